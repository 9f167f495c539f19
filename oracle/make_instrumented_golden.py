"""A direct link  reference's own code == keyed draw contract  (run in the build container only).

    python oracle/make_instrumented_golden.py            # writes tests/golden/ref_keyed_*.npz

BASELINE.json's first correctness criterion: "when the GPU kernels are fed the same pre-drawn uniform random streams as
an instrumented reference run on a small population, the daily integer state trajectories and per-age counts must be
bit-exact".  The other fixtures pin  oracle(MT stream) == unmodified reference  and the GPU is compared with
oracle(keyed draws); this script closes the triangle without the oracle's control flow in between:

  * it reads `/root/reference/seir_individual.py` WHERE IT LIES and writes an instrumented copy into a scratch
    directory (never into this repo): every `np.random.*` / `functions.threshold_*` call of the day loop, of
    `do_contact_tracing` and of the initial cases' activation draw is replaced, at its own source line, by the keyed
    draw of that site (include/seir_draws.h: same agent, day, site, sub-counter as the contract names), bound through
    ctypes from `liboracle_seir.so` and called from numba's nopython code.  Nothing else of the reference changes:
    its statement order, its `continue`s, its row copies, its recursion, its as-compiled tracing switch, and its own
    Mersenne-Twister prologue (`np.random.choice(..., replace=False)` at :181 / :187) all run as the reference wrote them;
  * the table of edits below is the whole instrumentation: (line, text that must be there, replacement) -- a line
    that does not hold the expected text aborts the script, so a different reference cannot be instrumented wrongly;
  * the instrumented `run_model` runs under numba on the populations of the frozen fixtures (`tests/golden/ref_*.npz`,
    tracing as compiled, i.e. from the fixture's own `t_tracing_start`) and all 20 outputs are frozen as
    `tests/golden/ref_keyed_<case>.npz`.

`tests/test_instrumented_reference.py` then demands, bit for bit: oracle(replay) == these outputs (CPU, plus a live
regeneration where /root/reference exists) and GPU(replay mode, through the C ABI) == these outputs (`-m gpu`).
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# draw sites of include/seir_draws.h
ACT, DAY, SEV, CRIT, HH, INF, POIS, KEEP, TR_HH, TR_OUT, INIT = 1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 13
_ORD_COMM = "(16 + (contact_age << 8) + j) << 8"   # SD_ORD_COMM_REPLAY(contact_age, j) << 8
_ORD_HH = "j << 8"                                  # SD_ORD_HH(j) << 8

RAND = "np.random.rand()"
EXPO = "functions.threshold_exponential("
LOGN = "functions.threshold_log_normal("


def _u(site, sub, lane):
    return "kd_uniform(kd_seed, i, t, %d, %s, %s)" % (site, sub, lane)


def _e(site, sub, lane):
    return "kd_exp(kd_seed, i, t, %d, %s, %s, " % (site, sub, lane)


# (line of /root/reference/seir_individual.py, text that must be on it, what it becomes)
EDITS = [
    (61, "def do_contact_tracing(i, ", "def do_contact_tracing(kd_seed, i, "),
    (72, RAND, _u(TR_HH, "j", 0)),
    (77, "do_contact_tracing(contact, ", "do_contact_tracing(kd_seed, contact, "),
    (83, RAND, _u(TR_OUT, "j", 0)),
    (88, "do_contact_tracing(contact, ", "do_contact_tracing(kd_seed, contact, "),
    (91, "fraction_stay_home, params):", "fraction_stay_home, params, kd_enlam_pre, kd_enlam_post):"),
    (163, "np.random.seed(seed)", "np.random.seed(seed); kd_seed = seed; kd_enlam = kd_enlam_pre"),
    (227, LOGN, "kd_lognormal(kd_seed, initial_infected[i], 0, %d, 0, " % INIT),
    (240, "contact_matrix = contact_matrix/lockdown_factor",
     "contact_matrix = contact_matrix/lockdown_factor; kd_enlam = kd_enlam_post"),
    (262, RAND, _u(ACT, 0, 0)),
    (263, EXPO, _e(ACT, 0, 1)),
    (267, EXPO, _e(ACT, 0, 1)),
    (270, EXPO, _e(ACT, 1, 0)),
    (282, RAND, _u(DAY, 0, 0)),
    (289, "do_contact_tracing(i, ", "do_contact_tracing(kd_seed, i, "),
    (303, "do_contact_tracing(i, ", "do_contact_tracing(kd_seed, i, "),
    (306, RAND, _u(SEV, 0, 0)),
    (307, EXPO, _e(SEV, 0, 1)),
    (310, EXPO, _e(SEV, 0, 1)),
    (316, RAND, _u(CRIT, 0, 0)),
    (317, EXPO, _e(CRIT, 0, 1)),
    (320, EXPO, _e(CRIT, 0, 1)),
    (346, RAND, _u(HH, "j >> 1", "j & 1")),
    (352, EXPO, _e(INF, _ORD_HH, 0)),
    (356, LOGN, "kd_lognormal(kd_seed, i, t, %d, %s, " % (INF, _ORD_HH)),
    (370, "np.random.poisson(contact_matrix[age[i], contact_age])",
     "kd_poisson(kd_seed, i, t, %d, contact_age << 8, kd_enlam[age[i], contact_age])" % POIS),
    (373, RAND, _u(KEEP, "(contact_age << 8) | j", 0)),
    (374, "np.random.choice(age_groups[contact_age])",
     "age_groups[contact_age][kd_index(kd_seed, i, t, %d, (contact_age << 8) | j, age_groups[contact_age].shape[0])]" % KEEP),
    (381, EXPO, _e(INF, _ORD_COMM, 0)),
    (385, LOGN, "kd_lognormal(kd_seed, i, t, %d, %s, " % (INF, _ORD_COMM)),
]

HEADER = '''
import ctypes as _C
_kd = _C.CDLL(%r)
def _kd_bind(name, restype, *argtypes):
    f = getattr(_kd, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f
_K = (_C.c_uint64, _C.c_uint32, _C.c_uint32, _C.c_uint32, _C.c_uint32)   # seed, agent, day, site, sub
kd_uniform = _kd_bind("oracle_kd_uniform", _C.c_double, *_K, _C.c_int)
kd_exp = _kd_bind("oracle_kd_exp_days", _C.c_double, *_K, _C.c_int, _C.c_double)
kd_lognormal = _kd_bind("oracle_kd_lognormal_days", _C.c_double, *_K, _C.c_double, _C.c_double)
kd_poisson = _kd_bind("oracle_kd_poisson", _C.c_int64, *_K, _C.c_double)
kd_index = _kd_bind("oracle_kd_index", _C.c_int64, *_K, _C.c_int64)
'''


def instrumented_source(lib_path):
    """The reference's module text with the EDITS applied (one replacement per listed line, checked)."""
    lines = open(os.path.join(ref_shim.REFERENCE_ROOT, "seir_individual.py")).read().split("\n")
    for ln, old, new in EDITS:
        text = lines[ln - 1]
        if text.count(old) != 1:
            raise RuntimeError("seir_individual.py:%d does not hold %r exactly once: %r" % (ln, old, text))
        lines[ln - 1] = text.replace(old, new)
    left = [k + 1 for k, text in enumerate(lines)
            if 45 <= k and ("np.random." in text or "functions.threshold" in text) and not text.lstrip().startswith("#")]
    # what stays on the reference's own stream: the seed call and the two prologue picks
    if left != [163, 181, 187]:
        raise RuntimeError("draw sites left uninstrumented: %r" % left)
    first_def = next(k for k, text in enumerate(lines) if text.startswith('"""3. RUN SIMULATION'))
    lines.insert(first_def, HEADER % lib_path)
    return "\n".join(lines)


_module = None


def load_instrumented():
    """Write the instrumented copy into a scratch directory and import it (numba compiles it on first call)."""
    global _module
    if _module is None:
        import oracle  # oracle/oracle.py: builds liboracle_seir.so
        lib_path = oracle.build()
        ref_shim.load()  # numpy aliases, reference root on sys.path (for `functions`, `sample_*`), scratch CWD
        scratch = tempfile.mkdtemp(prefix="seir_keyed_ref_")
        path = os.path.join(scratch, "seir_individual_keyed.py")
        with open(path, "w") as f:
            f.write(instrumented_source(lib_path))
        spec = importlib.util.spec_from_file_location("seir_individual_keyed", path)
        _module = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = _module
        spec.loader.exec_module(_module)
    return _module


def typed_params(params):
    import numba
    d = numba.typed.Dict.empty(key_type=numba.types.unicode_type, value_type=numba.types.float64)
    for k, v in params.items():
        d[k] = float(v)
    return d


def run_case(case, params=None, seed=None):
    """The instrumented reference on a fixture's population (util.Case): the 20-tuple of :392."""
    mod = load_instrumented()
    params = dict(case.params if params is None else params)
    tb = case.tables(params)   # the product's host tables: exp(-C) before / after the lockdown
    seed = case.seed if seed is None else seed
    return mod.run_model(int(seed), *case.model_args(typed_params(params)),
                         np.ascontiguousarray(tb["enlam_pre"]), np.ascontiguousarray(tb["enlam_post"]))


def freeze(name):
    import util  # tests/util.py
    c = util.Case(name)
    out = run_case(c)
    T = int(c.params["T"])
    fx = {"seed": np.int64(c.seed), "history_packed": util.pack_history(out[:9])}
    for k in util.VECS:
        fx["out_" + k] = np.asarray(out[util.VEC_INDEX[k]])
    assert np.array_equal(np.asarray(out[17]), c.age)
    path = os.path.join(GOLDEN, "ref_keyed_%s.npz" % name)
    np.savez_compressed(path, **fx)
    hp = fx["history_packed"]
    print(name, "n", c.n, "T", T, "final S/E/M/Sev/C/R/D", [int(((hp[-1] & 7) == k).sum()) for k in range(7)],
          "Q", int(((hp[-1] >> 3) & 1).sum()), "Doc", int(((hp[-1] >> 4) & 1).sum()),
          "%.0f kB" % (os.path.getsize(path) / 1e3))


LARGE = ("ensemble_tracing_italy_n20000", 9)   # (fixture whose population is used, seed): 20 000 agents, tracing from day 25


def large_summary(out, age, T):
    """What is frozen of a large run: per-day per-age counts, a digest of the packed history, the ten vectors (their
    values are small integers, -1 and inf: exact in float32)."""
    import hashlib
    import util
    packed = util.pack_history(out[:9])
    fx = {"tallies": util.tallies_from_history(out[:9], age, T).astype(np.int32),
          "history_sha256": np.array(hashlib.sha256(np.ascontiguousarray(packed).tobytes()).hexdigest())}
    for k in util.VECS:
        v = np.asarray(out[util.VEC_INDEX[k]])
        assert np.array_equal(v.astype(np.float32).astype(np.float64), v, equal_nan=True)
        fx["out_" + k] = v.astype(np.float32)
    return fx


def freeze_large():
    """One run large enough for the staged (dense-day) pass to run over many CTAs: thousands of list entries a day."""
    import util
    name, seed = LARGE
    c = util.Case(name)
    out = run_case(c, seed=seed)
    T = int(c.params["T"])
    fx = large_summary(out, c.age, T)
    fx["seed"] = np.int64(seed)
    path = os.path.join(GOLDEN, "ref_keyed_large_italy_n20000.npz")
    np.savez_compressed(path, **fx)
    tal = fx["tallies"]
    print("large: n", c.n, "T", T, "infected", int(c.n - tal[-1, 0].sum()), "peak active", int(tal[:, [1, 2, 4, 5]].sum(axis=(1, 2)).max()),
          "Q end", int(tal[-1, 8].sum()), "%.0f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    import util
    which = sys.argv[1:] or (list(util.CASES) + ["large"])
    for name in which:
        if name == "large":
            freeze_large()
        else:
            freeze(name)
