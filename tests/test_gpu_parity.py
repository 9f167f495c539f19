"""-m gpu : the CUDA day step (through the C ABI, libseir_b200.so) against the CPU oracle under the
SAME keyed Philox draws, bit-exact: nine histories, ten per-agent vectors, per-day per-age tallies.
Populations / parameters come from the golden fixtures frozen from the reference."""
import numpy as np
import pytest

import util
from covid19_demography_b200 import tables as tables_mod

pytestmark = pytest.mark.gpu

CASES = ["italy_default", "italy_seeded", "china_default", "italy_policy", "italy_fastiso", "italy_tracing"]


def _compare(eng, tup, ex, c, T, replica=0):
    packed = eng.history("packed", replica=replica)
    want = util.pack_history(tup[:9])
    if not np.array_equal(packed, want):
        bad = np.argwhere(packed != want)
        t, i = bad[0]
        raise AssertionError("history differs at %d agent-days; first (t=%d, i=%d): gpu=%#x oracle=%#x"
                             % (len(bad), t, i, packed[t, i], want[t, i]))
    for k in util.VECS:
        got = eng.agent_vector(k, replica=replica)
        exp = tup[util.VEC_INDEX[k]]
        assert np.array_equal(got, exp, equal_nan=True), "vector %s: %d mismatches" % (k, (got != exp).sum())
    for k in ("time_to_isolate", "time_severe", "time_to_critical"):
        got = eng.agent_vector(k, replica=replica)
        exp = np.where(ex[k] >= 65535, np.inf, ex[k])  # the device record saturates never-reached times
        assert np.array_equal(got, exp, equal_nan=True), "vector %s" % k
    assert np.array_equal(eng.agent_vector("num_infected_by_outside", replica=replica),
                          ex["num_infected_by_outside"].astype(np.float64))
    tal = eng.tallies(replica)
    assert np.array_equal(tal, util.tallies_from_history(tup[:9], c.age, T)), "per-day per-age tallies differ"
    for k, name in enumerate(util.HIST):
        plane = eng.history(name, replica=replica)
        assert np.array_equal(plane, np.asarray(tup[k])), name


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["native", "replay"])
def test_gpu_equals_keyed_oracle(gpu_engine_module, name, mode):
    c = util.Case(name)
    params = c.no_tracing_params()
    tup, ex, home, init = c.keyed_oracle(mode, params)
    util.check_invariants(tup, c.n, int(params["T"]))
    eng = c.gpu_engine(gpu_engine_module, params, mode)
    eng.run([c.seed], [home], [init])
    _compare(eng, tup, ex, c, int(params["T"]))
    cnt = eng.counters()
    assert cnt["infections"] == int((tup[15] > 0).sum())
    assert cnt["launches"] >= 2
    eng.close()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("sparse", ["never", "mixed"])
def test_staged_and_warp_cooperative_passes_agree(gpu_engine_module, monkeypatch, name, sparse):
    """The fixtures are small: by default every day of a native run is a warp-cooperative pass (sparse_entry).  The same
    runs through the staged pass only, and switching between the two from day to day (lists above 40 entries are staged;
    the holes a warp-cooperative day leaves in the transmitting list are then skipped and squeezed out)."""
    if sparse == "never":
        monkeypatch.setenv("SEIR_SPARSE_ROUNDS", "0")
    else:
        monkeypatch.setenv("SEIR_SPARSE_MAX", "40")
    c = util.Case(name)
    params = c.no_tracing_params()
    tup, ex, home, init = c.keyed_oracle("native", params)
    for launch in ("persistent", "per_day"):
        eng = c.gpu_engine(gpu_engine_module, params, "native", launch=launch)
        eng.run([c.seed], [home], [init])
        _compare(eng, tup, ex, c, int(params["T"]))
        assert eng.counters()["infections"] == int((tup[15] > 0).sum())
        eng.close()


@pytest.mark.parametrize("mode", ["native", "replay"])
def test_per_day_launch_equals_persistent(gpu_engine_module, mode):
    c = util.Case("italy_seeded")
    params = c.no_tracing_params()
    tup, ex, home, init = c.keyed_oracle(mode, params)
    eng = c.gpu_engine(gpu_engine_module, params, mode, launch="per_day")
    eng.run([c.seed], [home], [init])
    _compare(eng, tup, ex, c, int(params["T"]))
    assert eng.counters()["launches"] >= int(params["T"])
    eng.close()


def test_replicas_are_independent_runs(gpu_engine_module):
    """Seed ensembles: R replicas in one launch == R single runs (no communication)."""
    c = util.Case("italy_policy")
    params = c.no_tracing_params()
    seeds = [5, 6, 7]
    eng = c.gpu_engine(gpu_engine_module, params, "native", n_replicas=len(seeds))
    runs = [c.keyed_oracle("native", params, seed=s) for s in seeds]
    eng.run(seeds, [r[2] for r in runs], [r[3] for r in runs])
    for r, (tup, ex, _, _) in enumerate(runs):
        _compare(eng, tup, ex, c, int(params["T"]), replica=r)
    # rerun on the same handle: state is fully re-initialised
    eng.run(seeds[::-1], [r[2] for r in runs[::-1]], [r[3] for r in runs[::-1]])
    _compare(eng, runs[0][0], runs[0][1], c, int(params["T"]), replica=2)
    eng.close()


@pytest.mark.parametrize("launch", ["persistent", "per_day"])
def test_uneven_replicas(gpu_engine_module, launch):
    """An ensemble whose replicas differ wildly in size (no initial case, one, five, 50, 500): the CTAs are dealt to the
    replicas anew in every pass, in proportion to their active lists -- every replica still equals its own single run."""
    import oracle
    c = util.Case("italy_seeded")
    params = c.no_tracing_params()
    rng = np.random.default_rng(3)
    inits = [np.sort(rng.choice(c.n, size=k, replace=False)).astype(np.int64) for k in (0, 1, 5, 50, 500)]
    seeds = [41, 42, 43, 44, 45]
    home = np.zeros(c.n, dtype=bool)
    eng = c.gpu_engine(gpu_engine_module, params, "native", launch=launch, n_replicas=len(seeds))
    eng.run(seeds, [home] * len(seeds), inits)
    for r, (s, init) in enumerate(zip(seeds, inits)):
        tup, ex = oracle.run_model(s, *c.model_args(params), provider="native", tracing_enabled=1, home_real=home,
                                   initial_infected=init, tables=c.tables(params))
        _compare(eng, tup, ex, c, int(params["T"]), replica=r)
    assert eng.tallies(0)[:, 0].sum(axis=1).tolist() == [c.n] * int(params["T"])  # (nobody is ever infected in replica 0)
    eng.close()


def test_draw_contract_on_device(gpu_engine_module):
    """include/seir_draws.h evaluated by the device == by the host compiler, bit for bit."""
    import ctypes as C
    import oracle
    m = 100000
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.random(m // 2), np.exp(rng.uniform(-700, 700, m - m // 2))])
    e, l, p, g = gpu_engine_module.probe_draws(1234, m, 4.6, 1.621, 0.418, float(np.exp(-0.4)), x)
    lib = oracle.lib()
    y = np.empty(m)
    lib.oracle_sd_log(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_int64(m))
    assert np.array_equal(g, y)
    lib.oracle_sd_lognormal_days(C.c_uint64(1234), C.c_int64(m), C.c_double(1.621), C.c_double(0.418),
                                 C.c_void_p(y.ctypes.data))
    assert np.array_equal(l, y)
    yi = np.empty(m, dtype=np.int64)
    lib.oracle_sd_poisson(C.c_uint64(1234), C.c_int64(m), C.c_double(float(np.exp(-0.4))), C.c_void_p(yi.ctypes.data))
    assert np.array_equal(p, yi)
    assert np.isfinite(e).all() and e.mean() == pytest.approx(4.6, rel=0.02)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["native", "replay"])
def test_gpu_equals_keyed_oracle_with_tracing_as_compiled(gpu_engine_module, name, mode):
    """The fixtures' own parameters: as compiled, the reference switches contact tracing on at t_tracing_start
    (day 37 of 60 for the Italy defaults, day 10 in italy_tracing).  do_contact_tracing
    (seir_individual.py:60-88) on the device == the oracle's sequential recursion, bit for bit."""
    c = util.Case(name)
    params = dict(c.params)
    T = int(params["T"])
    tup, ex, home, init = c.keyed_oracle(mode, params)
    tracing_ran = 1 <= params["t_tracing_start"] < T
    util.check_invariants(tup, c.n, T, tracing_off=not tracing_ran)
    eng = c.gpu_engine(gpu_engine_module, params, mode)
    eng.run([c.seed], [home], [init])
    _compare(eng, tup, ex, c, T)
    eng.close()


@pytest.mark.parametrize("launch", ["persistent", "per_day"])
def test_tracing_with_replicas_and_launch_modes(gpu_engine_module, launch):
    c = util.Case("italy_tracing")
    params = dict(c.params)
    seeds = [21, 22, 23]
    runs = [c.keyed_oracle("native", params, seed=s) for s in seeds]
    assert any(np.asarray(r[0][8])[-1][np.asarray(r[0][6])[-1] | np.asarray(r[0][7])[-1]].any() for r in runs), \
        "the tracing fixture should quarantine some recovered/dead agents (SURVEY.md section 4)"
    eng = c.gpu_engine(gpu_engine_module, params, "native", launch=launch, n_replicas=len(seeds))
    eng.run(seeds, [r[2] for r in runs], [r[3] for r in runs])
    for k, (tup, ex, _, _) in enumerate(runs):
        _compare(eng, tup, ex, c, int(params["T"]), replica=k)
    eng.close()


def test_tracing_with_a_full_edge_cache(gpu_engine_module, monkeypatch):
    """The parallel tracer keeps the day's successful edges for the searches of the roots; with room for only 40 of them
    most agents fall back to evaluating their edges from the log again -- same result."""
    monkeypatch.setenv("SEIR_DEBUG_TEDGE_CAP", "40")
    c = util.Case("italy_tracing")
    params = dict(c.params)
    tup, ex, home, init = c.keyed_oracle("native", params)
    eng = c.gpu_engine(gpu_engine_module, params, "native")
    eng.run([c.seed], [home], [init])
    _compare(eng, tup, ex, c, int(params["T"]))
    eng.close()


def test_tracing_off_is_the_papers_semantics(gpu_engine_module):
    """tracing_enabled = 0: `tracing_enabled` False at :241, whatever t_tracing_start says."""
    c = util.Case("italy_tracing")
    params = dict(c.params)
    tup, ex, home, init = c.keyed_oracle("native", params, tracing_enabled=0)
    util.check_invariants(tup, c.n, int(params["T"]), tracing_off=True)
    eng = c.gpu_engine(gpu_engine_module, params, "native", tracing_enabled=0)
    eng.run([c.seed], [home], [init])
    _compare(eng, tup, ex, c, int(params["T"]))
    eng.close()


@pytest.mark.parametrize("tracing", [False, True])
def test_more_replicas_than_ctas(gpu_engine_module, tracing):
    """Sweeps at small n: 320 replicas on a grid of 296 CTAs (every CTA serves one or two replicas per pass; with
    tracing, the parallel tracer loops over them too)."""
    c = util.Case("italy_tracing" if tracing else "italy_policy")
    params = dict(c.params) if tracing else c.no_tracing_params()
    R = 320
    seeds = list(range(3000, 3000 + R))
    eng = c.gpu_engine(gpu_engine_module, params, "native", n_replicas=R)
    pro = [tables_mod.prologue(s, c.age_groups, c.n, c.fraction_stay_home, params["initial_infected_fraction"]) for s in seeds]
    eng.run(seeds, [p[0] for p in pro], [p[1] for p in pro])
    for k in (0, 1, 150, 295, 296, 319):
        tup, ex, _, _ = c.keyed_oracle("native", params, seed=seeds[k])
        _compare(eng, tup, ex, c, int(params["T"]), replica=k)
    eng.close()


def test_per_agent_household_infectiousness(gpu_engine_module):
    """p_infect_household is a per-agent vector in the reference's signature (:108); the drivers pass a constant
    one (detected and kept as a scalar), a varying one takes the vector path."""
    c = util.Case("italy_seeded")
    rng = np.random.default_rng(11)
    c.p_infect_household = rng.uniform(0.0, 0.25, size=c.n)
    params = c.no_tracing_params()
    for mode in ("native", "replay"):
        tup, ex, home, init = c.keyed_oracle(mode, params)
        eng = c.gpu_engine(gpu_engine_module, params, mode)
        eng.run([c.seed], [home], [init])
        _compare(eng, tup, ex, c, int(params["T"]))
        eng.close()


def test_degenerate_runs(gpu_engine_module):
    """No initial case: everybody stays S.  T = 2: one day.  Everybody infected at t = 0."""
    c = util.Case("italy_default")
    params = c.no_tracing_params(initial_infected_fraction=0.0)
    tup, ex, home, init = c.keyed_oracle("native", params)
    assert len(init) == 0
    eng = c.gpu_engine(gpu_engine_module, params, "native")
    eng.run([c.seed], [home], [init])
    _compare(eng, tup, ex, c, int(params["T"]))
    assert eng.tallies()[:, 0].sum(axis=1).tolist() == [c.n] * int(params["T"])
    eng.close()
    for over in ({"T": 2.0, "initial_infected_fraction": 0.05}, {"T": 12.0, "initial_infected_fraction": 1.0}):
        params = c.no_tracing_params(**over)
        tup, ex, home, init = c.keyed_oracle("native", params)
        eng = c.gpu_engine(gpu_engine_module, params, "native")
        eng.run([c.seed], [home], [init])
        _compare(eng, tup, ex, c, int(params["T"]))
        eng.close()


def test_bad_inputs_raise(gpu_engine_module):
    c = util.Case("italy_default")
    params = c.no_tracing_params()
    hh = c.households.copy()
    hh[0, 0] = 7  # break the contiguous layout
    c2 = util.Case("italy_default")
    c2.households = hh
    with pytest.raises((RuntimeError, ValueError), match="contiguous"):
        c2.gpu_engine(gpu_engine_module, params, "native")
    eng = c.gpu_engine(gpu_engine_module, params, "native")
    with pytest.raises((RuntimeError, ValueError)):
        eng.run([0], [None], [np.array([c.n + 5])])
    eng.close()
