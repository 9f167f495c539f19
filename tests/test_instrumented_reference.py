"""The reference's OWN code under the keyed draw contract (BASELINE.json: "fed the same pre-drawn uniform random
streams as an instrumented reference run on a small population ... bit-exact").

`tests/golden/ref_keyed_<case>.npz` holds the 20 outputs of `/root/reference/seir_individual.py`'s `run_model` after
`oracle/make_instrumented_golden.py` replaced each of its random call sites -- and nothing else -- by the keyed draw of
that site; its control flow, tracing recursion included, ran under numba as the reference wrote it.  Demanded here,
bit for bit on all six fixtures with the fixtures' own parameters (tracing as compiled):

  * CPU: oracle(provider = replay) == those outputs, and (where /root/reference exists) a live regeneration of the
    instrumented copy on a seed / parameter set that is not frozen;
  * `-m gpu`: the CUDA day step in replay mode, through the C ABI, == those outputs -- the GPU against the reference's
    code without the oracle in between -- plus the per-day per-age tallies the day loop counted against the column
    sums of the reference's histories.
"""
import os

import numpy as np
import pytest

import ref_shim
import util


def _frozen(name):
    return np.load(os.path.join(util.GOLDEN, "ref_keyed_%s.npz" % name))


@pytest.mark.parametrize("name", util.CASES)
def test_oracle_replay_equals_instrumented_reference(name):
    c = util.Case(name)
    fx = _frozen(name)
    assert int(fx["seed"]) == c.seed
    tup, ex, home, init = c.keyed_oracle("replay", dict(c.params))
    assert np.array_equal(util.pack_history(tup[:9]), fx["history_packed"])
    for k in util.VECS:
        assert np.array_equal(tup[util.VEC_INDEX[k]], fx["out_" + k], equal_nan=True), k
    # the frozen runs are not trivial: an epidemic happened, and where tracing ran it left its mark
    hp = fx["history_packed"]
    assert ((hp[-1] & 7) != 0).sum() > 100
    T = int(c.params["T"])
    if name in ("italy_default", "italy_seeded", "italy_tracing"):
        assert 1 <= c.params["t_tracing_start"] < T
        done = ((hp[-1] & 7) >= 5)
        assert (((hp[-1] >> 3) & 1).astype(bool) & done).any(), "tracing should quarantine some recovered/dead agents"


def _numba_ok():
    try:
        import numba  # noqa: F401
        return True
    except Exception:
        return False


@pytest.mark.skipif(not (ref_shim.available() and _numba_ok()), reason="/root/reference or numba is not on this box")
def test_live_instrumented_reference_on_an_unfrozen_run():
    """Regenerate the instrumented copy from the reference where it lies and run a seed / parameter set that no
    fixture holds: heavy tracing from day 8, short isolation delays, a lockdown on day 20."""
    import make_instrumented_golden as mig
    c = util.Case("italy_tracing")
    params = dict(c.params)
    params.update(t_tracing_start=8.0, t_lockdown=20.0, lockdown_factor=3.0, mean_time_to_isolate_asympt=6.0,
                  initial_infected_fraction=0.01)
    seed = 77
    out = mig.run_case(c, params=params, seed=seed)
    tup, ex, home, init = c.keyed_oracle("replay", params, seed=seed)
    assert np.array_equal(util.pack_history(out[:9]), util.pack_history(tup[:9]))
    for k in util.VECS:
        assert np.array_equal(np.asarray(out[util.VEC_INDEX[k]]), tup[util.VEC_INDEX[k]], equal_nan=True), k
    assert (np.asarray(out[15]) > 0).sum() > 300


@pytest.mark.gpu
@pytest.mark.parametrize("name", util.CASES)
@pytest.mark.parametrize("launch", ["persistent", "per_day"])
@pytest.mark.parametrize("passes", ["default", "staged_only", "mixed"])
def test_gpu_replay_equals_instrumented_reference(gpu_engine_module, monkeypatch, name, launch, passes):
    # both pass forms of the day loop against the reference's code: the fixtures are small, so by default their days run
    # as warp-cooperative passes; staged_only / mixed force the staged (dense-day) pass on all / on the longer days
    if passes == "staged_only":
        monkeypatch.setenv("SEIR_SPARSE_ROUNDS", "0")
    elif passes == "mixed":
        monkeypatch.setenv("SEIR_SPARSE_MAX", "40")
    c = util.Case(name)
    fx = _frozen(name)
    params = dict(c.params)
    T = int(params["T"])
    # the reference's own prologue ran in the instrumented copy (MT picks of :181, :187): the host prologue of the
    # drop-in run_model (seir_prologue_mt) makes exactly those
    home, init = gpu_engine_module.prologue_mt(c.seed, gpu_engine_module.group_csr(c.age_groups), c.n,
                                               c.fraction_stay_home, params["initial_infected_fraction"])
    eng = c.gpu_engine(gpu_engine_module, params, "replay", launch=launch)
    eng.run([c.seed], [home], [init])
    packed = eng.history("packed")
    want = fx["history_packed"]
    if not np.array_equal(packed, want):
        bad = np.argwhere(packed != want)
        t, i = bad[0]
        raise AssertionError("history differs at %d agent-days; first (t=%d, i=%d): gpu=%#x reference=%#x"
                             % (len(bad), t, i, packed[t, i], want[t, i]))
    for k in util.VECS:
        assert np.array_equal(eng.agent_vector(k), fx["out_" + k], equal_nan=True), k
    assert np.array_equal(eng.tallies(0), util.tallies_from_history(util.unpack_history(want), c.age, T)), \
        "per-day per-age counts differ"
    eng.close()


# ---- one large run: 20 000 agents, 12 208 infections, thousands of list entries a day (the staged pass over many CTAs),
# tracing from day 25 as compiled.  Frozen: per-day per-age counts, a digest of the packed history, the ten vectors.
def _large():
    import make_instrumented_golden as mig
    name, seed = mig.LARGE
    return util.Case(name), seed, np.load(os.path.join(util.GOLDEN, "ref_keyed_large_italy_n20000.npz")), mig


def _check_large(fx, packed, tallies, vector_of):
    import hashlib
    assert np.array_equal(tallies, fx["tallies"].astype(np.int64)), "per-day per-age counts differ"
    assert hashlib.sha256(np.ascontiguousarray(packed).tobytes()).hexdigest() == str(fx["history_sha256"]), "history differs"
    for k in util.VECS:
        assert np.array_equal(vector_of(k), fx["out_" + k].astype(np.float64), equal_nan=True), k


def test_oracle_replay_equals_large_instrumented_run():
    c, seed, fx, _ = _large()
    assert int(fx["seed"]) == seed and c.n == 20000
    tup, ex, home, init = c.keyed_oracle("replay", dict(c.params), seed=seed)
    T = int(c.params["T"])
    _check_large(fx, util.pack_history(tup[:9]), util.tallies_from_history(tup[:9], c.age, T), lambda k: tup[util.VEC_INDEX[k]])
    assert c.n - fx["tallies"][-1, 0].sum() > 10000 and fx["tallies"][:, [1, 2, 4, 5]].sum(axis=(1, 2)).max() > 4000


@pytest.mark.gpu
@pytest.mark.parametrize("passes", ["default", "staged_only"])
def test_gpu_replay_equals_large_instrumented_run(gpu_engine_module, monkeypatch, passes):
    if passes == "staged_only":
        monkeypatch.setenv("SEIR_SPARSE_ROUNDS", "0")
    c, seed, fx, _ = _large()
    params = dict(c.params)
    home, init = gpu_engine_module.prologue_mt(seed, gpu_engine_module.group_csr(c.age_groups), c.n,
                                               c.fraction_stay_home, params["initial_infected_fraction"])
    eng = c.gpu_engine(gpu_engine_module, params, "replay")
    eng.run([seed], [home], [init])
    _check_large(fx, eng.history("packed"), eng.tallies(0), eng.agent_vector)
    assert eng.counters()["infections"] > 10000
    eng.close()
