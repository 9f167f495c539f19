"""The CPU oracle (MT provider) against outputs frozen from the UNMODIFIED reference.

The reference has no tests or golden vectors; the fixtures under tests/golden/ were produced by
running its own numba `run_model` in the build container (oracle/make_golden.py).  All 20 outputs of
`seir_individual.py:392` must be bit-identical -- including the runs in which contact tracing is
active (as compiled, tracing switches on at t_tracing_start)."""
import numpy as np
import pytest

import oracle
import util


@pytest.mark.parametrize("name", util.CASES)
def test_oracle_mt_equals_reference_fixture(name):
    c = util.Case(name)
    tup, ex = oracle.run_model(c.seed, *c.model_args(), provider="mt", tracing_enabled=1)
    want = util.unpack_history(c.fx["history_packed"])
    for k, a, b in zip(util.HIST, want, tup[:9]):
        assert np.array_equal(a, b), "history %s differs from the reference" % k
    for k in util.VECS:
        got = tup[util.VEC_INDEX[k]]
        assert np.array_equal(c.fx["out_" + k], got, equal_nan=True), "vector %s differs from the reference" % k
    assert np.array_equal(tup[17], c.age)


@pytest.mark.parametrize("name", ["italy_default", "italy_policy"])
def test_host_prologue_equals_reference_stream(name):
    """Home_real / initial_infected from NumPy's legacy RandomState == the reference's numba MT stream
    (restated in the oracle, which is pinned above)."""
    from covid19_demography_b200 import tables
    c = util.Case(name)
    _, ex = oracle.run_model(c.seed, *c.model_args(), provider="mt", tracing_enabled=1)
    home, init = tables.prologue(c.seed, c.age_groups, c.n, c.fraction_stay_home,
                                 c.params["initial_infected_fraction"])
    assert np.array_equal(home, ex["home_real"])
    assert np.array_equal(init, ex["initial_infected"])
    if name == "italy_policy":
        assert home.sum() > 0


def test_oracle_mt_generator_is_numpys_mt19937():
    import ctypes as C
    lib = oracle.lib()
    y = np.empty(2000)
    lib.oracle_mt_doubles(C.c_uint32(12345), C.c_void_p(y.ctypes.data), C.c_int64(2000))
    assert np.array_equal(y, np.random.RandomState(12345).random_sample(2000))
