"""CPU tier: `seir_prologue_mt` (host routine of libseir_b200.so, include/seir_b200.h) makes the reference's own prologue
picks for a seed (seir_individual.py:163, 176-187) -- checked against `tables.prologue` (NumPy's RandomState, the same
MT19937 shuffle as numba's) on ragged cases, and against the picks the reference itself made in the frozen fixtures
(through the oracle's MT provider, which is pinned bit for bit on the reference's outputs)."""
import numpy as np
import pytest

import oracle
import util
from covid19_demography_b200 import engine, tables


def _groups(age):
    return tuple(np.where(age == a)[0] for a in range(101))


@pytest.mark.parametrize("n", [1, 2, 3, 10, 101, 1000, 30000])
def test_equals_the_numpy_replay(n):
    rng = np.random.default_rng(n)
    age = rng.integers(0, 101, n)
    if n > 50:
        age[age == 7] = 8   # an empty age group (:180 skips it without touching the stream)
    groups = _groups(age)
    csr = engine.group_csr(groups)
    policies = [np.zeros(101), np.where(np.arange(101) >= 60, 0.8, 0.0), rng.random(101), np.full(101, 0.01),
                np.ones(101)]
    for seed in (0, 1, 77, 2**32 - 1):
        for frac in (0.0, 0.001, 0.01, 0.2, 0.6, 1.0):
            for fsh in policies:
                h0, i0 = tables.prologue(seed, groups, n, fsh, frac)
                h1, i1 = engine.prologue_mt(seed, csr, n, fsh, frac)
                assert np.array_equal(h0, h1) and np.array_equal(i0, i1), (seed, frac)
                assert h1.dtype == np.bool_ and i1.dtype == np.int64 and len(i1) == int(frac * n)


@pytest.mark.parametrize("name", ["italy_seeded", "italy_policy", "china_default"])
def test_equals_the_picks_of_the_reference_run(name):
    c = util.Case(name)
    tup, ex = oracle.run_model(c.seed, *c.model_args(), provider="mt", tracing_enabled=1)
    # the MT provider reproduces the reference's 20 outputs bit for bit (test_oracle_golden.py): its picks are the reference's
    assert np.array_equal(util.pack_history(tup[:9]), c.fx["history_packed"])
    home, init = engine.prologue_mt(c.seed, engine.group_csr(c.age_groups), c.n, c.fraction_stay_home,
                                    c.params["initial_infected_fraction"])
    assert np.array_equal(home, ex["home_real"].astype(bool))
    assert np.array_equal(init, ex["initial_infected"])
    if name == "italy_policy":
        assert home.sum() > 100   # the 60+ shelter in place in this fixture


def test_argument_errors():
    age = np.arange(50) % 101
    csr = engine.group_csr(_groups(age))
    with pytest.raises(ValueError):
        engine.prologue_mt(0, csr, 49, None, 0.0)            # the groups do not partition 49 agents
    with pytest.raises(ValueError):
        engine.prologue_mt(0, csr, 50, np.zeros(7), 0.0)     # one fraction per age
    with pytest.raises(ValueError):
        engine.prologue_mt(0, csr, 50, np.full(101, np.nan), 0.0)
    with pytest.raises(ValueError):
        engine.prologue_mt(0, csr, 50, None, -0.5)
    with pytest.raises(RuntimeError, match="above 1"):
        engine.prologue_mt(0, csr, 50, np.full(101, 2.5), 0.0)
    with pytest.raises(RuntimeError, match="more initial cases"):
        engine.prologue_mt(0, csr, 50, None, 1.5)
