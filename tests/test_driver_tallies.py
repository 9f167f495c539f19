"""The drivers' per-run statistics (run_simulation.py:448-460, parameter_sweep.py:556-603) from the fused
tallies + exposure-day cohorts == the same statistics computed the drivers' way (NumPy passes over the nine
bool[T, n] histories and the per-agent vectors).  CPU tier: tallies / cohorts built from oracle outputs;
gpu tier: from the device (`seir_get_tallies`, `seir_get_cohorts`)."""
import warnings

import numpy as np
import pytest

import util
from covid19_demography_b200 import driver_tallies as dt


def cohorts_from_outputs(tup, age, T):
    """[T, 4, 101] the slow way."""
    te, nib = tup[15], tup[9]
    D_end, R_end = np.asarray(tup[7])[-1], np.asarray(tup[6])[-1]
    out = np.zeros((T, 4, 101), dtype=np.int64)
    for t in range(T):
        m = te == t
        out[t, 0] = np.bincount(age[m], minlength=101)
        out[t, 1] = np.bincount(age[m], weights=nib[m], minlength=101).astype(np.int64)
        out[t, 2] = np.bincount(age[m & D_end], minlength=101)
        out[t, 3] = np.bincount(age[m & R_end], minlength=101)
    return out


def drivers_way(tup, age, n, T):
    """Restatement of the drivers' statements on the full arrays."""
    S, E, Mild, Documented, Severe, Critical, R, D, Q = [np.asarray(a) for a in tup[:9]]
    nib, te = tup[9], tup[15]
    out = {k: np.zeros(T) for k in ("r_0_over_time", "cfr_over_time", "fraction_over_70_time",
                                    "fraction_below_30_time", "median_age_time")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(T):
            m = te == t
            if m.sum() > 0:
                out["r_0_over_time"][t] = nib[m].mean()
                out["cfr_over_time"][t] = D[-1, m].sum() / (D[-1, m].sum() + R[-1, m].sum())
                out["fraction_over_70_time"][t] = (age[m] >= 70).mean()
                out["fraction_below_30_time"][t] = (age[m] < 30).mean()
                out["median_age_time"][t] = np.median(age[m])
        out["total_infections_time"] = n - S.sum(axis=1) - E.sum(axis=1)
        out["total_documented_time"] = Documented.sum(axis=1)
        out["total_deaths_time"] = D.sum(axis=1)
        out["dead_by_age"] = np.array([D[-1, age == a].sum() for a in range(101)])
        out["r0_total"] = nib[np.logical_and(te <= 20, te > 0)].mean()
        G = len(dt.AGE_GROUPS_SWEEP)
        out["infected_by_age_by_time"] = np.zeros((G, T))
        out["CFR_by_age_by_time"] = np.zeros((G, T))
        out["dead_by_age_by_time"] = np.zeros((G, T))
        for t in range(T):
            for g, (lo, hi) in enumerate(dt.AGE_GROUPS_SWEEP):
                grp = np.logical_and(age >= lo, age <= hi)
                out["infected_by_age_by_time"][g, t] = grp.sum() - np.logical_and(S[t], grp).sum()
                d = np.logical_and(np.logical_and(D[-1], te == t), grp).sum()
                r = np.logical_and(np.logical_and(R[-1], te == t), grp).sum()
                out["CFR_by_age_by_time"][g, t] = d / (d + r)
                out["dead_by_age_by_time"][g, t] = np.logical_and(D[t], grp).sum()
    return out


def check(tallies, cohorts, want, n):
    got = dt.cohort_statistics(cohorts)
    for k in ("r_0_over_time", "cfr_over_time", "fraction_over_70_time", "fraction_below_30_time", "median_age_time"):
        assert np.allclose(got[k], want[k], rtol=0, atol=1e-12, equal_nan=True), k
    tot = dt.totals(tallies, n)
    for k in ("total_infections_time", "total_documented_time", "total_deaths_time", "dead_by_age"):
        assert np.array_equal(tot[k], want[k]), k
    assert dt.r0_total(cohorts) == pytest.approx(want["r0_total"], abs=1e-12)
    ag = dt.by_age_group(tallies, cohorts)
    for k in ("infected_by_age_by_time", "CFR_by_age_by_time", "dead_by_age_by_time"):
        assert np.allclose(ag[k], want[k], rtol=0, atol=1e-12, equal_nan=True), k
    pt = dt.per_time(tallies)
    assert np.array_equal(pt["S"] + pt["E"] + pt["Mild"] + pt["Severe"] + pt["Critical"] + pt["R"] + pt["D"],
                          np.full(len(pt["S"]), n))


@pytest.mark.parametrize("name", ["italy_seeded", "italy_tracing"])
def test_driver_statistics_from_tallies(name):
    c = util.Case(name)
    T = int(c.params["T"])
    tup, _, _, _ = c.keyed_oracle("native", c.params)
    check(util.tallies_from_history(tup[:9], c.age, T), cohorts_from_outputs(tup, c.age, T),
          drivers_way(tup, c.age, c.n, T), c.n)


def test_median_from_histogram():
    rng = np.random.default_rng(3)
    for m in (1, 2, 5, 6, 101):
        ages = rng.integers(0, 101, size=m)
        assert dt._median_from_histogram(np.bincount(ages, minlength=101)) == np.median(ages)
    assert np.isnan(dt._median_from_histogram(np.zeros(101)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["italy_seeded", "china_default"])
def test_device_cohorts_and_driver_statistics(gpu_engine_module, name):
    c = util.Case(name)
    T = int(c.params["T"])
    tup, _, home, init = c.keyed_oracle("native", c.params)
    eng = c.gpu_engine(gpu_engine_module, c.params, "native")
    eng.run([c.seed], [home], [init])
    coh = eng.cohorts()
    assert np.array_equal(coh, cohorts_from_outputs(tup, c.age, T))
    check(eng.tallies(), coh, drivers_way(tup, c.age, c.n, T), c.n)
    eng.close()
