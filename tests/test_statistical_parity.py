"""Statistical parity (tier T4; BASELINE.json north_star): with native keyed draws
the ensemble of daily infection / ICU / death curves must agree with the UNMODIFIED reference's ensemble.

Two reference ensembles (64 seeds each, one fixed 20 000-agent Italy population) were frozen by
oracle/make_golden.py::run_ensemble from the reference's own numba run_model: one in which tracing never
switches on, one in which the reference traces contacts from day 25 (as it does as compiled).  Tolerances, stated:
  * two-sample Kolmogorov-Smirnov on final deaths, peak ICU (Critical), final attack size, peak Mild, final
    Documented and final Q: p-value > 1e-3 for each (a 64-vs-64 KS test rejects a 15 % shift of the median
    at that level);
  * confidence-interval overlap on the daily means of cumulative infections, Critical and deaths at
    days 10, 20, ..., 59: |mean_a - mean_b| <= 4 combined standard errors + 1 agent.
The CPU tier checks the oracle's keyed providers (both community modes) this way; the gpu tier
checks the CUDA path (64 replicas in ONE persistent launch) the same way."""
import numpy as np
import pytest
import scipy.stats as st

import oracle
import util
from covid19_demography_b200 import ensemble, tables as tables_mod

N_SEEDS = 64
KS_P_MIN = 1e-3
CI_SIGMAS = 4.0


FIXTURES = {"no_tracing": "ensemble_italy_n20000",            # t_tracing_start = 1000: the paper's semantics
            "tracing": "ensemble_tracing_italy_n20000"}        # tracing from day 25, as the reference runs as compiled


def reference_ensemble(which="no_tracing"):
    c = util.Case(FIXTURES[which])
    return c, c.fx["curves"].astype(np.int64)  # [64, T, 9]


def assert_same_distribution(a, b, n, label):
    """a, b: [seeds, T, 9] per-day plane sums (order S,E,Mild,Documented,Severe,Critical,R,D,Q)."""
    stats = {
        "final deaths": lambda c: c[:, -1, 7],
        "peak ICU": lambda c: c[:, :, 5].max(axis=1),
        "final attack size": lambda c: n - c[:, -1, 0],
        "peak Mild": lambda c: c[:, :, 2].max(axis=1),
        "final Documented": lambda c: c[:, -1, 3],
        "final Q": lambda c: c[:, -1, 8],
    }
    for name, f in stats.items():
        p = st.ks_2samp(f(a), f(b)).pvalue
        assert p > KS_P_MIN, "%s: KS p=%.2e on %s (means %.1f vs %.1f)" % (label, p, name, f(a).mean(), f(b).mean())
    T = a.shape[1]
    for day in list(range(10, T, 10)) + [T - 1]:
        for name, col in (("infections", lambda c: n - c[:, day, 0]), ("Critical", lambda c: c[:, day, 5]),
                          ("deaths", lambda c: c[:, day, 7])):
            xa, xb = col(a).astype(float), col(b).astype(float)
            se = np.sqrt(xa.var(ddof=1) / len(xa) + xb.var(ddof=1) / len(xb))
            assert abs(xa.mean() - xb.mean()) <= CI_SIGMAS * se + 1.0, \
                "%s: day %d %s means %.1f vs %.1f (se %.2f)" % (label, day, name, xa.mean(), xb.mean(), se)


def _oracle_ensemble(c, mode, seeds):
    T = int(c.params["T"])
    curves = np.zeros((len(seeds), T, 9), dtype=np.int64)
    for k, s in enumerate(seeds):
        tup, _, _, _ = c.keyed_oracle(mode, c.params, seed=s)
        for j in range(9):
            curves[k, :, j] = np.asarray(tup[j]).sum(axis=1)
    return curves


@pytest.mark.parametrize("which,mode", [("no_tracing", "native"), ("no_tracing", "replay"), ("tracing", "native")])
def test_keyed_oracle_ensemble_matches_reference_ensemble(which, mode):
    c, ref = reference_ensemble(which)
    got = _oracle_ensemble(c, mode, list(range(500, 500 + N_SEEDS)))
    assert_same_distribution(got, ref, c.n, "oracle/%s/%s" % (which, mode))


@pytest.mark.parametrize("which", ["no_tracing", "tracing"])
def test_reference_ensemble_fixture_is_sane(which):
    c, ref = reference_ensemble(which)
    assert ref.shape == (N_SEEDS, int(c.params["T"]), 9)
    assert (ref[:, :, [0, 1, 2, 4, 5, 6, 7]].sum(axis=2) == c.n).all()  # compartments partition the agents
    assert (np.diff(ref[:, :, 0], axis=1) <= 0).all() and (np.diff(ref[:, :, 7], axis=1) >= 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("which,mode", [("no_tracing", "native"), ("no_tracing", "replay"), ("tracing", "native"),
                                        ("tracing", "replay")])
def test_gpu_ensemble_matches_reference_ensemble(gpu_engine_module, which, mode):
    """64 seeds as 64 replicas of one persistent launch on the device (with the `tracing` fixture: 64 concurrent
    parallel tracers)."""
    c, ref = reference_ensemble(which)
    eng = c.gpu_engine(gpu_engine_module, c.params, mode, n_replicas=N_SEEDS)
    seeds = list(range(900, 900 + N_SEEDS))
    res = ensemble.run_seeds(eng, seeds, c.age_groups, c.fraction_stay_home, c.params["initial_infected_fraction"])
    got = ensemble.curves_from_tallies(res["tallies"])
    assert_same_distribution(got, ref, c.n, "gpu/%s/%s" % (which, mode))
    # and the device ensemble is the keyed oracle's, bit for bit, on a few of its members
    for k in (0, 17, 63):
        tup, _, _, _ = c.keyed_oracle(mode, c.params, seed=seeds[k])
        want = np.stack([np.asarray(tup[j]).sum(axis=1) for j in range(9)], axis=1)
        assert np.array_equal(got[k], want)
    eng.close()


# --------------------------------------------------------------------------------------------------------------
# The tightened tier (round 2): 256 reference seeds against 256 keyed seeds on the same population, KS p > 0.01,
# plus what the daily curves cannot show -- WHO gets infected (final attack counts per age decade) and HOW the
# infections are distributed over the infectors (the offspring numbers num_infected_by, incl. the double counting of
# same-day multiple hits).
N_LARGE = 256
KS_P_MIN_LARGE = 0.01


def reference_ensemble_large():
    c = util.Case(FIXTURES["no_tracing"])
    fx = np.load(util.os.path.join(util.GOLDEN, "ref_ensemble256_italy_n20000.npz"))
    return c, fx


def assert_same_distribution_large(curves, attack_by_age, offspring, fx, n, label):
    ref = fx["curves"].astype(np.int64)
    stats = {
        "final deaths": lambda c: c[:, -1, 7], "peak ICU": lambda c: c[:, :, 5].max(axis=1),
        "final attack size": lambda c: n - c[:, -1, 0], "peak Mild": lambda c: c[:, :, 2].max(axis=1),
        "final Documented": lambda c: c[:, -1, 3], "final Q": lambda c: c[:, -1, 8],
        "day of peak Mild": lambda c: c[:, :, 2].argmax(axis=1), "infections by day 30": lambda c: n - c[:, 30, 0],
    }
    for name, f in stats.items():
        p = st.ks_2samp(f(curves), f(ref)).pvalue
        assert p > KS_P_MIN_LARGE, "%s: KS p=%.3g on %s (means %.1f vs %.1f)" % (label, p, name, f(curves).mean(), f(ref).mean())

    def close(xa, xb, what):
        se = np.sqrt(xa.var(ddof=1) / len(xa) + xb.var(ddof=1) / len(xb))
        assert abs(xa.mean() - xb.mean()) <= CI_SIGMAS * se + 0.5, "%s: %s means %.2f vs %.2f (se %.3f)" % (label, what, xa.mean(), xb.mean(), se)
    dec = np.minimum(np.arange(101) // 10, 9)
    for d in range(10):   # final attack counts per age decade
        close(attack_by_age[:, dec == d].sum(axis=1).astype(float), fx["attack_by_age"][:, dec == d].sum(axis=1).astype(float),
              "attack count, ages %d-%d" % (10 * d, 10 * d + 9))
    fa = offspring / offspring.sum(axis=1, keepdims=True)   # share of infected agents that infected k others, per seed
    fb = fx["offspring"] / fx["offspring"].sum(axis=1, keepdims=True)
    for k in range(8):
        close(fa[:, k] * 1000, fb[:, k] * 1000, "per mille of infectors with %d offspring" % k)
    mean_a = (offspring * np.arange(16)).sum(axis=1) / offspring.sum(axis=1)
    mean_b = (fx["offspring"] * np.arange(16)).sum(axis=1) / fx["offspring"].sum(axis=1)
    assert st.ks_2samp(mean_a, mean_b).pvalue > KS_P_MIN_LARGE, "%s: mean offspring number" % label


def _summaries(tup, age):
    te, nib = tup[15], tup[9]
    return (np.stack([np.asarray(tup[j]).sum(axis=1) for j in range(9)], axis=1), np.bincount(age[te >= 0], minlength=101),
            np.bincount(np.minimum(nib[nib >= 0].astype(np.int64), 15), minlength=16))


def test_keyed_oracle_large_ensemble_matches_reference():
    c, fx = reference_ensemble_large()
    outs = [_summaries(c.keyed_oracle("native", c.params, seed=20000 + s)[0], c.age) for s in range(N_LARGE)]
    assert_same_distribution_large(np.stack([o[0] for o in outs]), np.stack([o[1] for o in outs]),
                                   np.stack([o[2] for o in outs]), fx, c.n, "oracle/native/256")


@pytest.mark.gpu
def test_gpu_large_ensemble_matches_reference(gpu_engine_module):
    """256 seeds as four launches of 64 replicas, prologue picks made on the device (keyed permutation)."""
    c, fx = reference_ensemble_large()
    R = 64
    eng = c.gpu_engine(gpu_engine_module, c.params, "native", n_replicas=R)
    curves, attack, off = [], [], []
    for b in range(N_LARGE // R):
        seeds = [30000 + b * R + r for r in range(R)]
        eng.run(seeds, keyed=(c.fraction_stay_home, c.params["initial_infected_fraction"]))
        for r in range(R):
            curves.append(eng.tallies(r).sum(axis=2))
            te, nib = eng.agent_vector("time_exposed", replica=r), eng.agent_vector("num_infected_by", replica=r)
            attack.append(np.bincount(c.age[te >= 0], minlength=101))
            off.append(np.bincount(np.minimum(nib[nib >= 0].astype(np.int64), 15), minlength=16))
    eng.close()
    assert_same_distribution_large(np.stack(curves), np.stack(attack), np.stack(off), fx, c.n, "gpu/native/256")

