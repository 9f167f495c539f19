"""The per-run statistics the reference's drivers compute from run_model's outputs, derived from the
device-side tallies instead of O(T * n) NumPy passes over the nine bool[T, n] histories.

    tallies  int64[T, 9, 101]   per-day per-age compartment counts   (engine.Engine.tallies)
    cohorts  int64[T, 4, 101]   per exposure day and age: exposed, sum num_infected_by, dead / recovered at the end
                                                                      (engine.Engine.cohorts)

Reference statements restated here: run_simulation.py:448-460, simulate_agepolicies.py:617-631,
parameter_sweep.py:556-603.  Divisions by zero give nan exactly where NumPy gives nan/inf warnings
in the reference (an empty cohort keeps the drivers' initial value, which the caller supplies).
"""
import numpy as np

PLANES = ("S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q")
AGE_GROUPS_SWEEP = ((0, 14), (15, 24), (25, 39), (40, 69), (70, 100))  # parameter_sweep.py:538


def per_time(tallies):
    """{plane: X.sum(axis=1)} -- S_per_time ... Q_per_time (parameter_sweep.py:556-565)."""
    s = np.asarray(tallies).sum(axis=2)   # [T, 9]: one reduction for the nine planes
    return {k: s[:, j] for j, k in enumerate(PLANES)}


def _median_from_histogram(counts):
    """np.median of the multiset {age a repeated counts[a] times}."""
    c = np.asarray(counts, dtype=np.int64)
    m = int(c.sum())
    if m == 0:
        return np.nan
    cum = np.cumsum(c)
    lo = int(np.searchsorted(cum, (m - 1) // 2 + 1))
    hi = int(np.searchsorted(cum, m // 2 + 1))
    return 0.5 * (lo + hi)


def _medians_from_histograms(counts):
    """_median_from_histogram for every row of counts[T, 101] at once (rows without agents: nan)."""
    c = np.asarray(counts, dtype=np.int64)
    m = c.sum(axis=1)
    cum = np.cumsum(c, axis=1)
    # searchsorted(cum_row, k) = number of entries of the row below k
    lo = (cum < ((m - 1) // 2 + 1)[:, None]).sum(axis=1)
    hi = (cum < (m // 2 + 1)[:, None]).sum(axis=1)
    return np.where(m > 0, 0.5 * (lo + hi), np.nan)


def cohort_statistics(cohorts, fill=0.0):
    """run_simulation.py:448-453: per exposure day t (rows with no exposure keep `fill`):
    r_0_over_time, cfr_over_time, fraction_over_70_time, fraction_below_30_time, median_age_time."""
    ci = np.asarray(cohorts)
    tot = ci.sum(axis=2).astype(np.float64)   # [T, 4] (counts: exact in float64 whatever the order of the sums)
    m, d, r = tot[:, 0], tot[:, 2], tot[:, 3]
    has = m > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        vals = {"r_0_over_time": tot[:, 1] / m,
                "cfr_over_time": np.where(d + r > 0, d / (d + r), np.nan),
                "fraction_over_70_time": ci[:, 0, 70:].sum(axis=1).astype(np.float64) / m,
                "fraction_below_30_time": ci[:, 0, :30].sum(axis=1).astype(np.float64) / m,
                "median_age_time": _medians_from_histograms(ci[:, 0])}
    return {k: np.where(has, v, fill).astype(np.float64) for k, v in vals.items()}


def r0_total(cohorts, first=1, last=20):
    """parameter_sweep.py:567: mean num_infected_by of the agents with 0 < time_exposed <= 20."""
    c = np.asarray(cohorts)
    m = c[first:last + 1, 0].sum()
    return c[first:last + 1, 1].sum() / m if m > 0 else np.nan


def totals(tallies, n):
    """run_simulation.py:454-459: total_infections_time = n - S - E, total_documented_time, total_deaths_time,
    dead_by_age (D[-1] by age)."""
    t = np.asarray(tallies)
    S, E = t[:, 0].sum(axis=1), t[:, 1].sum(axis=1)
    return {"total_infections_time": n - S - E, "total_documented_time": t[:, 3].sum(axis=1),
            "total_deaths_time": t[:, 7].sum(axis=1), "dead_by_age": t[-1, 7].copy()}


def by_age_group(tallies, cohorts, groups=AGE_GROUPS_SWEEP):
    """parameter_sweep.py:582-600: per age group and day -- infected (group size - S[t]), group size,
    CFR of the day's exposure cohort, dead (D[t])."""
    t = np.asarray(tallies)
    T = t.shape[0]
    G = len(groups)
    size_by_age = t[0, [0, 1, 2, 4, 5, 6, 7]].sum(axis=0)  # the compartments partition the agents
    if all(groups[g][1] + 1 == groups[g + 1][0] for g in range(G - 1)) and groups[-1][1] == t.shape[2] - 1:
        # consecutive age bands up to the last age (the sweep's): one segmented reduction per quantity
        starts = [lo for lo, _ in groups]
        size_g = np.add.reduceat(size_by_age, starts).astype(np.float64)
        dr = np.add.reduceat(np.asarray(cohorts)[:, 2:4, :], starts, axis=2).astype(np.float64)   # [T, 2, G]
        with np.errstate(invalid="ignore", divide="ignore"):
            cfr = dr[:, 0, :] / (dr[:, 0, :] + dr[:, 1, :])
        return {"infected_by_age_by_time": size_g[:, None] - np.add.reduceat(t[:, 0, :], starts, axis=1).T,
                "total_age_by_time": np.repeat(size_g[:, None], T, axis=1),
                "CFR_by_age_by_time": np.ascontiguousarray(cfr.T),
                "dead_by_age_by_time": np.add.reduceat(t[:, 7, :], starts, axis=1).T.astype(np.float64)}
    c = np.asarray(cohorts).astype(np.float64)
    out = {"infected_by_age_by_time": np.zeros((G, T)), "total_age_by_time": np.zeros((G, T)),
           "CFR_by_age_by_time": np.full((G, T), np.nan), "dead_by_age_by_time": np.zeros((G, T))}
    for g, (lo, hi) in enumerate(groups):
        sl = slice(lo, hi + 1)
        out["total_age_by_time"][g, :] = size_by_age[sl].sum()
        out["infected_by_age_by_time"][g, :] = size_by_age[sl].sum() - t[:, 0, sl].sum(axis=1)
        out["dead_by_age_by_time"][g, :] = t[:, 7, sl].sum(axis=1)
        d, r = c[:, 2, sl].sum(axis=1), c[:, 3, sl].sum(axis=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            out["CFR_by_age_by_time"][g, :] = d / (d + r)
    return out
