"""The per-job result files of `parameter_sweep.py` (:617-733), written from the device tallies.

One job (SLURM array index) = N_SIMS_PER_JOB runs of one parameter combination; the reference writes 19
tab-separated files without header (`na_rep='NA'`), one row per run, named

    {sim_name}_paramsweep_n{params['n']}_i{combo index}_N{trial}_p{p_infect_given_contact}_m{mortality multiplier}_s{start day}_{datatype}.csv

`combine_results.py`, `parse_parameter_sweep.py` and the Policy_Section scripts read exactly these, so a sweep run
through `sweep.run_sweep` can feed the reference's own post-processing.  Shapes: r0_tot, mse [N]; thirteen
[N, T] series; four [N, 5 * T] age-group series flattened group-major like the reference's reshape (:727).
"""
import os

import numpy as np

from . import driver_tallies as _dt

N_T_SERIES = (("susceptible", "S"), ("exposed", "E"), ("deaths", "D"), ("mild", "Mild"), ("severe", "Severe"),
              ("critical", "Critical"), ("recovered", "R"), ("quarantine", "Q"))
COHORT_SERIES = (("r0_time", "r_0_over_time"), ("cfr_time", "cfr_over_time"), ("frac_over_70_time", "fraction_over_70_time"),
                 ("frac_below_30_time", "fraction_below_30_time"), ("median_age_time", "median_age_time"))
AGE_SERIES = (("infected_age_time", "infected_by_age_by_time"), ("total_age_time", "total_age_by_time"),
              ("cfr_age_time", "CFR_by_age_by_time"), ("dead_age_time", "dead_by_age_by_time"))
DATATYPES = ("r0_tot", "mse") + tuple(k for k, _ in N_T_SERIES + COHORT_SERIES + AGE_SERIES)


def file_stem(sim_name, n, combo_index, trial, pigc, mm, sd):
    """parameter_sweep.py:618 (params['n'] is a float there: '%s' gives e.g. 10000000.0)."""
    return "%s_paramsweep_n%s_i%s_N%s_p%s_m%s_s%s" % (sim_name, float(n), combo_index, trial, pigc, mm, sd)


def job_arrays(tallies, cohorts, actual_deaths=None, first_day=None):
    """{datatype: array} for one job from tallies int64[N, T, 9, 101] and cohorts int64[N, T, 4, 101].
    `mse` needs the observed cumulative deaths for days [first_day, first_day + len(actual_deaths)) (:611-614)."""
    tallies, cohorts = np.asarray(tallies), np.asarray(cohorts)
    N, T = tallies.shape[0], tallies.shape[1]
    out = {"r0_tot": np.array([_dt.r0_total(cohorts[i]) for i in range(N)])}
    per = [_dt.per_time(tallies[i]) for i in range(N)]
    if actual_deaths is not None:
        obs = np.asarray(actual_deaths, dtype=np.float64)
        out["mse"] = np.array([np.mean((per[i]["D"][first_day:first_day + len(obs)] - obs) ** 2) for i in range(N)])
    else:
        out["mse"] = np.full(N, np.nan)
    for name, plane in N_T_SERIES:
        out[name] = np.stack([per[i][plane] for i in range(N)]).astype(np.float64)
    coh = [_dt.cohort_statistics(cohorts[i]) for i in range(N)]
    for name, key in COHORT_SERIES:
        out[name] = np.stack([coh[i][key] for i in range(N)])
    ages = [_dt.by_age_group(tallies[i], cohorts[i]) for i in range(N)]
    for name, key in AGE_SERIES:
        out[name] = np.stack([ages[i][key].reshape(-1) for i in range(N)])  # [N, 5 * T], group-major (:727)
    return out


AGEPOLICY_DATATYPES = ("r0_tot", "mse", "infected", "susceptible", "exposed", "deaths", "mild", "severe", "critical",
                       "recovered", "quarantine", "documented")
"""The twelve per-job files of `simulate_agepolicies.py` (:645-716)."""


def agepolicy_file_stem(sim_name, n, combo_index, trial, pigc, mm, sd):
    """simulate_agepolicies.py:646."""
    return "%s_agepolicy_n%s_i%s_N%s_p%s_m%s_s%s" % (sim_name, float(n), combo_index, trial, pigc, mm, sd)


def agepolicy_job_arrays(tallies, cohorts, n, actual_deaths=None, first_day=None):
    """{datatype: array} of simulate_agepolicies.py from tallies int64[N, T, 9, 101] and cohorts int64[N, T, 4, 101];
    `infected` is n - S per day (:617), `documented` the Documented column sums."""
    base = job_arrays(tallies, cohorts, actual_deaths=actual_deaths, first_day=first_day)
    out = {k: base[k] for k in ("r0_tot", "mse", "susceptible", "exposed", "deaths", "mild", "severe", "critical",
                                "recovered", "quarantine")}
    t = np.asarray(tallies)
    out["infected"] = (float(n) - t[:, :, 0, :].sum(axis=2)).astype(np.float64)
    out["documented"] = t[:, :, 3, :].sum(axis=2).astype(np.float64)
    return out


def _format_tsv(a):
    """The bytes `pd.DataFrame(a).to_csv(sep="\t", index=False, header=False, na_rep="NA")` writes for a 1-D or 2-D numeric
    array (parameter_sweep.py:716-729): float64 as the shortest repr ("3.0", "0.1", "1e-05", "inf"), NaN as NA, integers
    as integers, one line per row.  pandas spends 0.4 ms per call before the first byte; a job has 19 files of a few
    hundred numbers, and a sweep thousands of jobs (tests/test_sweep.py compares the two byte for byte)."""
    a = np.asarray(a)
    if a.ndim == 1:
        a = a[:, None]
    if a.dtype.kind in "iub":
        rows = ["\t".join(map(str, r)) for r in a.tolist()]
        return ("\n".join(rows) + "\n") if rows else ""
    # floats: the library's host routine (seir_format_tsv) -- formatted value by value in Python this took longer than
    # the device run of the job; _format_tsv_py below is the same in Python (tests compare the two and pandas)
    from . import engine as _engine
    return _engine.format_tsv(a).decode("ascii")


def _format_tsv_bytes(a):
    """_format_tsv as bytes (what goes into the file)."""
    a = np.asarray(a)
    if a.dtype.kind in "iub":
        return _format_tsv(a).encode("ascii")
    from . import engine as _engine
    return _engine.format_tsv(a)


def _format_tsv_py(a):
    """_format_tsv for floats in plain Python (the checked alternative of seir_format_tsv)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    rows = ["\t".join("NA" if x != x else repr(x) for x in r) for r in a.tolist()]
    return ("\n".join(rows) + "\n") if rows else ""


def write_job_files(directory, stem, arrays, datatypes=None):
    """One file per datatype, the reference's to_csv settings (:716-729).  Returns the paths."""
    os.makedirs(directory, exist_ok=True)
    paths = {}
    for name in (DATATYPES if datatypes is None else datatypes):
        path = os.path.join(directory, "%s_%s.csv" % (stem, name))
        data = _format_tsv_bytes(arrays[name])
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)  # (no buffered text layer: one write per file)
        try:
            view = memoryview(data)
            while len(view):
                view = view[os.write(fd, view):]
        finally:
            os.close(fd)
        paths[name] = path
    return paths
