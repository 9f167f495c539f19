"""In-process parameter sweeps and seed ensembles over the GPUs of one box.

The reference spreads a sweep with a SLURM job array: job INDEX picks its parameter combination and
seeds by arithmetic (`parameter_sweep.py:107-126`, `simulate_agepolicies.py:165-175`):

    JOBS_PER_COMBO = N_SIMS_PER_COMBO // N_SIMS_PER_JOB
    combo = INDEX // JOBS_PER_COMBO ;  trial = INDEX % JOBS_PER_COMBO
    seed  = INDEX * N_SIMS_PER_JOB + SEED_OFFSET, then +1 before every run          (:112, :548)

`job_seeds` / `work_items` restate that arithmetic so results stay comparable run for run; `run_sweep`
deals the jobs round-robin to the ranks (one process per GPU, no communication), keeps ONE device-
resident population per rank, switches parameter sets with `Engine.set_params` and packs the seeds of
a job into replica batches.
"""
import numpy as np

from . import ensemble as _ensemble
from . import tables as _tables


def job_seeds(index, n_sims_per_job, seed_offset=0):
    """Seeds of the runs of SLURM job `index` (parameter_sweep.py:112, :546-548)."""
    base = index * n_sims_per_job + seed_offset
    return [base + k + 1 for k in range(n_sims_per_job)]


def work_items(n_combos, n_sims_per_combo, n_sims_per_job, seed_offset=0):
    """All jobs of a sweep in SLURM-array order: dicts {index, combo, trial, seeds}."""
    if n_sims_per_job < 1 or n_sims_per_combo % n_sims_per_job:
        raise ValueError("N_SIMS_PER_COMBO must be a multiple of N_SIMS_PER_JOB")
    jobs_per_combo = n_sims_per_combo // n_sims_per_job
    return [{"index": idx, "combo": idx // jobs_per_combo, "trial": idx % jobs_per_combo,
             "seeds": job_seeds(idx, n_sims_per_job, seed_offset)}
            for idx in range(n_combos * jobs_per_combo)]


def run_sweep(engine, combos, items, age_groups, fraction_stay_home, rank=0, world=1, on_job=None):
    """Run this rank's share of `items` on `engine`.

    combos[k] = {"params": mapping with the 24 run_model keys, "tables": dict for seir_tables} -- what the
    drivers rebuild per combination (transmissibility, mortality multiplier -> severity tables, start
    date -> T and the lockdown day).  Jobs are dealt round-robin over the ranks in combo-major order, so a
    rank changes parameter set only when its next job belongs to another combination.
    Returns {job index: {"tallies": int64[n_seeds, T, 9, 101], "cohorts": int64[n_seeds, T, 4, 101], "seeds": [...]}}.
    """
    mine = [items[i] for i in _ensemble.partition(len(items), rank, world)]
    out = {}
    current = None
    for job in mine:
        k = job["combo"]
        if k != current:
            engine.set_params(params=combos[k]["params"], tables=combos[k]["tables"])
            current = k
        p = _tables.params_to_dict(combos[k]["params"])
        res = _ensemble.run_seeds(engine, job["seeds"], age_groups, fraction_stay_home,
                                  p["initial_infected_fraction"], want_cohorts=True)
        out[job["index"]] = {"tallies": res["tallies"], "cohorts": res["cohorts"], "seeds": list(job["seeds"])}
        if on_job is not None:
            on_job(job, out[job["index"]])
    return out


def gather_sweep(local_results, n_items, dist=None):
    """All jobs' results on every rank, keyed by job index."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        pieces = [local_results]
    else:
        pieces = [None] * dist.get_world_size()
        dist.all_gather_object(pieces, local_results)
    merged = {}
    for piece in pieces:
        for idx, res in piece.items():
            if idx in merged:
                raise RuntimeError("job %d was run by two ranks" % idx)
            merged[idx] = res
    missing = [i for i in range(n_items) if i not in merged]
    if missing:
        raise RuntimeError("jobs %s were run by no rank" % missing[:8])
    return merged
