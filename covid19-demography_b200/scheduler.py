"""A sweep scheduler: one worker per GPU pulling jobs from a shared queue.

The reference spreads a sweep as a SLURM job array (`job.parameter_sweep.sh:15`, one single-threaded process per
array index, `parameter_sweep.py:107-126`) and repairs failures by resubmitting by hand (`combine_results.py:91-92`
silently skips missing files).  Here:

  * a JOB is the reference's array index: N_SIMS_PER_JOB seeds of one parameter combination, its result the
    reference's 19 (parameter sweep) or 12 (age policies) tab-separated files (`sweep_io`);
  * the QUEUE is the output directory itself: a job whose files all exist is done (resumable by construction, the
    same idempotence the file names give the reference), a job somebody is working on has a claim file created with
    O_EXCL (atomic on a local or NFS file system); a claim whose process is gone is taken over;
  * a WORKER (one process per GPU, nothing shared but the directory) pulls the next unclaimed jobs until its
    engine's replica slots are full -- the jobs of one launch may belong to DIFFERENT combinations: their day numbers
    and table sets travel as per-replica overrides (`Engine.run(overrides=...)`, `Engine.set_table_bank`) -- runs
    them in ONE persistent launch, writes the files (on a writer thread, overlapping the next launch) and goes on.

Seeds follow the reference's arithmetic (`sweep.job_seeds`), so results stay comparable job for job.
"""
import os
import socket
import threading
import time

import numpy as np

from . import sweep as _sweep
from . import sweep_io as _io
from . import tables as _tables


class FileQueue:
    """Claims on jobs through files in `directory` (shared by all workers of a sweep)."""

    def __init__(self, directory, stale_after_s=6 * 3600):
        self.dir = directory
        self.stale_after_s = stale_after_s
        os.makedirs(directory, exist_ok=True)
        self.me = "%s:%d" % (socket.gethostname(), os.getpid())

    def _claim_path(self, stem):
        return os.path.join(self.dir, stem + ".claim")

    def done(self, stem, datatypes):
        return all(os.path.exists(os.path.join(self.dir, "%s_%s.csv" % (stem, d))) for d in datatypes)

    def _owner_alive(self, path):
        try:
            text = open(path).read().strip()
            if ":" not in text:   # created a moment ago, its owner has not written its name yet (O_EXCL create, then write)
                return time.time() - os.path.getmtime(path) < 10.0
            host, pid = text.rsplit(":", 1)
            if host != socket.gethostname():
                return time.time() - os.path.getmtime(path) < self.stale_after_s
            os.kill(int(pid), 0)
            return True
        except (OSError, ValueError):
            return False

    def claim(self, stem):
        """True if this process now owns the job."""
        path = self._claim_path(stem)
        for _ in range(2):
            try:
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
                os.write(fd, self.me.encode())
                os.close(fd)
                return True
            except FileExistsError:
                if self._owner_alive(path):
                    return False
                try:   # the owner died (a killed sweep): take the job over
                    os.remove(path)
                except FileNotFoundError:
                    pass
        return False

    def release(self, stem):
        try:
            os.remove(self._claim_path(stem))
        except FileNotFoundError:
            pass


class SweepPlan:
    """The jobs of a sweep and what each combination means for the device.

    combos[k] = {"params": the 24 run_model keys, "tables": dict for seir_tables (tables.build_tables),
                 "label": (p_infect_given_contact, mortality multiplier, start day) for the file names}
    kind = "paramsweep" (19 files, parameter_sweep.py:617-733) or "agepolicy" (12 files, simulate_agepolicies.py:645-716).
    """

    def __init__(self, sim_name, combos, n_sims_per_combo, n_sims_per_job, seed_offset=0, kind="paramsweep",
                 fraction_stay_home=None, actual_deaths=None, first_day=None):
        self.sim_name, self.combos, self.kind = sim_name, combos, kind
        self.items = _sweep.work_items(len(combos), n_sims_per_combo, n_sims_per_job, seed_offset)
        self.datatypes = _io.DATATYPES if kind == "paramsweep" else _io.AGEPOLICY_DATATYPES
        self.fraction_stay_home = np.zeros(_tables.N_AGES) if fraction_stay_home is None else np.asarray(fraction_stay_home, dtype=np.float64)
        self.actual_deaths, self.first_day = actual_deaths, first_day
        self.T_max = max(int(_tables.params_to_dict(c["params"])["T"]) for c in combos)

    def stem(self, job):
        c = self.combos[job["combo"]]
        n = _tables.params_to_dict(c["params"])["n"]
        pigc, mm, sd = c["label"]
        f = _io.file_stem if self.kind == "paramsweep" else _io.agepolicy_file_stem
        return f(self.sim_name, n, job["combo"], job["trial"], pigc, mm, sd)

    def overrides(self, combo):
        p = _tables.params_to_dict(self.combos[combo]["params"])
        return {"T": int(p["T"]), "t_lockdown": int(p["t_lockdown"]), "t_stayhome_start": int(min(p["t_stayhome_start"], 2**31 - 1)),
                "t_tracing_start": int(min(p["t_tracing_start"], 2**31 - 1)), "table_set": combo}


def _write(plan, directory, job, tallies, cohorts, n):
    if plan.kind == "paramsweep":
        arrays = _io.job_arrays(tallies, cohorts, actual_deaths=plan.actual_deaths, first_day=plan.first_day)
    else:
        arrays = _io.agepolicy_job_arrays(tallies, cohorts, n, actual_deaths=plan.actual_deaths, first_day=plan.first_day)
    tmp = os.path.join(directory, ".tmp_%d_%d" % (os.getpid(), job["index"]))
    paths = _io.write_job_files(tmp, plan.stem(job), arrays, datatypes=plan.datatypes)
    for p in paths.values():   # a job is done when ALL its files are in place: move them in last
        os.replace(p, os.path.join(directory, os.path.basename(p)))
    os.rmdir(tmp)


def run_worker(engine, plan, directory, prologue="keyed", age_groups=None, max_jobs=None, log=None):
    """Pull and run jobs until the queue is empty.  `engine` must have been created with T >= plan.T_max and
    `engine.set_table_bank([c["tables"] for c in plan.combos], ...)` is done here.  prologue = "keyed": the
    device draws Home_real / initial_infected itself; "reference": tables.prologue replays the reference's
    Mersenne-Twister picks on the host (needs age_groups).  Returns the number of jobs this worker ran."""
    q = FileQueue(directory)
    engine.set_table_bank([c["tables"] for c in plan.combos],
                          [_tables.params_to_dict(c["params"])["p_infect_given_contact"] for c in plan.combos])
    R = engine.n_replicas
    per_job = len(plan.items[0]["seeds"])
    if per_job > R:
        raise ValueError("a job's %d seeds do not fit the engine's %d replica slots" % (per_job, R))
    jobs_per_launch = R // per_job
    writer, ran, cursor = None, 0, 0
    errors = []
    while cursor < len(plan.items) and (max_jobs is None or ran < max_jobs):
        batch = []
        while cursor < len(plan.items) and len(batch) < jobs_per_launch and (max_jobs is None or ran + len(batch) < max_jobs):
            job = plan.items[cursor]
            cursor += 1
            stem = plan.stem(job)
            if q.done(stem, plan.datatypes) or not q.claim(stem):
                continue
            if q.done(stem, plan.datatypes):   # (its owner finished and released it between the two tests above)
                q.release(stem)
                continue
            batch.append(job)
        if not batch:
            continue
        seeds, ov, fracs = [], [], []
        for job in batch:
            p = _tables.params_to_dict(plan.combos[job["combo"]]["params"])
            for s in job["seeds"]:
                seeds.append(s)
                ov.append(plan.overrides(job["combo"]))
                fracs.append(p["initial_infected_fraction"])
        used = len(seeds)
        while len(seeds) < R:   # (idle slots repeat the last run; their results are dropped)
            seeds.append(seeds[-1]); ov.append(ov[-1]); fracs.append(fracs[-1])
        if prologue == "keyed":
            engine.run(seeds, keyed=[(plan.fraction_stay_home, f) for f in fracs], overrides=ov)
        else:
            picks = [_tables.prologue(s, age_groups, engine.n, plan.fraction_stay_home, f) for s, f in zip(seeds, fracs)]
            engine.run(seeds, [h for h, _ in picks], [i for _, i in picks], overrides=ov)
        results = []
        for j, job in enumerate(batch):
            rs = range(j * per_job, (j + 1) * per_job)
            T = ov[j * per_job]["T"]
            results.append((job, np.stack([engine.tallies(r)[:T] for r in rs]), np.stack([engine.cohorts(r)[:T] for r in rs])))
        assert used == len(batch) * per_job
        if writer is not None:
            writer.join()

        def flush(results=results):
            for job, tal, coh in results:
                try:
                    _write(plan, directory, job, tal, coh, engine.n)
                except Exception as e:   # noqa: BLE001  (reported by the worker below)
                    errors.append((job["index"], e))
                finally:
                    q.release(plan.stem(job))
        writer = threading.Thread(target=flush)
        writer.start()   # the files are written while the next launch runs
        ran += len(batch)
        if log is not None:
            log("launch of %d job(s), %d run(s): %.2f ms on the device" % (len(batch), used, sum(engine.last_run_ms3())))
    if writer is not None:
        writer.join()
    if errors:
        raise RuntimeError("writing the results of job %d failed: %r" % errors[0])
    return ran


def missing_jobs(plan, directory):
    """Indices of the jobs whose files are not all there (what a restarted sweep still has to do)."""
    q = FileQueue(directory)
    return [job["index"] for job in plan.items if not q.done(plan.stem(job), plan.datatypes)]
