"""Wall time of a FIXED sweep workload on N GPUs of one box through the scheduler (covid19-demography_b200/scheduler.py):
one worker process per GPU pulling jobs from the shared output directory, jobs of different grid points batched into
one launch, the reference's per-job files written.  The efficiency the driver cannot see from a per-step maximum.

    python profiles/sweep_scaling.py --gpus 4 [--workload c5|c2] [--resume-test]

  c5: parameter_sweep.py's grid, 360 combinations (12 transmissibilities x 5 mortality multipliers x 6 start days,
      :54-56) x 2 jobs x 2 seeds = 1440 runs at 1 M agents, 64 replica slots per GPU, 19 files per job;
  c2: 512 seeds of the C2 workload (10 M agents, T = 60) as 128 jobs of 4 seeds, 8 replica slots per GPU.
Prints one JSON line: wall seconds, runs per hour, and (with --resume-test) the wall time of a sweep that was killed
half way and restarted.
"""
import argparse
import json
import multiprocessing as mp
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def make_plan(workload):
    import bench
    from covid19_demography_b200 import scheduler, tables
    if workload == "c5":
        n, R = 1_000_000, 64
        pop = bench.make_population(n)
        gs = np.array([len(g) for g in pop[4]])
        combos = []
        p0, tabs = bench.workload_params(n, 73)
        p0["initial_infected_fraction"] = 5.0 / n
        base = None
        for pigc in np.linspace(0.015, 0.048, 12):
            for mm in (1.0, 2.0, 3.0, 4.0, 5.0):
                sev = [np.clip(tabs[x] * (mm / 4.0) ** (1.0 / 3.0), 0.0, 1.0) for x in ("p_mild_severe", "p_severe_critical", "p_critical_death")]
                for d0 in (10, 13, 16, 19, 22, 25):   # start day: moves T and the lockdown day (parameter_sweep.py:134-175)
                    p = dict(p0, p_infect_given_contact=float(pigc), T=float(48 + d0), t_lockdown=float(21 + d0))
                    tb = tables.build_tables(tabs["contact_matrix"], gs, tables.params_to_dict(p), tabs["mean_time_to_isolate_factor"],
                                             *sev, with_replay=False, alias_from=base)
                    base = base or tb
                    combos.append({"params": p, "tables": tb, "label": (round(float(pigc), 4), mm, d0)})
        plan = scheduler.SweepPlan("c5", combos, 4, 2, seed_offset=0)
        return plan, pop, tabs, R, n
    n, R = 10_000_000, 8
    pop = bench.make_population(n)
    p, tabs = bench.workload_params(n, 60)
    tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in pop[4]]), tables.params_to_dict(p),
                             tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                             tabs["p_critical_death"], with_replay=False)
    plan = scheduler.SweepPlan("c2", [{"params": p, "tables": tb, "label": (0.029, 4, 22)}], 512, 4, seed_offset=0)
    return plan, pop, tabs, R, n


def worker(rank, workload, directory, max_jobs, ready, go, out):
    from covid19_demography_b200 import engine, scheduler, tables
    plan, pop, tabs, R, n = make_plan(workload)
    hh, age, diab, hyp, groups = pop
    base = dict(plan.combos[0]["params"], T=float(plan.T_max))
    eng = engine.Engine(tables.params_to_dict(base), hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]),
                        plan.combos[0]["tables"], n_replicas=R, device=rank, history="on_demand")
    ready.put(rank)
    go.wait()
    t0 = time.perf_counter()
    ran = scheduler.run_worker(eng, plan, directory, max_jobs=max_jobs)
    out.put((rank, ran, time.perf_counter() - t0))
    eng.close()


def run(workload, gpus, directory, max_jobs=None):
    ctx = mp.get_context("spawn")
    ready, out, go = ctx.Queue(), ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=worker, args=(r, workload, directory, max_jobs, ready, go, out)) for r in range(gpus)]
    for p in procs:
        p.start()
    for _ in procs:
        ready.get(timeout=1800)   # populations built, engines created: the clock starts when every worker is ready
    t0 = time.perf_counter()
    go.set()
    res = [out.get(timeout=3600) for _ in procs]
    wall = time.perf_counter() - t0
    for p in procs:
        p.join(timeout=120)
    return wall, sorted(res)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workload", default="c5", choices=["c5", "c2"])
    ap.add_argument("--resume-test", action="store_true")
    args = ap.parse_args()
    d = tempfile.mkdtemp(prefix="sweep_%s_" % args.workload)
    runs_total = {"c5": 1440, "c2": 512}[args.workload]
    line = {"workload": args.workload, "gpus": args.gpus, "runs": runs_total}
    if args.resume_test:
        jobs = {"c5": 720, "c2": 128}[args.workload]
        w1, r1 = run(args.workload, args.gpus, d, max_jobs=jobs // (2 * args.gpus))
        w2, r2 = run(args.workload, args.gpus, d)
        line.update({"first_half_wall_s": w1, "first_half_jobs": sum(r[1] for r in r1), "restart_wall_s": w2,
                     "restart_jobs": sum(r[1] for r in r2), "files": len([f for f in os.listdir(d) if f.endswith(".csv")])})
    else:
        wall, res = run(args.workload, args.gpus, d)
        line.update({"wall_s": wall, "runs_per_hour": 3600.0 * runs_total / wall, "jobs_per_worker": [r[1] for r in res],
                     "worker_wall_s": [round(r[2], 3) for r in res], "files": len([f for f in os.listdir(d) if f.endswith(".csv")])})
    shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(line))
