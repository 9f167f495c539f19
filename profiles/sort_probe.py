"""Experiment: what would ID-ordered active lists be worth?  (debug build SEIR_NVCC_EXTRA=-DSEIR_SORT_PROBE, per-day
launches; with SEIR_DEBUG_SORT=<mask> the host sorts tomorrow's list segments by agent id after every day kernel;
the figure is the sum of the day kernels' own durations.)   python profiles/sort_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from covid19_demography_b200 import engine, population, tables

n = 10_000_000
hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
for label, ninit, lockdown in (("C2", 5, None), ("stress (0.1 % infected at t = 0, no lockdown)", 10000, 1.0e6)):
    p, tabs = bench.workload_params(n, 60)
    if lockdown:
        p["t_lockdown"] = lockdown
    tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                             tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                             tabs["p_critical_death"], with_replay=False)
    for sort in (0, 7, 2, 5):  # bit 0: agents that may transmit, bit 1: new infectees, bit 2: quarantined agents
        os.environ["SEIR_DEBUG_SORT"] = str(sort)
        eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1,
                            launch="per_day")
        rng = np.random.default_rng(100)
        ms = []
        for k in range(3):
            init = [rng.choice(n, size=ninit, replace=False).astype(np.int64)]
            eng.run([k * 64], None, init)
            ms.append(eng.last_run_ms()[0])
        print("%-48s lists %s: day kernels %.3f ms (runs: %s), infections %d"
              % (label, {0: "in arrival order", 7: "all sorted by id", 2: "only the new infectees sorted", 5: "only the survivors sorted"}[sort], ms[-1], " ".join("%.3f" % x for x in ms),
                 eng.counters()["infections"]), flush=True)
        eng.close()
