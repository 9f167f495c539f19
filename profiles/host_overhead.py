"""Where the end-to-end time of a C2 run goes on the host side: wall time of Engine.run / Engine.tallies against the
device time of the same run.   python profiles/host_overhead.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from covid19_demography_b200 import engine, population, tables

n = 10_000_000
p, tabs = bench.workload_params(n, 60)
hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                         tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                         tabs["p_critical_death"], with_replay=False)
eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1)
rng = np.random.default_rng(100)
inits = [engine.pin(rng.choice(n, size=5, replace=False).astype(np.int64)) for _ in range(24)]
rows = []
for k, init in enumerate(inits):
    t0 = time.perf_counter()
    eng.run([k], None, [init])
    t1 = time.perf_counter()
    eng.tallies(0)
    t2 = time.perf_counter()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, eng.last_run_ms()[1]))
r = np.array(rows[4:])
print("per run (ms), median of %d: Engine.run wall %.3f | device (events) %.3f | run overhead %.3f | Engine.tallies %.3f"
      % (len(r), np.median(r[:, 0]), np.median(r[:, 2]), np.median(r[:, 0] - r[:, 2]), np.median(r[:, 1])))
