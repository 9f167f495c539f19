import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from covid19_demography_b200 import engine, population, tables
n = 10_000_000
p, tabs = bench.workload_params(n, 60); p["t_tracing_start"] = 37.0
hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p), tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"], tabs["p_critical_death"], with_replay=False)
eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1)
rng = np.random.default_rng(100)
for k in range(3):
    eng.run([k * 64], None, [rng.choice(n, size=5, replace=False).astype(np.int64)])
us = eng.day_times_us(); st = eng.stage_times_us()
print("loop ms", eng.last_run_ms3())
for t in range(36, 60):
    print("day %2d total %8.1f us | pass end %.1f | roots+barrier %.1f  bfs %.1f  reps %.1f  sizes %.1f  warps %.1f  ctas+barrier %.1f" % (
        t, us[t-1], st[t,7], st[t,1]-st[t,7], st[t,2]-st[t,1], st[t,3]-st[t,2], st[t,4]-st[t,3], st[t,5]-st[t,4], st[t,6]-st[t,5]))
