"""Per-day counters of the parallel tracer on C2 with tracing from day 37 (debug build: SEIR_NVCC_EXTRA=-DSEIR_TRACE_STATS):
roots, agents reached by the frontier step, whole-CTA components and the largest one, node expansions of the exact
searches (CTA / warp path), levels.   python profiles/trace_stats.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
    init = [rng.choice(n, size=5, replace=False).astype(np.int64)]
eng.run([2 * 64], None, init)   # (one run: the counters accumulate)
out = np.zeros((eng.T + 2) * 16, dtype=np.int64)
eng.lib.seir_get_fine_times(eng._h, out.ctypes.data)
f = out.reshape(-1, 16)
us = eng.day_times_us()
print("day   us     roots reached  bigcomps maxsize maxroots | cta: searches levels expansions | warp expansions | frontier levels")
for t in range(37, 60):
    r = f[t]
    print("%3d %7.0f %6d %7d %6d %7d %7d | %6d %6d %9d | %8d | %4d" % (t, us[t - 1], r[0], r[1], r[2], r[3], r[9], r[7], r[8], r[4], r[5], r[6]))
