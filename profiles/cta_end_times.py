import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from covid19_demography_b200 import engine, population, tables
n = 10_000_000
p, tabs = bench.workload_params(n, 60)
hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p), tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"], tabs["p_critical_death"], with_replay=False)
eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1)
rng = np.random.default_rng(100)
for k in range(4):
    eng.run([k * 64], None, [rng.choice(n, size=5, replace=False).astype(np.int64)])
out = np.zeros((eng.T + 2) * 8, dtype=np.int64)
eng.lib.seir_get_stage_times(eng._h, out.ctypes.data)
day = np.zeros(eng.T + 2, dtype=np.int64)
eng.lib.seir_get_day_times(eng._h, day.ctypes.data)
t = int(os.environ.get("PASS", "55"))
rel = (out[:296] - day[t]) / 1e3
print("pass", t, "duration %.1f us" % ((day[t + 1] - day[t]) / 1e3))
print("CTA end times (us after pass start): min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % (rel.min(), np.percentile(rel, 10), np.median(rel), np.percentile(rel, 90), rel.max()))
order = np.argsort(rel)
print("slowest CTAs:", [(int(b), round(float(rel[b]), 1)) for b in order[-8:]])
print("fastest CTAs:", [(int(b), round(float(rel[b]), 1)) for b in order[:8]])
print("by CTA index parity (two CTAs share an SM?): even %.1f odd %.1f ; first 148 %.1f last 148 %.1f" % (rel[0::2].mean(), rel[1::2].mean(), rel[:148].mean(), rel[148:].mean()))
