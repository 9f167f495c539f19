"""Where the end-to-end step of bench.py goes: the same loop as its e2e pass (Engine.run + Engine.tallies per step, host
buffers), with wall-clock stamps around every call and the device time (CUDA events inside seir_run) of every run.
    python profiles/e2e_breakdown.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from covid19_demography_b200 import engine, tables

n, T, K = 10_000_000, 60, 20
p, tabs = bench.workload_params(n, T)
hh, age, diab, hyp, groups = bench.make_population(n)
tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                         tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                         tabs["p_critical_death"], with_replay=False)
eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1)
rng = np.random.default_rng(7)
home = np.zeros(n, dtype=np.uint8)
ins = [([k], [home], [engine.pin(rng.choice(n, size=5, replace=False).astype(np.int64))]) for k in range(K)]
out = [np.zeros((T, 9, 101), dtype=np.int64) for _ in range(K)]
for k in range(5):
    eng.run(*ins[k]); eng.tallies(0, out=out[k])
rows = []
t_all = time.perf_counter()
for k in range(K):
    t0 = time.perf_counter()
    eng.run(*ins[k])
    t1 = time.perf_counter()
    eng.tallies(0, out=out[k])
    t2 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1))
t_all = time.perf_counter() - t_all
dev = []
for k in range(K):   # the same runs again, device times read after each
    eng.run(*ins[k])
    dev.append(sum(eng.last_run_ms3()))
r = np.array(rows) * 1e3
print("per step (ms, mean of %d): loop %.3f = Engine.run %.3f + Engine.tallies %.3f + loop overhead %.3f | device events %.3f | run - device %.3f"
      % (K, t_all / K * 1e3, r[:, 0].mean(), r[:, 1].mean(), t_all / K * 1e3 - r.sum(axis=1).mean(), np.mean(dev), r[:, 0].mean() - np.mean(dev)))
