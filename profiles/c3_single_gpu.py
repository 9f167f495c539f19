"""C3 (Hubei-scale: 59 M agents, China dates T = 127, lockdown day 69 /50, China contact matrix) on ONE B200.
    python profiles/c3_single_gpu.py [n] [n_initial]
Prints init / day loop / materialise device times, agent-days/s and the per-day timeline of the heavy days."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from covid19_demography_b200 import engine, population, tables

n = int(sys.argv[1]) if len(sys.argv) > 1 else 59_000_000
n_init = int(sys.argv[2]) if len(sys.argv) > 2 else 5
fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_china_default.npz"))
p = dict(zip(fx["param_keys"].tolist(), [float(v) for v in fx["param_vals"]]))
p["n"] = float(n)
p["initial_infected_fraction"] = n_init / n
p["t_tracing_start"] = 1.0e6
T = int(p["T"])
t0 = time.perf_counter()
hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1, country="China")
print("population %.1f s" % (time.perf_counter() - t0), flush=True)
tb = tables.build_tables(fx["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                         fx["mean_time_to_isolate_factor"], fx["p_mild_severe"], fx["p_severe_critical"],
                         fx["p_critical_death"], with_replay=False)
t0 = time.perf_counter()
eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, float(fx["p_infect_household_scalar"])), tb, n_replicas=1)
print("engine create %.1f s" % (time.perf_counter() - t0), flush=True)
del hh
rng = np.random.default_rng(3)
for k in range(4):
    init = [rng.choice(n, size=n_init, replace=False).astype(np.int64)]
    eng.run([50 + k], None, init)
    a, b, c = eng.last_run_ms3()
    tal = eng.tallies(0)
    act = tal[:, [1, 2, 4, 5], :].sum(axis=(1, 2))
    print("run %d: init %.3f ms, day loop %.3f ms, materialise %.3f ms -> %.3e agent-days/s (device total); peak active %d, infected %d"
          % (k, a, b, c, n * (T - 1) / ((a + b + c) / 1e3), act.max(), n - tal[-1, 0].sum()), flush=True)
us = eng.day_times_us()
for t in range(1, len(us) + 1, 6):
    print("pass %3d %8.2f us active(t-1)=%d" % (t, us[t - 1], act[t - 1] if t - 1 < len(act) else -1))
