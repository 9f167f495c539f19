"""Per-day device timeline of one C2 run (persistent kernel, %globaltimer stamps of CTA 0) next to the
day's active-agent count.   python profiles/day_timeline.py [n] [replicas]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from covid19_demography_b200 import engine, population, tables

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p, tabs = bench.workload_params(n, int(os.environ.get("T", "60")))  # T=30: only sparse passes
if os.environ.get("TRACING"):  # the reference as compiled: tracing switches on at day 37 (run_simulation.py:115-116)
    p["t_tracing_start"] = float(os.environ["TRACING"])
hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                         tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                         tabs["p_critical_death"], with_replay=False)
eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=R)
rng = np.random.default_rng(100)
for k in range(4):
    init = [rng.choice(n, size=int(os.environ.get("NINIT", "5")), replace=False).astype(np.int64) for _ in range(R)]
    eng.run([k * 64 + r for r in range(R)], None, init)
us = eng.day_times_us()
tal = eng.tallies(0)
act = tal[:, [1, 2, 4, 5], :].sum(axis=(1, 2))
print("init/loop/materialize ms:", eng.last_run_ms3(), "sum of passes us: %.1f" % us.sum())
st = eng.stage_times_us()
for t in range(1, len(us) + 1):
    print("pass %2d  %7.2f us   active(t-1)=%-7d stages(setup,A,E,D,C,H,W,end): %s"
          % (t, us[t - 1], act[t - 1] if t - 1 < len(act) else -1, " ".join("%6.2f" % x for x in st[t, :8])))
if os.environ.get("FINE"):
    import ctypes as C
    out = np.zeros((eng.T + 2) * 16, dtype=np.int64)
    eng.lib.seir_get_fine_times.argtypes = [C.c_void_p, C.c_void_p]
    eng.lib.seir_get_fine_times(eng._h, out.ctypes.data)
    f = out.reshape(-1, 16).astype(float)
    day = np.zeros(eng.T + 2, dtype=np.int64)
    eng.lib.seir_get_day_times(eng._h, day.ctypes.data)
    names = "grab entry rec tallies dayblk progr+store hhloads pushes ret append".split()
    if os.environ.get("FINE") == "sparse":  # the stamps of sparse_entry (warp 0 of CTA 0)
        names = "entry+loads dayblk newrec progress hh+knuth gathers tiso hits store dayblk_again".split()
    if os.environ.get("FINE") == "sparse":  # SM cycle counter of CTA 0 / warp 0, relative to the call of sparse_rounds
        order = [14, 10, 11, 0, 1, 2, 3, 4, 5, 6, 7, 8, 13, 15]
        names = "top call entry loads-issued dayblk newrec progress hh+knuth gathers tiso hits store end barrier".split()
        for t in range(1, eng.T):
            base = f[t, 14]
            print("pass %2d cycles: " % t + " ".join("%s=%d" % (nm, f[t, k] - base) for nm, k in zip(names, order) if f[t, k] >= base))
    else:
        for t in range(1, eng.T):
            rel = [(f[t, k] - day[t]) / 1e3 if f[t, k] > 0 else float("nan") for k in range(10)]
            print("pass %2d fine: " % t + " ".join("%s=%.2f" % (n, x) for n, x in zip(names, rel)))
