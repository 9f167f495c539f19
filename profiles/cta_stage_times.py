"""Per-CTA stage stamps of chosen passes (debug build: SEIR_NVCC_EXTRA=-DSEIR_CTA_STAMPS; the pass is picked at
run time with SEIR_DEBUG_PASS): where does the slowest CTA of a pass spend its time?
    python profiles/cta_stage_times.py 10 24 40 55"""
import os, sys
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
names = "setup A E D C H W end".split()
for t in [int(a) for a in sys.argv[1:]] or [10]:
    os.environ["SEIR_DEBUG_PASS"] = str(t)
    eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1)
    rng = np.random.default_rng(100)
    for k in range(4):
        eng.run([k * 64], None, [rng.choice(n, size=5, replace=False).astype(np.int64)])
    out = np.zeros((eng.T + 2) * 16, dtype=np.int64)
    eng.lib.seir_get_fine_times(eng._h, out.ctypes.data)
    day = np.zeros(eng.T + 2, dtype=np.int64)
    eng.lib.seir_get_day_times(eng._h, day.ctypes.data)
    tal = eng.tallies(0)
    st = (out.reshape(-1, 8)[: (eng.T + 2) * 2] - day[t]) / 1e3
    dur = np.diff(np.concatenate([np.zeros((len(st), 1)), st], axis=1), axis=1)
    print("pass %d: %.1f us, active(t-1) = %d; per-CTA stamps of the first %d CTAs (us after the pass start)"
          % (t, (day[t + 1] - day[t]) / 1e3, tal[t - 1, [1, 2, 4, 5]].sum(), len(st)))
    print("   stage durations, median over CTAs: " + " ".join("%s=%.2f" % (nm, x) for nm, x in zip(names, np.median(dur, axis=0))))
    print("   stage durations, max over CTAs   : " + " ".join("%s=%.2f" % (nm, x) for nm, x in zip(names, dur.max(axis=0))))
    order = np.argsort(st[:, 7])
    for b in list(order[-5:][::-1]) + list(order[:2]):
        print("   CTA %3d ends %6.2f : " % (b, st[b, 7]) + " ".join("%s=%.2f" % (nm, x) for nm, x in zip(names, dur[b])))
    eng.close()
