import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from covid19_demography_b200 import engine, tables
n = 10_000_000
p, tabs = bench.workload_params(n, 60)
pop = bench.make_population(n)
hh, age, diab, hyp, groups = pop
tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                         tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                         tabs["p_critical_death"], with_replay=False)
eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=1)
rng = np.random.default_rng(100)
inits = [engine.pin(rng.choice(n, size=5, replace=False).astype(np.int64)) for _ in range(40)]
home = np.zeros(n, dtype=np.uint8)
lib = eng.lib
orig = lib.seir_run_ex
ccall = [0.0]
class W:
    def __call__(self, *a):
        t = time.perf_counter(); r = orig(*a); ccall[0] = time.perf_counter() - t; return r
lib_run = W()
rows = []
for k, init in enumerate(inits):
    eng.lib = type("L", (), {"__getattr__": lambda s, nm: lib_run if nm == "seir_run_ex" else getattr(lib, nm)})()
    t0 = time.perf_counter()
    eng.run([k], [home], [init])
    t1 = time.perf_counter()
    eng.lib = lib
    tal = eng.tallies(0)
    t2 = time.perf_counter()
    dev = sum(eng.last_run_ms3())
    rows.append(((t1 - t0) * 1e3, ccall[0] * 1e3, dev, (t2 - t1) * 1e3))
r = np.array(rows[5:])
print("median ms: Engine.run wall %.3f | C call %.3f | device events %.3f | tallies %.3f" % tuple(np.median(r, axis=0)))
print("python prep in Engine.run %.3f | C call minus device %.3f" % (np.median(r[:,0]-r[:,1]), np.median(r[:,1]-r[:,2])))
