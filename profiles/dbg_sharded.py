"""debug: one sharded fixture run with a replicated phase (threshold 40), per-rank diagnostics"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch.distributed as dist
import util, ctypes as C
from covid19_demography_b200 import engine, sharding, tables as tables_mod
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
c = util.Case(sys.argv[1] if len(sys.argv) > 1 else "china_default")
params = c.no_tracing_params()
eng = engine.Engine(tables_mod.params_to_dict(params), c.households, c.age, c.age_groups, c.diabetes,
                    c.hypertension, c.p_infect_household, c.tables(params), n_replicas=1, device=local, mode="native")
bounds = sharding.attach(eng, c.households, dist, threshold=int(os.environ.get("THR", "40")))
if os.environ.get("FIRST"):
    eng.lib.seir_shard_set_threshold(eng._h, 0)
    home, init = tables_mod.prologue(c.seed, c.age_groups, c.n, c.fraction_stay_home, params["initial_infected_fraction"])
    eng.run([c.seed], [home], [init])
    print(rank, "first run ok", eng.counters(), flush=True)
    if os.environ.get("READ"):
        packed = sharding.gather_owned(eng.history("packed"), bounds, dist, axis=1)
        tal = sharding.sum_over_ranks(eng.tallies(0), dist)
        vecs = {k: sharding.gather_owned(eng.agent_vector(k), bounds, dist) for k in util.VECS}
        print(rank, "read ok", flush=True)
    eng.lib.seir_shard_set_threshold(eng._h, int(os.environ.get("THR", "40")))
seed = c.seed + 7
home, init = tables_mod.prologue(seed, c.age_groups, c.n, c.fraction_stay_home, params["initial_infected_fraction"])
print(rank, "n", c.n, "init", len(init), "T", params["T"], flush=True)
try:
    eng.run([seed], [home], [init])
    ho = np.zeros(8, dtype=np.int32)
    eng.lib.seir_get_handover.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    eng.lib.seir_get_handover(eng._h, ho.ctypes.data, 8)
    print(rank, "ok handover", ho.tolist(), eng.counters(), flush=True)
except Exception as e:
    print(rank, "FAILED", e, flush=True)
dist.barrier()
