"""Per-day timeline of the sharded C3 run (heavy variant by default) next to the same run on one GPU: rank 0's stamps.
    torchrun --nproc-per-node N profiles/sharded_timeline.py [n_initial]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import bench
from covid19_demography_b200 import engine, sharding, tables

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("N", "59000000"))
n_initial = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
p, tabs = bench.c3_params(n, n_initial)
pop = bench.make_population(n, "China")
rng = np.random.default_rng(3)
runs = [([50 + k], None, [rng.choice(n, size=n_initial, replace=False).astype(np.int64)]) for k in range(2)]
eng = bench._engine(engine, tables, p, tabs, pop, 1, local)
single = None
if rank == 0:
    for r in runs:
        eng.run(*r)
    single = (eng.day_times_us().copy(), eng.stage_times_us().copy(), eng.last_run_ms3())
dist.barrier()
sharding.attach(eng, pop[0], dist)
for r in runs:
    eng.run(*r)
us, st = eng.day_times_us(), eng.stage_times_us()
tal = torch.from_numpy(eng.tallies(0)).to("cuda:%d" % local)
dist.all_reduce(tal)
tal = tal.cpu().numpy()
if rank == 0:
    act = tal[:, [1, 2, 4, 5], :].sum(axis=(1, 2))
    new = -np.diff(tal[:, 0, :].sum(axis=1))
    print("one GPU init/loop/mat ms:", single[2], " sharded (%d GPUs):" % world, eng.last_run_ms3())
    print("day  active(t-1)  new(t)   one-GPU us | sharded us   stages of rank 0's CTA 0 (setup,A,E,D,C,H,W,end)")
    for t in range(1, len(us) + 1):
        if t % int(os.environ.get("EVERY", "4")) and t > 12:
            continue
        print("%3d  %9d %8d   %8.1f | %8.1f   %s" % (t, act[t - 1], new[t - 1] if t - 1 < len(new) else 0, single[0][t - 1], us[t - 1],
                                                   " ".join("%6.1f" % x for x in st[t, :8])))
dist.destroy_process_group()
