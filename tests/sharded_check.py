"""Multi-GPU parity check of the sharded path (run under torchrun on a box with >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/sharded_check.py

One population split over WORLD_SIZE GPUs by whole households, cross-shard community contacts resolved in
peer memory over NVLink, one device-side barrier per day -- compared bit for bit with the CPU oracle under
the same keyed draws (nine histories, per-agent vectors, per-day per-age tallies), both community modes."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE):
    sys.path.insert(0, p)

import torch.distributed as dist  # noqa: E402

import util  # noqa: E402
from covid19_demography_b200 import engine, sharding, tables as tables_mod  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    failures = 0
    cases = (("italy_seeded", ("native", "replay")), ("ensemble_italy_n20000", ("native",)), ("china_default", ("native",)))
    if len(sys.argv) > 1:  # (debugging: only the named cases)
        cases = tuple(c_ for c_ in cases if c_[0] in sys.argv[1:])
    for name, modes in cases:
        c = util.Case(name)
        params = c.no_tracing_params()
        T = int(params["T"])
        for mode in modes:
            eng = engine.Engine(tables_mod.params_to_dict(params), c.households, c.age, c.age_groups, c.diabetes,
                                c.hypertension, c.p_infect_household, c.tables(params), n_replicas=1, device=local,
                                mode=mode)
            # the first seed sharded from day 1, the second replicated until a day lists more than 40 agents
            bounds = sharding.attach(eng, c.households, dist, threshold=0)
            for seed in (c.seed, c.seed + 7):
                eng.lib.seir_shard_set_threshold(eng._h, 0 if seed == c.seed else 40)
                home, init = tables_mod.prologue(seed, c.age_groups, c.n, c.fraction_stay_home,
                                                 params["initial_infected_fraction"])
                eng.run([seed], [home], [init])
                packed = sharding.gather_owned(eng.history("packed"), bounds, dist, axis=1)
                tal = sharding.sum_over_ranks(eng.tallies(0), dist)
                vecs = {k: sharding.gather_owned(eng.agent_vector(k), bounds, dist) for k in util.VECS}
                cnt = eng.counters()
                if rank == 0:
                    tup, ex, _, _ = c.keyed_oracle(mode, params, seed=seed)
                    ok = np.array_equal(packed, util.pack_history(tup[:9]))
                    ok = ok and np.array_equal(tal, util.tallies_from_history(tup[:9], c.age, T))
                    for k in util.VECS:
                        ok = ok and np.array_equal(vecs[k], tup[util.VEC_INDEX[k]], equal_nan=True)
                    print("sharded x%d %-22s %-6s seed %d: %s (bounds %s, infections %d, day loop %.3f ms on rank 0)"
                          % (world, name, mode, seed, "bit-exact vs oracle" if ok else "MISMATCH", bounds.tolist(),
                             int((tup[15] > 0).sum()), eng.last_run_ms()[0]), flush=True)
                    failures += 0 if ok else 1
                dist.barrier()
            eng.close()
    res = [failures]
    dist.broadcast_object_list(res, src=0)
    dist.destroy_process_group()
    sys.exit(1 if res[0] else 0)


if __name__ == "__main__":
    main()
