"""CPU tier: the host logic of household-aligned sharding (bounds, ownership gather), world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest

import util
from covid19_demography_b200 import population, sharding


def test_household_starts_and_bounds():
    hh, age, _, _, _ = population.synthetic_population(5000, seed=3)
    starts = sharding.household_starts(hh)
    assert starts[0]
    # a start's row lists only later agents; a non-start's row begins with an earlier agent
    idx = np.arange(len(hh))
    assert ((hh[~starts, 0] < idx[~starts]) & (hh[~starts, 0] >= 0)).all()
    for world in (2, 3, 8):
        b = sharding.shard_bounds(hh, world)
        assert b[0] == 0 and b[-1] == len(hh) and (np.diff(b) > 0).all()
        assert starts[b[1:-1]].all(), "cuts must fall on household starts"
        assert np.abs(np.diff(b) - len(hh) / world).max() <= 6
        # no household row points across a cut
        for g in range(world):
            rows = hh[b[g]:b[g + 1]]
            members = rows[rows >= 0]
            assert members.min() >= b[g] and members.max() < b[g + 1]
    c = util.Case("italy_default")
    b = sharding.shard_bounds(c.households, 2)
    assert sharding.household_starts(c.households)[b[1]]
    with pytest.raises(ValueError):
        sharding.shard_bounds(hh[:3], 8)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bounds = np.array([0, 4, 10], dtype=np.int64)
        full = np.arange(30).reshape(3, 10) * (rank + 1)           # each rank only trusts its own columns
        vec = np.arange(10, dtype=np.float64) + 100 * rank
        q.put((rank, sharding.gather_owned(full, bounds, dist, axis=1), sharding.gather_owned(vec, bounds, dist),
               sharding.sum_over_ranks(np.full((2, 2), rank + 1), dist)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_owned_two_ranks():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = np.arange(30).reshape(3, 10)
    want_hist = np.concatenate([base[:, :4], 2 * base[:, 4:]], axis=1)
    want_vec = np.concatenate([np.arange(4.0), np.arange(4.0, 10.0) + 100])
    for rank, hist, vec, tot in got:
        assert np.array_equal(hist, want_hist) and np.array_equal(vec, want_vec)
        assert np.array_equal(tot, np.full((2, 2), 3))
