"""Multi-process host logic of seed ensembles / sweeps, world_size 2 over gloo on
the CPU: the round-robin deal, replica batching and the gather give the same per-seed results whatever
the number of ranks.  The compute inside each rank is the CPU oracle here (test infrastructure) -- on
the GPU box the same `ensemble` functions drive engine.Engine (tests/test_statistical_parity.py)."""
import os
import socket

import numpy as np
import pytest

import util
from covid19_demography_b200 import ensemble

SEEDS = [11, 12, 13, 14, 15]


def _curves_of(seeds):
    c = util.Case("italy_seeded")
    params = c.no_tracing_params(T=25.0)
    out = []
    for s in seeds:
        tup, _, _, _ = c.keyed_oracle("native", params, seed=s)
        out.append(np.stack([np.asarray(tup[j]).sum(axis=1) for j in range(9)], axis=1))
    return np.asarray(out, dtype=np.int64)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = ensemble.partition(len(SEEDS), rank, world)
        local = _curves_of([SEEDS[i] for i in mine])
        full = ensemble.gather(mine, local, len(SEEDS), dist)
        q.put((rank, mine, full))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_partition_and_batches():
    for world in (1, 2, 3, 8):
        parts = [ensemble.partition(13, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(13))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert ensemble.batches([1, 2, 3, 4, 5], 2) == [[1, 2], [3, 4], [5]]
    assert ensemble.partition(0, 0, 4) == []
    with pytest.raises(ValueError):
        ensemble.partition(4, 4, 4)
    with pytest.raises(ValueError):
        ensemble.batches([1], 0)


def test_gather_detects_missing_and_duplicate_items():
    a = np.arange(6).reshape(3, 2)
    assert np.array_equal(ensemble.gather([0, 1, 2], a, 3), a)
    with pytest.raises(RuntimeError, match="no rank"):
        ensemble.gather([0, 1], a[:2], 3)
    with pytest.raises(RuntimeError, match="two ranks"):
        ensemble.gather([0, 0, 1], a, 3)


def test_two_rank_ensemble_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _curves_of(SEEDS)
    owned = []
    for rank, mine, full in got:
        assert np.array_equal(full, want), "rank %d assembled a different ensemble" % rank
        owned += mine
    assert sorted(owned) == list(range(len(SEEDS)))


def test_threaded_prologue_gives_the_same_runs():
    """host_threads > 0 only moves the prologue draws of the next batch onto host threads."""
    from test_sweep import _FakeEngine
    c = util.Case("italy_policy")
    calls = {}
    for threads in (0, 3):
        eng = _FakeEngine()
        eng.n = c.n
        eng.marker = 1
        seen = []
        orig = eng.run

        def run(seeds, home, init, _orig=orig, _seen=seen):
            _seen.append((tuple(seeds), [h.sum() for h in home], [tuple(i.tolist()) for i in init]))
            _orig(seeds, home, init)
        eng.run = run
        res = ensemble.run_seeds(eng, [5, 6, 7, 8, 9], c.age_groups, c.fraction_stay_home, 0.004, host_threads=threads)
        calls[threads] = (seen, res["tallies"].copy())
    assert calls[0][0] == calls[3][0] and np.array_equal(calls[0][1], calls[3][1])
    assert [s[0] for s in calls[0][0]] == [(5, 6), (7, 8), (9, 9)]
