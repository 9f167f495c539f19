"""One population across the GPUs of one box (BASELINE.json config C3).

Households are contiguous index ranges (sample_households.py:344-349), so a cut at a household start
keeps household transmission (seir_individual.py:339-359) inside a shard.  Community contacts
(:361-390) pick their target uniformly over the GLOBAL age group: when the target lives on another
GPU the sender's kernel reads the target's `winner` word, applies the `atomicMax`, sets the inbox bits
and appends the target to the owner's list of tomorrow directly in the owner's memory over NVLink
(CUDA IPC mappings, system-scope atomics); the GPUs meet at a device-side barrier once per day.  No
host round trip and no NCCL call inside the day loop: torch.distributed only carries the 384-byte IPC
handles at set-up and the small results at the end.

Every rank holds the full-size arrays in global agent indexing (C3: 11 GB per GPU) and steps the agents
it owns; results are bit-identical to the single-GPU run because every draw is keyed on
(agent, day, site) -- never on who executes it.

A run starts REPLICATED: while the epidemic is small (a day's lists below `threshold` entries) every GPU steps
the whole population on its own -- deterministic runs stay identical without exchanging a byte or meeting at a
barrier, so a sparse day costs what it costs on one GPU -- and on the first longer day all GPUs switch, for good,
to sharded passes.  Nothing moves at the switch: everybody holds everything up to that day.
"""
import numpy as np


def household_starts(households):
    """bool[n]: agent i is the first member of its household (rows list the OTHER members in ascending order)."""
    hh = np.asarray(households)
    first = hh[:, 0]
    idx = np.arange(hh.shape[0], dtype=first.dtype)
    return (first == -1) | (first > idx)


def shard_bounds(households, world):
    """int64[world + 1]: near-equal contiguous ranges cut at household starts."""
    n = np.asarray(households).shape[0]
    starts = np.flatnonzero(household_starts(households))
    bounds = [0]
    for g in range(1, world):
        target = (n * g) // world
        k = int(np.searchsorted(starts, target))
        cut = int(starts[min(k, len(starts) - 1)])
        if cut <= bounds[-1]:
            raise ValueError("population too small for %d shards" % world)
        bounds.append(cut)
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def attach(engine, households, dist, threshold=None):
    """Exchange the IPC handles over an initialised torch.distributed group and attach the peers.  `threshold`: see
    Engine.shard_attach (the run stays replicated -- every GPU steps everybody, no exchange -- while the day's lists
    are shorter)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    bounds = shard_bounds(households, world)
    mine = engine.shard_export()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tobytes())
    handles = np.stack([np.frombuffer(b, dtype=np.uint8) for b in gathered])
    engine.shard_attach(rank, world, handles, bounds, threshold=threshold)
    dist.barrier()
    return bounds


def gather_owned(local_array, bounds, dist, axis=-1):
    """Assemble a per-agent result (vector [n] or history [T, n]) from the owners' slices, on every rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    a = np.asarray(local_array)
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(int(bounds[rank]), int(bounds[rank + 1]))
    pieces = [None] * world
    dist.all_gather_object(pieces, np.ascontiguousarray(a[tuple(sl)]))
    return np.concatenate(pieces, axis=axis)


def sum_over_ranks(local_array, dist):
    """Element-wise sum of a small per-rank array (the [T, 9, 101] tallies add up over the shards)."""
    pieces = [None] * dist.get_world_size()
    dist.all_gather_object(pieces, np.asarray(local_array))
    return np.sum(pieces, axis=0)
