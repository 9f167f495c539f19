"""Seed ensembles and parameter sweeps over the GPUs of one box.

The reference runs ensembles as a Python `for` loop over seeds inside one process
(`run_simulation.py:437-444`, `simulate_agepolicies.py:605-613`, `parameter_sweep.py:546-554`) and
spreads processes with a SLURM job array (`job.parameter_sweep.sh:15`): runs never talk to each other.
Here the same work is dealt round-robin to one process per GPU (`partition`), each process packs its
share into replica batches that share one device-resident population (`run_seeds`), and the small
per-run results (`[T, 9, 101]` tallies, counters) are gathered on rank 0 (`gather`).  There is no
data-path collective: `torch.distributed` is only used for the final gather.
"""
import numpy as np

from . import tables as _tables


def partition(n_items, rank, world):
    """Indices of the work items rank `rank` of `world` owns: rank, rank + world, ... (round-robin, so
    every rank gets the same count +-1 whatever the order of the items)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside [0, %d)" % (rank, world))
    return list(range(rank, n_items, world))


def batches(items, batch):
    """Consecutive batches of at most `batch` items."""
    if batch < 1:
        raise ValueError("batch must be >= 1")
    return [items[k:k + batch] for k in range(0, len(items), batch)]


def run_seeds(engine, seeds, age_groups, fraction_stay_home, initial_infected_fraction, want_vectors=(),
              want_cohorts=False, fast_prologue=False, host_threads=0, keyed_prologue=False):
    """Run `seeds` on `engine` (an engine.Engine with R replica slots) in batches of R.

    Per seed the reference's own prologue picks are used (`tables.prologue`, seir_individual.py:176-187).
    `fast_prologue` swaps in `tables.prologue_fast` (same distributions, O(picks) instead of two O(n) permutations).
    `keyed_prologue` lets the DEVICE make the picks (keyed permutation, include/seir_draws.h sd_perm; the host
    restatement is tables.prologue_keyed): nothing but 101 doubles per run crosses PCIe, no host work per seed.
    `host_threads` > 0 draws the prologues of the NEXT batch on that many host threads while the device runs the
    current one (NumPy's generators and ctypes' foreign calls release the GIL), so a long ensemble is bound by
    the device, not by the host prologue (0.3 s per seed at n = 10 M for the reference's stream).
    A short last batch is padded by repeating its last seed (the padding runs are discarded).
    Returns {"tallies": int64[len(seeds), T, 9, 101], "counters": [dict], "vectors": {name: float64[len(seeds), n]},
    "cohorts": int64[len(seeds), T, 4, 101] (with want_cohorts)}.
    """
    R = engine.n_replicas
    n = engine.n
    tallies = np.zeros((len(seeds), engine.T, 9, _tables.N_AGES), dtype=np.int64)
    counters = []
    vectors = {k: np.zeros((len(seeds), n)) for k in want_vectors}
    cohorts = np.zeros((len(seeds), engine.T, 4, _tables.N_AGES), dtype=np.int64) if want_cohorts else None
    pick = _tables.prologue_fast if fast_prologue else _tables.prologue
    chunks = batches(list(seeds), R)
    padded = [c + [c[-1]] * (R - len(c)) for c in chunks]

    def draw(batch):
        if keyed_prologue:
            return None
        return [pick(s, age_groups, n, fraction_stay_home, initial_infected_fraction) for s in batch]

    pool = None
    if host_threads > 0 and len(chunks) > 0 and not keyed_prologue:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=host_threads)

        def draw_async(batch):
            futs = [pool.submit(pick, s, age_groups, n, fraction_stay_home, initial_infected_fraction) for s in batch]
            return lambda: [f.result() for f in futs]
        pending = draw_async(padded[0])
    done = 0
    try:
        for b, chunk in enumerate(chunks):
            if pool is not None:
                pro = pending()
                if b + 1 < len(chunks):
                    pending = draw_async(padded[b + 1])  # overlaps with the device run below
            else:
                pro = draw(padded[b])
            if keyed_prologue:
                engine.run(padded[b], keyed=(fraction_stay_home, initial_infected_fraction))
            else:
                engine.run(padded[b], [p[0] for p in pro], [p[1] for p in pro])
            for r in range(len(chunk)):
                tallies[done + r] = engine.tallies(r)
                counters.append(engine.counters(r))
                if want_cohorts:
                    cohorts[done + r] = engine.cohorts(r)
                for k in want_vectors:
                    vectors[k][done + r] = engine.agent_vector(k, replica=r)
            done += len(chunk)
    finally:
        if pool is not None:
            pool.shutdown(wait=True)
    return {"tallies": tallies, "counters": counters, "vectors": vectors, "cohorts": cohorts}


def gather(local_indices, local_array, n_items, dist=None):
    """Assemble the per-item results of all ranks in item order on every rank.

    local_array[k] belongs to item local_indices[k].  With `dist` (an initialised torch.distributed
    module; gloo or nccl) the ranks exchange (indices, array) pairs; without it the call is the
    identity for a single process."""
    local_array = np.asarray(local_array)
    out = np.zeros((n_items,) + local_array.shape[1:], dtype=local_array.dtype)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        pieces = [(list(local_indices), local_array)]
    else:
        pieces = [None] * dist.get_world_size()
        dist.all_gather_object(pieces, (list(local_indices), local_array))
    seen = np.zeros(n_items, dtype=bool)
    for idx, arr in pieces:
        for k, i in enumerate(idx):
            if seen[i]:
                raise RuntimeError("work item %d was produced by two ranks" % i)
            seen[i] = True
            out[i] = arr[k]
    if not seen.all():
        raise RuntimeError("work items %s were produced by no rank" % np.flatnonzero(~seen)[:8].tolist())
    return out


def curves_from_tallies(tallies):
    """[..., T, 9, 101] -> [..., T, 9]: the per-day column sums the drivers compute from the histories
    (run_simulation.py:448-456)."""
    return np.asarray(tallies).sum(axis=-1)
