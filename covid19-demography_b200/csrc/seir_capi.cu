// seir_capi.cu -- kernels' launch side and the C ABI declared in include/seir_b200.h.
// Built into libseir_b200.so by covid19-demography_b200/build.py (nvcc, sm_100a, -lineinfo).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <atomic>
#include <vector>

#include "seir_kernels.cuh"
#ifdef SEIR_SORT_PROBE  // (experiment) what would ID-ordered active lists be worth?  see the per-day branch of seir_run
#include <thrust/execution_policy.h>
#include <thrust/sort.h>
#endif

using namespace seir;

#define SEIR_STR2(x) #x
#define SEIR_STR(x) SEIR_STR2(x)

// ------------------------------------------------------------------ kernels
extern __shared__ __align__(16) unsigned char seir_smem_raw[];

template <bool TR>
__global__ void __launch_bounds__(kThreads, kMinBlocks) seir_persistent_kernel(const __grid_constant__ DevCtx cx)
{
    PassSmem &sm = *reinterpret_cast<PassSmem *>(seir_smem_raw);
    // the host resets the barrier counter to 0 before each launch
    stage_tables(cx, sm);
    unsigned int epoch = 0, epoch_m = 0;  // grid barriers passed so far: this GPU's own, and the ones across the GPUs of a sharded run
    const int reps_per_sweep = gridDim.x / cx.blocks_per_rep;  // >= 1
    const int my_slot = blockIdx.x / cx.blocks_per_rep, my_part = blockIdx.x % cx.blocks_per_rep;
    // A sharded run starts REPLICATED: while the epidemic is small every GPU steps the whole population on its own (the
    // runs are deterministic, so the GPUs stay identical without exchanging a byte or meeting at a barrier -- a sparse
    // day costs what it costs on one GPU).  On the first day whose lists exceed shard_threshold entries all GPUs switch,
    // for good, to sharded passes: each keeps the agents it owns, cross-shard hits go to the owner's memory over NVLink
    // and the GPUs meet once per day.  No state moves at the switch: everybody holds everything up to that day.
    bool sharded_now = cx.n_shards > 1 && cx.shard_threshold == 0u;
    if (sharded_now) {  // no GPU touches a peer's arrays before every peer has finished (re-)initialising them
        ++epoch_m;
        grid_barrier_multi(cx, cx.barrier, epoch_m * gridDim.x, cx.epoch_base + epoch_m);
    }
    const bool defer_counters = cx.n_rep == 1;  // this CTA serves the one replica throughout: its run counters are added once, below
    if (threadIdx.x < 3) sm.run_cnt[threadIdx.x] = 0;
    int bank_staged = -1;  // the table set whose kept-contact table sits in shared memory
    // (a sharded run's replicated sparse / dense days may have been run by the single-replica kernels: go on from their day)
    const int t_first = cx.start_word >= 0 ? __ldcg(cx.ctl + cx.start_word) + 1 : 1;
    for (int t = t_first; t <= cx.T_loop; ++t) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.day_ns[t] = global_timer_ns();
        if (cx.n_shards > 1 && !sharded_now) {  // (the day's list lengths are the same on every GPU: so is the decision)
            __syncthreads();
            if (threadIdx.x == 0) {
                const uint32_t *c0 = cx.list_cnt + t;
                sm.base_new = __ldcg(c0) + __ldcg(c0 + (cx.T + 2)) + __ldcg(c0 + 2 * (cx.T + 2));
            }
            __syncthreads();
            if (sm.base_new > cx.shard_threshold) {
                sharded_now = true;
                ++epoch_m;
                grid_barrier_multi(cx, cx.barrier, epoch_m * gridDim.x, cx.epoch_base + epoch_m);
            }
            __syncthreads();
        }
        if (sharded_now) {  // every GPU's copy of the infection days, from the owners' lists of today's new infectees
            if (xday_refresh(cx, sm, t)) grid_barrier(cx.barrier, ++epoch * gridDim.x);
        }
        Pass<TR>::template day_pass<true>(cx, sm, t, my_slot, my_part, reps_per_sweep, bank_staged, defer_counters, sharded_now);
        if (t < cx.T_loop) {
            if (sharded_now) {
                ++epoch_m;
                grid_barrier_multi(cx, cx.barrier, epoch_m * gridDim.x, cx.epoch_base + epoch_m);
                continue;  // (tracing is not available in sharded runs)
            }
            grid_barrier(cx.barrier, ++epoch * gridDim.x);
            if (TR && cx.trace_from >= 1 && t >= cx.trace_from) {  // :241-242 tracing is on: resolve the day's traces
                // any root today?  (grid-uniform: the counts are final after the barrier; read by ONE warp per CTA --
                // the same word requested by every warp of the grid queues up at its L2 slice)
                if (threadIdx.x < 32) {
                    bool mine = false;
                    for (int r = threadIdx.x; r < cx.n_rep; r += 32) mine = mine || __ldcg(cx.root_cnt + (size_t)r * (cx.T + 2) + t) != 0u;
                    const bool warp_any = __any_sync(0xFFFFFFFFu, mine);
                    if (threadIdx.x == 0) sm.base_new = warp_any ? 1u : 0u;
                }
                __syncthreads();
                const bool any = sm.base_new != 0u;
                __syncthreads();  // (base_new is reused below)
                if (any) {
                    if (cx.tcomp) trace_parallel(cx, sm, t, epoch);  // components in parallel, ordered inside
                    else trace_pass(cx, *reinterpret_cast<TraceSmem *>(sm.out_old), t);
                    grid_barrier(cx.barrier, ++epoch * gridDim.x);
                    if (blockIdx.x == 0 && threadIdx.x == 0) cx.stage_ns[(size_t)t * 8 + 6] = global_timer_ns();
                }
            }
        }
    }
    if (defer_counters && threadIdx.x < 3 && my_slot < cx.n_rep && sm.run_cnt[threadIdx.x])
        atomicAdd(&cx.counters[my_slot * 4 + 1 + threadIdx.x], (unsigned long long)sm.run_cnt[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) cx.day_ns[cx.T + 1] = global_timer_ns();
}

// The day loop of ONE replica on one GPU without contact tracing (C1, C2, C3 on one GPU: the bench's default) is run by
// two kernels in turn.  seir_sparse_kernel takes the days whose lists fit one entry per warp (warp-cooperative passes,
// sparse_entry): one CTA per SM, so a thread may use 128 registers -- the pass is ONE warp's dependent chain, and a spill
// is a round trip to L2 behind the barrier's invalidated L1 -- a barrier over 148 CTAs, a code footprint of a few KB.
// seir_dense_solo_kernel takes the others (staged passes, two CTAs per SM).  Each leaves the first day it does not take
// in ctl[k + 1] and the host has queued the next launch already: no host round trip at a hand-over.
__global__ void __launch_bounds__(kSparseThreads, 1) seir_sparse_kernel(const __grid_constant__ DevCtx cx, int k)
{
    PassSmem &sm = *reinterpret_cast<PassSmem *>(seir_smem_raw);
    int t = __ldcg(cx.ctl + k) + 1;
    if (t > cx.T_loop) {  // nothing left to do: hand the word on
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.ctl[k + 1] = t - 1;
        return;
    }
    stage_tables(cx, sm);
    unsigned int *bar = cx.barrier + (size_t)(1 + k) * kBarStride;  // this launch's barrier word (zeroed by the host with the others)
    unsigned int epoch = 0;
    if (threadIdx.x < 3) sm.run_cnt[threadIdx.x] = 0;
    if (threadIdx.x < 4) sm.cnt[threadIdx.x] = 0;
    int bank_staged = -1;
    if (threadIdx.x == 0) {
        Pass<false>::fill_view(cx, sm, t, 0, false);
        Pass<false>::load_list_lengths(cx, sm, t, 0);
    }
    __syncthreads();
    for (; t <= cx.T_loop; ++t) {
        // (the view and the list lengths of day t are in shared memory: set up before the loop, or in the shadow of the barrier below)
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.day_ns[t] = global_timer_ns();  // (a dense day: the other kernel stamps it again)
#ifdef SEIR_FINE_STAMPS
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.fine_ns[(size_t)t * 16 + 14] = (unsigned long long)clock64();
#endif
        const int T = sm.rp.T;
        const bool finalize = t >= T;
        const uint32_t n_new = sm.list_n[0], n_old = finalize ? 0u : sm.list_n[1], n_pk = finalize ? 0u : sm.list_n[2];
        if (sm.list_n[0] + sm.list_n[1] + sm.list_n[2] > cx.sparse_cap) break;  // (grid-uniform: a dense day)
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.stage_ns[(size_t)t * 8 + 0] = global_timer_ns();
        if (bank_staged != sm.rp.bank) {
            const double *src = cx.bank + (size_t)sm.rp.bank * kBankStride + kBankNatEn;
            for (int q = threadIdx.x; q < 4 * kAges; q += kSparseThreads) sm.nat_enlam[q] = src[q];
            bank_staged = sm.rp.bank;
            __syncthreads();
        }
#ifdef SEIR_FINE_STAMPS
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.fine_ns[(size_t)t * 16 + 10] = (unsigned long long)clock64();
#endif
        Pass<false>::sparse_rounds(cx, sm, t, n_new, n_old, n_pk, cx.list_new + (size_t)(t & 1) * cx.n_pad, cx.list_old + (size_t)(t & 1) * cx.n_pad,
                                   cx.pk_pool + (size_t)(t & (kParkRing - 1)) * cx.n_pad, (int)blockIdx.x, (int)gridDim.x, finalize, false);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const unsigned long long now = global_timer_ns();
            for (int q = 1; q < 8; ++q) cx.stage_ns[(size_t)t * 8 + q] = now;
        }
        // The grid barrier (flat form of grid_barrier) with the next pass's set-up in its shadow: between its arrival and
        // its poll thread 0 does what needs nothing from the other CTAs (this CTA's run counters, tomorrow's view), and as
        // soon as the poll succeeds it fetches tomorrow's list lengths -- the closing __syncthreads() hands both on.
        __syncthreads();
        if (t < cx.T_loop) ++epoch;
        if (threadIdx.x == 0 && t < cx.T_loop) {
            unsigned int seen;
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(bar), "r"(1u) : "memory");
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
            } while (seen < epoch * gridDim.x);
            Pass<false>::load_list_lengths(cx, sm, t + 1, 0);
        }
        if (threadIdx.x == 32) {  // (another warp: in the shadow of thread 0's round trips even in the last CTA to arrive)
            Pass<false>::flush_pass_counters(cx, sm, 0, true);
            if (t < cx.T_loop) Pass<false>::fill_view(cx, sm, t + 1, 0, false);
        }
        __syncthreads();
#ifdef SEIR_FINE_STAMPS
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.fine_ns[(size_t)t * 16 + 15] = (unsigned long long)clock64();
#endif
    }
    if (threadIdx.x < 3 && sm.run_cnt[threadIdx.x]) atomicAdd(&cx.counters[1 + threadIdx.x], (unsigned long long)sm.run_cnt[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cx.ctl[k + 1] = t - 1;
        if (t > cx.T_loop) cx.day_ns[cx.T + 1] = global_timer_ns();
    }
}

// The first days of a run (a handful of cases: at most one list entry per warp of EIGHT CTAs) are run by one thread-block
// CLUSTER: the same warp-cooperative passes, but the CTAs of a day meet at the hardware cluster barrier
// (barrier.cluster.arrive.release / wait.acquire, SASS: UCGABAR_ARV / UCGABAR_WAIT) instead of an atomic word in L2 polled
// by 148 CTAs, and the other 140 SMs stay idle.  Launched in front of seir_sparse_kernel; hands over like the others.
#ifndef SEIR_CLUSTER_CTAS
#define SEIR_CLUSTER_CTAS 8   // portable cluster size; one CTA per SM (the sparse pass's 126 registers x 512 threads)
#endif
constexpr int kClusterCtas = SEIR_CLUSTER_CTAS;
__global__ void __launch_bounds__(kSparseThreads, 1) seir_cluster_kernel(const __grid_constant__ DevCtx cx, int k, uint32_t cap)
{
    PassSmem &sm = *reinterpret_cast<PassSmem *>(seir_smem_raw);
    int t = __ldcg(cx.ctl + k) + 1;
    if (t > cx.T_loop) {  // nothing left to do: hand the word on
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.ctl[k + 1] = t - 1;
        return;
    }
    stage_tables(cx, sm);
    if (threadIdx.x < 3) sm.run_cnt[threadIdx.x] = 0;
    if (threadIdx.x < 4) sm.cnt[threadIdx.x] = 0;
    int bank_staged = -1;
    if (threadIdx.x == 0) {
        Pass<false>::fill_view(cx, sm, t, 0, false);
        Pass<false>::load_list_lengths(cx, sm, t, 0);
    }
    __syncthreads();
    for (; t <= cx.T_loop; ++t) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.day_ns[t] = global_timer_ns();
        const int T = sm.rp.T;
        const bool finalize = t >= T;
        const uint32_t n_new = sm.list_n[0], n_old = finalize ? 0u : sm.list_n[1], n_pk = finalize ? 0u : sm.list_n[2];
        if (sm.list_n[0] + sm.list_n[1] + sm.list_n[2] > cap) break;  // (cluster-uniform: the day goes to the 148-CTA kernels)
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.stage_ns[(size_t)t * 8 + 0] = global_timer_ns();
        if (bank_staged != sm.rp.bank) {
            const double *src = cx.bank + (size_t)sm.rp.bank * kBankStride + kBankNatEn;
            for (int q = threadIdx.x; q < 4 * kAges; q += kSparseThreads) sm.nat_enlam[q] = src[q];
            bank_staged = sm.rp.bank;
            __syncthreads();
        }
        Pass<false>::sparse_rounds(cx, sm, t, n_new, n_old, n_pk, cx.list_new + (size_t)(t & 1) * cx.n_pad, cx.list_old + (size_t)(t & 1) * cx.n_pad,
                                   cx.pk_pool + (size_t)(t & (kParkRing - 1)) * cx.n_pad, (int)blockIdx.x, (int)gridDim.x, finalize, false);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const unsigned long long now = global_timer_ns();
            for (int q = 1; q < 8; ++q) cx.stage_ns[(size_t)t * 8 + q] = now;
        }
        __syncthreads();  // (every warp of this CTA is through with day t: its view may be replaced)
        // the cluster barrier, with the next pass's set-up between arrival and wait (what needs nothing from the other CTAs)
        if (t < cx.T_loop) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        if (threadIdx.x == 32) {
            Pass<false>::flush_pass_counters(cx, sm, 0, true);
            if (t < cx.T_loop) Pass<false>::fill_view(cx, sm, t + 1, 0, false);
        }
        if (t < cx.T_loop) {
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            if (threadIdx.x == 0) Pass<false>::load_list_lengths(cx, sm, t + 1, 0);
        }
        __syncthreads();
    }
    if (threadIdx.x < 3 && sm.run_cnt[threadIdx.x]) atomicAdd(&cx.counters[1 + threadIdx.x], (unsigned long long)sm.run_cnt[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cx.ctl[k + 1] = t - 1;
        if (t > cx.T_loop) cx.day_ns[cx.T + 1] = global_timer_ns();
    }
}

__global__ void __launch_bounds__(kThreads, kMinBlocks) seir_dense_solo_kernel(const __grid_constant__ DevCtx cx, int k, uint32_t yield_below, uint32_t yield_above)
{
    PassSmem &sm = *reinterpret_cast<PassSmem *>(seir_smem_raw);
    int t = __ldcg(cx.ctl + k) + 1;
    if (t > cx.T_loop) {  // nothing left to do: hand the word on
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.ctl[k + 1] = t - 1;
        return;
    }
    stage_tables(cx, sm);
    unsigned int *bar = cx.barrier + (size_t)k * kBarStride;  // (grid_barrier counts in the word one stride further)
    unsigned int epoch = 0;
    if (threadIdx.x < 3) sm.run_cnt[threadIdx.x] = 0;
    int bank_staged = -1;
    for (; t <= cx.T_loop; ++t) {
        if (blockIdx.x == 0 && threadIdx.x == 0) cx.day_ns[t] = global_timer_ns();
        if (Pass<false>::day_pass<false, true>(cx, sm, t, 0, (int)blockIdx.x, 1, bank_staged, true, false, yield_below, yield_above)) break;
        if (t < cx.T_loop) grid_barrier(bar, ++epoch * gridDim.x);
    }
    __syncthreads();
    if (threadIdx.x < 3 && sm.run_cnt[threadIdx.x]) atomicAdd(&cx.counters[1 + threadIdx.x], (unsigned long long)sm.run_cnt[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cx.ctl[k + 1] = t - 1;
        if (t > cx.T_loop) cx.day_ns[cx.T + 1] = global_timer_ns();
    }
}

template <bool TR>
__global__ void __launch_bounds__(kThreads, kMinBlocks) seir_day_kernel(const __grid_constant__ DevCtx cx, int t)
{
    PassSmem &sm = *reinterpret_cast<PassSmem *>(seir_smem_raw);
    if (blockIdx.x == 0 && threadIdx.x == 0) cx.day_ns[t] = global_timer_ns();
    stage_tables(cx, sm);
    int bank_staged = -1;
    Pass<TR>::template day_pass<true>(cx, sm, t, blockIdx.x / cx.blocks_per_rep, blockIdx.x % cx.blocks_per_rep, gridDim.x / cx.blocks_per_rep, bank_staged);
}

// contact tracing of day t as its own launch (per-day launch mode)
__global__ void __launch_bounds__(kThreads, kMinBlocks) seir_trace_kernel(const __grid_constant__ DevCtx cx, int t)
{
    trace_pass(cx, *reinterpret_cast<TraceSmem *>(seir_smem_raw), t);
}

// winner := 0, inbox := 0 (the only per-run state that must be cleared; records are built on infection)
// (and the run's small words -- tallies, run counters, list lengths, barrier and hand-over words: one launch instead of
// five memsets in front of it)
__global__ void seir_init_state_kernel(DevCtx cx, uint32_t *barrier_words, int n_barrier_words)
{
    {
        const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
        const int64_t n_tal = (int64_t)cx.n_rep * cx.T * TAL_ROWS * kAges, n_cnt = (int64_t)cx.n_rep * 3 * (cx.T + 2);
        for (int64_t k = tid; k < n_tal; k += nth) cx.tally[k] = 0u;
        for (int64_t k = tid; k < n_cnt; k += nth) cx.list_cnt[k] = 0u;
        for (int64_t k = tid; k < (int64_t)cx.n_rep * 4; k += nth) cx.counters[k] = 0ull;
        for (int64_t k = tid; k < n_barrier_words; k += nth) barrier_words[k] = 0u;
        for (int64_t k = tid; k <= kMaxLaunches; k += nth) cx.ctl[k] = 0;
    }
    // every warp instruction stores 512 consecutive bytes (lane l the l-th 16 B of them)
    const int64_t cells = cx.n_pad * (int64_t)cx.n_rep;   // (n_pad is a multiple of kVec = 16 agents)
    const uint4 z = make_uint4(0, 0, 0, 0);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    uint4 *wz = reinterpret_cast<uint4 *>(cx.winner);
    for (int64_t g = tid; g < cells / 2; g += nth) __stcs(wz + g, z);             // 8 B per agent
    uint4 *iz = reinterpret_cast<uint4 *>(cx.inbox);
    for (int64_t g = tid; g < cells / 64; g += nth) __stcs(iz + g, z);            // 2 bits per agent
    for (int64_t g = cells / 64 * 4 + tid; g < cells / 16; g += nth) cx.inbox[g] = 0u;   // (the words behind the last whole 16 B)
    if (cx.n_shards > 1) {  // (the infection days every GPU of a sharded run keeps)
        uint4 *xz = reinterpret_cast<uint4 *>(cx.xday);
        for (int64_t g = tid; g < cells / 8; g += nth) __stcs(xz + g, z);         // 2 B per agent
    }
    if (cx.log_head) {  // tracing state: log heads (the tracer's marks are stamped with the run number)
        uint4 *hz = reinterpret_cast<uint4 *>(cx.log_head);
        for (int64_t g = tid; g < cells / 4; g += nth) __stcs(hz + g, z);         // 4 B per agent
    }
}

// Keyed prologue (:176-187 on the device): replica blockIdx.y makes its own picks.  kcum[rep][0..102] are the prefix
// sums of the pick counts -- per age int(fraction_stay_home[a] * N_a), then the initial cases int(fraction * n) --
// or all 0 for a replica whose picks came from the host.  Pick r of group g is entry r of the keyed permutation of
// the group (include/seir_draws.h, sd_perm): no scan over the n candidates, no counter to contend for.
__global__ void seir_prologue_kernel(DevCtx cx, const int64_t *kcum_all, int32_t *init_ids, const int64_t *init_off, int home_used)
{
    __shared__ int64_t kcum[kAges + 2];
    __shared__ sd_perm perm[kAges + 1];
    const int rep = blockIdx.y;
    for (int g = threadIdx.x; g < kAges + 2; g += blockDim.x) kcum[g] = kcum_all[(size_t)rep * (kAges + 2) + g];
    __syncthreads();
    const int64_t total = kcum[kAges + 1];
    if (total == 0) return;
    const uint64_t seed = cx.seeds[rep];
    for (int g = threadIdx.x; g <= kAges; g += blockDim.x) {
        const uint32_t N = g < kAges ? (uint32_t)(cx.grp_off[g + 1] - cx.grp_off[g]) : (uint32_t)cx.n;
        if (kcum[g + 1] > kcum[g]) perm[g] = sd_perm_make(seed, g < kAges ? SD_SITE_HOME : SD_SITE_CASES, g < kAges ? (uint32_t)g : 0u, N);
    }
    __syncthreads();
    uint32_t *home = cx.home_rw + (size_t)rep * (cx.n_pad / 32);
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        int lo = 0, hi = kAges + 1;  // group g with kcum[g] <= q < kcum[g + 1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (kcum[mid] <= q) lo = mid; else hi = mid; }
        const int g = lo;
        const uint32_t x = sd_perm_at(&perm[g], (uint32_t)(q - kcum[g]));
        if (g < kAges) {
            if (home_used) {
                const uint32_t a = (uint32_t)cx.grp_members[cx.grp_off[g] + (int64_t)x];
                atomicOr(home + (a >> 5), 1u << (a & 31));
            }
        } else {
            init_ids[init_off[rep] + (q - kcum[kAges])] = (int32_t)x;
        }
    }
}

// SEIR_HISTORY_ON_DEMAND: rows [t0, t1) of one replica's packed history straight from the records
__global__ void seir_rows_kernel(DevCtx cx, int rep, int t0, int t1, uint8_t *out)
{
    const uint32_t *inbox = cx.inbox + (size_t)rep * (cx.n_pad / 16);
    const AgentRec *rec = cx.rec + (size_t)rep * cx.n_pad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cx.n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool ever = (inbox[i >> 4] & inbox_infected_bit((int)(i & 15))) != 0;
        if (!ever) {
            for (int t = t0; t < t1; ++t) out[(size_t)(t - t0) * cx.n_pad + i] = 0;
            continue;
        }
        const AgentRec r = load_rec(rec + i);
        const int t_q2 = ((r.flags() & F_TRX) && cx.trc) ? (int)(__ldcg(cx.trc + (size_t)rep * cx.n_pad + i) & 0xFFFFu) : 0xFFFF;
        for (int t = t0; t < t1; ++t) out[(size_t)(t - t0) * cx.n_pad + i] = state_on_day(r, t, t_q2);
    }
}

// :220-227 initial cases: E at t=0, activation time drawn at site INIT, time_to_isolate = 0
__global__ void seir_init_infected_kernel(DevCtx cx, const int32_t *init_ids, const int64_t *init_off)
{
    const int rep = blockIdx.y;
    const int64_t lo = init_off[rep], hi = init_off[rep + 1];
    for (int64_t k = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < hi; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = init_ids[k];
        const bool own = cx.n_shards <= 1 || (a >= cx.shard_lo[cx.my_shard] && a < cx.shard_lo[cx.my_shard + 1]);
        if (cx.n_shards > 1) cx.xday[(size_t)rep * cx.n_pad + a] = 1;  // infected on day 0 (every GPU of a sharded run keeps the whole array)
        if (!own && cx.shard_threshold == 0u) continue;  // another GPU's agent (a run that starts replicated steps it here too)
        const uint64_t seed = cx.seeds[rep];
        AgentRec r;
        r.clear();
        r.set_t_exposed(0);
        r.set_tt_activation(sd_enc16(sd_lognormal_days(seed, (uint32_t)a, 0, SD_SITE_INIT, 0, cx.act_mu, cx.act_sigma)));
        r.set_t_end(NEVER);
        r.set_t_q_on(NEVER);
        r.set_stat(cx.stat[a]);
        r.set_state(C_E);
        store_rec(cx.rec + (size_t)rep * cx.n_pad + a, r);
        cx.winner[(size_t)rep * cx.n_pad + a] = 1ull;  // day 0, never susceptible again
        atomicOr(cx.inbox + (size_t)rep * (cx.n_pad / 16) + (a >> 4), inbox_infected_bit((int)(a & 15)));
        cx.list_old[((size_t)rep * 2 + 1) * cx.n_pad + atomicAdd(cx.list_cnt + (size_t)rep * 3 * (cx.T + 2) + 1, 1u)] = (uint32_t)a;
        if (own) atomicAdd(cx.tally + (size_t)rep * cx.T * TAL_ROWS * kAges + TAL_NEW_E * kAges + (r.stat() & 127), 1u);
    }
}

// The reference's nine bool[T,n] outputs (:165-173, one row per day) as ONE packed byte per agent-day,
// written once from the per-agent records: a streaming T*n byte write.
// A WARP owns 512 consecutive agents (lane = one 16 B vector per row).  Its ever-infected agents are listed
// in shared memory and dealt round-robin to the lanes, so the per-row work is balanced whatever the
// clustering of infections; each agent's timeline is seven day thresholds held in registers, a row byte
// is a handful of branch-free compares dropped into a 512 B shared tile, and every row leaves the warp as
// full 16 B vector stores.
constexpr int kMatWarps = 8, kMatThreads = kMatWarps * 32, kMatPerLane = 4;
struct MatAgent { uint32_t a, b, c, d, e; };  // t_exp|t_inf<<16, t_sev|t_crit<<16, t_end|t_q_on<<16, t_doc|final<<16|local<<20, t_q2

__device__ __forceinline__ MatAgent mat_pack(const AgentRec &r, int local, const uint32_t *trc_of_agent)
{
    const uint32_t fl = r.flags();
    const uint32_t t_exp = r.t_exposed();
    const uint32_t t_inf = (fl & F_MILD) ? (uint32_t)t_infected_of(r) : 0xFFFFu;
    const uint32_t t_sev = (fl & F_SEV) ? (uint32_t)t_severe_of(r) : 0xFFFFu;
    const uint32_t t_cri = (fl & F_CRIT) ? (uint32_t)t_critical_of(r) : 0xFFFFu;
    const uint32_t t_doc = r.t_documented() ? r.t_documented() : 0xFFFFu;
    MatAgent m;
    m.a = t_exp | (t_inf << 16);
    m.b = t_sev | (t_cri << 16);
    m.c = (uint32_t)r.t_end() | ((uint32_t)r.t_q_on() << 16);
    m.d = t_doc | ((uint32_t)(r.state() & 7) << 16) | ((uint32_t)local << 20);
    m.e = ((fl & F_TRX) && trc_of_agent) ? (__ldcg(trc_of_agent) & 0xFFFFu) : 0xFFFFu;  // first day of a traced Q after the end (:72-73)
    return m;
}
// packed state byte on day t (same function as state_on_day, on the packed thresholds)
__device__ __forceinline__ uint32_t mat_byte(const MatAgent &m, uint32_t t)
{
    const uint32_t cnt = (t >= (m.a & 0xFFFFu)) + (t >= (m.a >> 16)) + (t >= (m.b & 0xFFFFu)) + (t >= (m.b >> 16));
    const bool ended = t >= (m.c & 0xFFFFu);
    const uint32_t q = ((t >= (m.c >> 16)) && !ended) || t >= m.e;
    const uint32_t doc = t >= (m.d & 0xFFFFu);
    const uint32_t comp = ended ? ((m.d >> 16) & 7u) : cnt;
    return comp | (q << 3) | (doc << 4);
}

__global__ void __launch_bounds__(kMatThreads) seir_materialize_kernel(const __grid_constant__ DevCtx cx)
{
    __shared__ __align__(16) uint8_t s_tile[kMatWarps][512];
    __shared__ uint16_t s_list[kMatWarps][512];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t groups_per_rep = cx.n_pad / 512;
    const int64_t total = groups_per_rep * cx.n_rep;
    const int64_t row_stride = cx.n_pad / kVec;
    uint8_t *tile = s_tile[wib];
    uint16_t *list = s_list[wib];
    for (int64_t g = (int64_t)blockIdx.x * kMatWarps + wib; g < total; g += (int64_t)gridDim.x * kMatWarps) {
        const int rep = (int)(g / groups_per_rep);
        const int64_t vi = (g % groups_per_rep) * 32 + lane;  // my vector (16 agents) in the row
        uint8_t *hist = cx.hist + (size_t)rep * cx.T * cx.n_pad;
        const int Tr = __ldg(&cx.rp[rep].T);  // the replica's own horizon (rows beyond it are never read)
        const AgentRec *rec = cx.rec + (size_t)rep * cx.n_pad + (g % groups_per_rep) * 512;
        uint4 *row = reinterpret_cast<uint4 *>(hist) + vi;
        uint32_t m = __ldcs(cx.inbox + (size_t)rep * row_stride + vi) & 0x0F0F0F0Fu;  // "ever infected" bits
        // ---- warp-wide list of infected agents (local index 0..511)
        const int mine = __popc(m);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        const int n_inf = __shfl_sync(0xFFFFFFFFu, incl, 31);
        int at = incl - mine;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            list[at++] = (uint16_t)(lane * 16 + (bit & 7) * 4 + (bit >> 3));  // inverse of inbox_infected_bit
        }
        __syncwarp();
        const uint4 z = make_uint4(0, 0, 0, 0);
        if (n_inf == 0) {
            for (int t = 0; t < Tr; ++t) __stcs(row + (int64_t)t * row_stride, z);
            continue;
        }
        // ---- my share of the agents: entries lane, lane+32, ... (at most kMatPerLane in registers)
        MatAgent ag[kMatPerLane];
        int t_first = Tr;
#pragma unroll
        for (int k = 0; k < kMatPerLane; ++k) {
            const int e = lane + 32 * k;
            ag[k].a = 0xFFFFFFFFu; ag[k].b = 0xFFFFFFFFu; ag[k].c = 0xFFFFFFFFu; ag[k].d = 0xFFFFu; ag[k].e = 0xFFFFu;
            if (e < n_inf) {
                const int local = list[e];
                const AgentRec r = load_rec(rec + local);
                ag[k] = mat_pack(r, local, cx.trc ? cx.trc + (size_t)rep * cx.n_pad + (g % groups_per_rep) * 512 + local : nullptr);
                t_first = min(t_first, (int)r.t_exposed());
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t_first = min(t_first, __shfl_xor_sync(0xFFFFFFFFu, t_first, o));
        const int n_slots = min(kMatPerLane, (n_inf + 31) / 32);  // warp-uniform
        for (int t = 0; t < t_first; ++t) __stcs(row + (int64_t)t * row_stride, z);  // S everywhere
        for (int t = t_first; t < Tr; ++t) {
            reinterpret_cast<uint4 *>(tile)[lane] = z;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kMatPerLane; ++k) {
                if (k < n_slots) {
                    const uint32_t byte = mat_byte(ag[k], (uint32_t)t);
                    if (byte) tile[ag[k].d >> 20] = (uint8_t)byte;
                }
            }
            __syncwarp();
            __stcs(row + (int64_t)t * row_stride, reinterpret_cast<const uint4 *>(tile)[lane]);
            __syncwarp();
        }
        // ---- more than 32*kMatPerLane infected agents in 512: patch the rest byte-wise
        for (int e = 32 * kMatPerLane + lane; e < n_inf; e += 32) {
            const int local = list[e];
            const AgentRec r = load_rec(rec + local);
            const int64_t i = (g % groups_per_rep) * 512 + local;
            const int t_q2 = ((r.flags() & F_TRX) && cx.trc) ? (int)(__ldcg(cx.trc + (size_t)rep * cx.n_pad + i) & 0xFFFFu) : 0xFFFF;
            for (int t = r.t_exposed(); t < Tr; ++t) hist[(size_t)t * cx.n_pad + i] = state_on_day(r, t, t_q2);
        }
        __syncwarp();
    }
}

// The reference's [T][9][101] daily counts (:392 planes S,E,Mild,Documented,Severe,Critical,R,D,Q summed by age -- what the
// drivers' X.sum(axis=1) and age masks compute) from the day loop's per-day CHANGES: one thread per (plane, age) runs through
// the days.  Launched behind the day loop on the copy stream, next to the history materialisation; its int64 output goes to
// pinned host memory, so seir_get_tallies is one memcpy.  One CTA per replica.
__global__ void __launch_bounds__(1024) seir_tally_accumulate_kernel(const uint32_t *raw_all, long long *out_all, const long long *gs, int T)
{
    const int k = threadIdx.x;  // plane * kAges + age
    if (k >= SEIR_N_TALLY * kAges) return;
    const int plane = k / kAges, a = k % kAges;
    const uint32_t *raw = raw_all + (size_t)blockIdx.x * T * TAL_ROWS * kAges + a;
    long long *out = out_all + (size_t)blockIdx.x * T * SEIR_N_TALLY * kAges + k;
    int r0 = 0, r1 = -1, r2 = -1;
    long long base = 0, sign = 1;
    switch (plane) {
        case 0: r0 = TAL_NEW_E; base = gs[a]; sign = -1; break;          // S = group size - infections so far
        case 1: r0 = TAL_NEW_E; r1 = TAL_D_E; break;
        case 2: r0 = TAL_D_MILD; break;
        case 3: r0 = TAL_NEW_DOC; break;
        case 4: r0 = TAL_D_SEV; break;
        case 5: r0 = TAL_D_CRIT; break;
        case 6: r0 = TAL_NEW_R; break;
        case 7: r0 = TAL_NEW_D; break;
        default: r0 = TAL_NEW_QE; r1 = TAL_D_Q; r2 = TAL_NEW_Q2; break;
    }
    // (a missing row reads row r0 again with weight 0: every load is unconditional, and 16 days of them are in flight
    // together -- the kernel runs next to the materialisation kernel, which keeps the memory system busy, so it is the
    // number of dependent round trips that counts)
    const int w1 = r1 >= 0 ? 1 : 0, w2 = r2 >= 0 ? 1 : 0;
    const uint32_t *p0 = raw + r0 * kAges, *p1 = raw + (w1 ? r1 : r0) * kAges, *p2 = raw + (w2 ? r2 : r0) * kAges;
    long long acc = 0;
    int t = 0;
    for (; t + 16 <= T; t += 16) {
        int32_t v0[16], v1[16], v2[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const size_t o = (size_t)(t + u) * TAL_ROWS * kAges;
            v0[u] = (int32_t)__ldcg(p0 + o);  // (signed rows wrap around 2^32)
            v1[u] = (int32_t)__ldcg(p1 + o);
            v2[u] = (int32_t)__ldcg(p2 + o);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            acc += (long long)v0[u] + (long long)(w1 * v1[u]) + (long long)(w2 * v2[u]);
            out[(size_t)(t + u) * SEIR_N_TALLY * kAges] = base + sign * acc;
        }
    }
    for (; t < T; ++t) {
        const size_t o = (size_t)t * TAL_ROWS * kAges;
        acc += (long long)(int32_t)__ldcg(p0 + o) + (long long)(w1 * (int32_t)__ldcg(p1 + o)) + (long long)(w2 * (int32_t)__ldcg(p2 + o));
        out[(size_t)t * SEIR_N_TALLY * kAges] = base + sign * acc;
    }
}

__global__ void seir_pack_home_kernel(const uint8_t *home_u8, uint32_t *bitmap, int64_t n, int64_t n_words)
{
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t x = 0;
        for (int b = 0; b < 32; ++b) {
            const int64_t i = w * 32 + b;
            if (i < n && home_u8[i]) x |= 1u << b;
        }
        bitmap[w] = x;
    }
}

__device__ __forceinline__ double dec16(uint16_t v) { return v == NEVER ? __longlong_as_double(0x7FF0000000000000ll) : (double)v; }

__global__ void seir_agent_vector_kernel(DevCtx cx, int rep, int which, double *out)
{
    const uint32_t *inbox = cx.inbox + (size_t)rep * (cx.n_pad / 16);
    const AgentRec *rec = cx.rec + (size_t)rep * cx.n_pad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cx.n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool ever = (inbox[i >> 4] & inbox_infected_bit((int)(i & 15))) != 0;
        double v;
        if (!ever) {
            v = (which == SEIR_VEC_NUM_INFECTED_BY || which == SEIR_VEC_TIME_EXPOSED ||
                 which == SEIR_VEC_NUM_INFECTED_ASYMPT || which == SEIR_VEC_NUM_INFECTED_OUTSIDE) ? -1.0 : 0.0;
        } else {
            const AgentRec r = rec[i];
            switch (which) {
                case SEIR_VEC_NUM_INFECTED_BY: v = r.n_inf_by(); break;
                case SEIR_VEC_TIME_DOCUMENTED: {  // :76, :85: every tracer visit rewrites time_documented
                    uint32_t td = r.t_documented();
                    if ((r.flags() & F_TRX) && cx.trc) {
                        const uint32_t last = cx.trc[(size_t)rep * cx.n_pad + i] >> 16;
                        if (last > td) td = last;
                    }
                    v = td;
                    break;
                }
                case SEIR_VEC_TIME_TO_ACTIVATION: v = dec16(r.tt_activation()); break;
                case SEIR_VEC_TIME_TO_DEATH: v = dec16(r.tt_death()); break;
                case SEIR_VEC_TIME_TO_RECOVERY: v = dec16(r.tt_recovery()); break;
                case SEIR_VEC_TIME_CRITICAL: v = (r.flags() & F_CRIT) ? t_critical_of(r) : 0; break;
                case SEIR_VEC_TIME_EXPOSED: v = r.t_exposed(); break;
                case SEIR_VEC_NUM_INFECTED_ASYMPT: v = r.n_inf_asympt(); break;
                case SEIR_VEC_TIME_INFECTED: v = (r.flags() & F_MILD) ? t_infected_of(r) : 0; break;
                case SEIR_VEC_TIME_TO_SEVERE: v = dec16(r.tt_severe()); break;
                case SEIR_VEC_TIME_TO_ISOLATE: v = dec16(r.tt_isolate()); break;
                case SEIR_VEC_TIME_SEVERE: v = (r.flags() & F_SEV) ? t_severe_of(r) : 0; break;
                case SEIR_VEC_TIME_TO_CRITICAL: v = dec16(r.tt_critical()); break;
                default: v = r.n_inf_outside(); break;
            }
        }
        out[i] = v;
    }
}

// Exposure-day cohorts (the drivers' per-run statistics: run_simulation.py:448-453, parameter_sweep.py:572-600):
// for every day d and age a, out[d][0][a] = agents exposed on day d, [1] = sum of their num_infected_by,
// [2] = of them dead at the end (D[-1]), [3] = of them recovered at the end (R[-1]).
__global__ void seir_cohort_kernel(DevCtx cx, int rep, unsigned long long *out)
{
    const uint32_t *inbox = cx.inbox + (size_t)rep * (cx.n_pad / 16);
    const AgentRec *rec = cx.rec + (size_t)rep * cx.n_pad;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cx.n; i += (int64_t)gridDim.x * blockDim.x) {
        if (!(inbox[i >> 4] & inbox_infected_bit((int)(i & 15)))) continue;
        if (cx.n_shards > 1 && (i < cx.shard_lo[cx.my_shard] || i >= cx.shard_lo[cx.my_shard + 1])) continue;  // (its owner counts it)
        const AgentRec r = load_rec(rec + i);
        const int d = r.t_exposed(), age = r.stat() & 127, comp = r.state() & 7;
        if (d >= cx.T) continue;
        unsigned long long *o = out + (size_t)d * 4 * kAges + age;
        atomicAdd(o, 1ull);
        if (r.n_inf_by()) atomicAdd(o + kAges, (unsigned long long)r.n_inf_by());
        if (comp == C_D) atomicAdd(o + 2 * kAges, 1ull);
        if (comp == C_R) atomicAdd(o + 3 * kAges, 1ull);
    }
}

__global__ void seir_history_plane_kernel(const uint8_t *rows, int64_t n, int64_t n_pad, int n_rows, int which, uint8_t *out)
{
    const int64_t total = (int64_t)n_rows * n;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = g / n, i = g % n;
        const uint8_t s = rows[row * n_pad + i];
        const int comp = s & 7;
        uint8_t v;
        switch (which) {
            case SEIR_H_S: v = comp == C_S; break;
            case SEIR_H_E: v = comp == C_E; break;
            case SEIR_H_MILD: v = comp == C_MILD; break;
            case SEIR_H_DOCUMENTED: v = (s & ST_DOC) != 0; break;
            case SEIR_H_SEVERE: v = comp == C_SEV; break;
            case SEIR_H_CRITICAL: v = comp == C_CRIT; break;
            case SEIR_H_R: v = comp == C_R; break;
            case SEIR_H_D: v = comp == C_D; break;
            case SEIR_H_Q: v = (s & ST_Q) != 0; break;
            default: v = s & 0x1F; break;
        }
        out[g] = v;
    }
}

__global__ void seir_probe_kernel(uint64_t seed, int64_t m, double mean, double mu, double sigma, double enlam,
                                  double *exp_out, double *logn_out, long long *pois_out, double *log_out,
                                  const double *log_in)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        sd_block b = sd_draw(seed, (uint32_t)i, 0, SD_SITE_ACT, 0);
        exp_out[i] = sd_exp_days(sd_lane(b, 0), mean);
        logn_out[i] = sd_lognormal_days(seed, (uint32_t)i, 0, SD_SITE_INIT, 0, mu, sigma);
        pois_out[i] = sd_poisson(seed, (uint32_t)i, 0, SD_SITE_NCNT, 0, enlam, SD_MAX_CONTACTS_NATIVE);
        log_out[i] = sd_log(log_in[i]);
    }
}

// ------------------------------------------------------------------ host side
static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            char b_[512];                                                                              \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(b_);                                                                           \
        }                                                                                              \
    } while (0)

template <class Tp> struct DevBuf {
    Tp *p = nullptr;
    size_t count = 0;
    cudaError_t alloc(size_t c)
    {
        if (c <= count && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), (c ? c : 1) * sizeof(Tp));
        if (e == cudaSuccess) count = c;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; count = 0; }
};

struct seir_handle {
    int device = 0, sms = 0, n_rep = 0;
    int grid = 0, occ = 0;
    seir_params prm{};
    int64_t n = 0, n_pad = 0;
    int T_alloc = 0;
    bool replay_tables = false;
    std::vector<int64_t> group_sizes;  // N_age
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    cudaEvent_t ev3 = nullptr;
    float loop_ms = 0.f, total_ms = 0.f, mat_ms = 0.f;
    long long launches = 0, visits = 0;
    DevBuf<uint16_t> stat;
    DevBuf<int32_t> grp_off, grp_members, init_ids, alias_idx;
    DevBuf<int64_t> init_off;
    DevBuf<double> p_hh, tables, enlam, stage_d;
    DevBuf<uint8_t> hist, stage_u8, hist_prev;
    DevBuf<AgentRec> rec_prev;          // records / infected bits / tracer words of the run set aside by seir_history_preserve:
    DevBuf<uint32_t> inbox_prev, trc_prev;   // the per-agent vectors of that run stay readable (which | SEIR_VEC_PREVIOUS)
    bool prev_vectors = false, prev_trc = false;
    int prev_T = 0;   // rows of the history set aside by seir_history_preserve (0: none)
    DevBuf<uint32_t> inbox, tally, home, list_old, list_new, list_cnt;
    DevBuf<uint32_t> pk_pool;   // day buckets of the parked agents
    uint32_t run_serial = 0;    // runs of this handle (DevCtx.run_nonce)
    DevBuf<int> ctl;            // hand-over words of the single-replica day loop's launches
    DevBuf<RepParams> rp;       // per-replica overrides of the current run
    DevBuf<double> bank;        // table sets (seir_set_table_bank); set 0 = the handle's own tables
    DevBuf<int64_t> kcum;       // keyed prologue: prefix sums of the pick counts per replica
    int n_sets = 1;
    std::vector<RepParams> rp_host;
    std::vector<int64_t> n_init_run;   // initial cases per replica of the last run
    bool home_valid = false;           // the home bitmaps hold the last run's picks
    int history_mode = SEIR_HISTORY_EAGER;
    DevBuf<unsigned long long> winner, counters, day_ns, stage_ns, fine_ns;
    DevBuf<uint16_t> xday;
    DevBuf<AgentRec> rec;
    DevBuf<uint64_t> seeds;
    DevBuf<unsigned int> barrier;
    // sharded run: one population across the GPUs of a box
    int n_shards = 0, my_shard = 0;
    uint32_t shard_threshold = 393216;   // list entries of a day from which a sharded run leaves its replicated phase
    int64_t shard_lo[kMaxShards + 1] = {0};
    void *peer_ptr[kMaxShards][SEIR_SHARD_HANDLES] = {{nullptr}};   // opened IPC mappings: winner, inbox, list_new, list_cnt, flags, xday
    DevBuf<unsigned int> flags;
    unsigned int run_count = 0;
    unsigned int epoch_base = 0, epoch_next = 0;   // day-barrier epochs of the sharded runs: a running counter (monotone whatever T does)
    std::vector<uint8_t> age8;         // ages, for the per-shard group sizes
    std::vector<int64_t> shard_group_sizes;
    // contact tracing (allocated when the run can switch tracing on)
    bool trace_alloc = false;
    unsigned int trace_runs = 0;
    uint32_t log_cap = 0;
    DevBuf<uint4> log_ent;
    DevBuf<uint32_t> log_head, log_cnt, rootbits, rootsum, root_cnt, trc, tstack, rootsum2, treached, troots, treps, tcsz;
    DevBuf<unsigned long long> tcomp, tlead, ttotal;
    DevBuf<TraceCtl> tctl;
    // results every caller reads after a run (tallies, counters): copied to pinned host memory by the run itself, on a
    // second stream, while the history is being materialised
    float probe_ms = 0.0f;             // (SEIR_SORT_PROBE) sum of the day kernels' durations
    unsigned char *in_host = nullptr;  // pinned staging of a run's small inputs
    size_t in_host_bytes = 0;
    uint32_t *tally_host = nullptr;
    DevBuf<long long> tally64, gs_dev;     // the accumulated [T][9][101] counts of every replica (seir_tally_accumulate_kernel); group sizes
    long long *tally64_host = nullptr;     // ... and their pinned host copy, filled by the run itself
    unsigned long long *counters_host = nullptr;
    bool host_results = false;  // the pinned copies hold the last run's results
    bool raw_home = false;      // ... including the raw per-day tally changes (tally_host)
    cudaStream_t copy_stream = nullptr;
    uint32_t cluster_cap = 0;   // a day with at most this many list entries runs in the one-cluster kernel (0: never)
    bool cluster_failed = false;  // a cluster launch was refused on this device: not tried again
    DevBuf<uint32_t> tcta, tbig, tedge_off, tedge_cnt, tedge;
    uint32_t tedge_cap = 0;
    DevBuf<unsigned long long> tmark;
    double p_hh_scalar = 0.0;
    bool p_hh_uniform = true;
    DevCtx cx{};
};

static const size_t kTabPms = 0, kTabPsc = 404, kTabPcd = 808, kTabIso = 1212, kTabIsoA = 1313, kTabNatEn = 1414,
                    kTabAliasP = 1414 + 404, kTabTotal = 1414 + 404 + 101 * 101;

static int upload_tables(seir_handle *h, const seir_tables *tb)
{
    if (!tb || !tb->p_mild_severe || !tb->p_severe_critical || !tb->p_critical_death || !tb->iso_mean ||
        !tb->iso_asympt_mean || !tb->nat_enlam || !tb->nat_alias_prob || !tb->nat_alias_idx)
        return fail("seir_tables: a required table pointer is NULL");
    std::vector<double> host(kTabTotal);
    memcpy(&host[kTabPms], tb->p_mild_severe, 404 * sizeof(double));
    memcpy(&host[kTabPsc], tb->p_severe_critical, 404 * sizeof(double));
    memcpy(&host[kTabPcd], tb->p_critical_death, 404 * sizeof(double));
    memcpy(&host[kTabIso], tb->iso_mean, 101 * sizeof(double));
    memcpy(&host[kTabIsoA], tb->iso_asympt_mean, 101 * sizeof(double));
    memcpy(&host[kTabNatEn], tb->nat_enlam, 404 * sizeof(double));
    memcpy(&host[kTabAliasP], tb->nat_alias_prob, 101 * 101 * sizeof(double));
    for (int k = 0; k < 101 * 101; ++k)
        if (tb->nat_alias_idx[k] < 0 || tb->nat_alias_idx[k] > 100) return fail("seir_tables.nat_alias_idx out of range");
    CK(h->alias_idx.alloc(101 * 101));
    CK(cudaMemcpyAsync(h->alias_idx.p, tb->nat_alias_idx, 101 * 101 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CK(h->tables.alloc(kTabTotal));
    CK(cudaMemcpyAsync(h->tables.p, host.data(), kTabTotal * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->replay_tables = tb->enlam_pre && tb->enlam_post;
    if (h->replay_tables) {
        CK(h->enlam.alloc(2 * 101 * 101));
        CK(cudaMemcpyAsync(h->enlam.p, tb->enlam_pre, 101 * 101 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->enlam.p + 101 * 101, tb->enlam_post, 101 * 101 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// one table set of the bank (what a sweep's grid changes, see kBank* in seir_kernels.cuh)
static void pack_bank_set(const seir_tables *tb, double p_contact, double *dst)
{
    memcpy(dst + kBankPms, tb->p_mild_severe, 404 * sizeof(double));
    memcpy(dst + kBankPsc, tb->p_severe_critical, 404 * sizeof(double));
    memcpy(dst + kBankPcd, tb->p_critical_death, 404 * sizeof(double));
    memcpy(dst + kBankNatEn, tb->nat_enlam, 404 * sizeof(double));
    for (int k = kBankPcontact; k < kBankStride; ++k) dst[k] = p_contact;
}

static int upload_bank(seir_handle *h, int n_sets, const seir_tables *sets, const double *p_contact)
{
    std::vector<double> host((size_t)n_sets * kBankStride);
    for (int k = 0; k < n_sets; ++k) {
        const seir_tables *tb = sets + k;
        if (!tb->p_mild_severe || !tb->p_severe_critical || !tb->p_critical_death || !tb->nat_enlam)
            return fail("seir_set_table_bank: a set needs p_mild_severe, p_severe_critical, p_critical_death and nat_enlam");
        pack_bank_set(tb, p_contact[k], host.data() + (size_t)k * kBankStride);
    }
    CK(h->bank.alloc(host.size()));
    CK(cudaMemcpyAsync(h->bank.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_sets = n_sets;
    return 0;
}

static int check_params(const seir_params *p, int64_t n)
{
    if (!p) return fail("seir_params is NULL");
    if (p->n != n) return fail("seir_params.n does not match the population");
    if (p->T < 2 || p->T > 4000) return fail("seir_params.T must be in [2, 4000]");
    if (p->mode != SEIR_MODE_NATIVE && p->mode != SEIR_MODE_REPLAY) return fail("seir_params.mode is invalid");
    if (p->launch != SEIR_LAUNCH_PERSISTENT && p->launch != SEIR_LAUNCH_PER_DAY) return fail("seir_params.launch is invalid");
    if (p->tracing_enabled && p->t_tracing_start >= 1 && p->t_tracing_start < p->T) {
        if (!(p->p_trace_household >= 0.0) || !(p->p_trace_outside >= 0.0))
            return fail("p_trace_household / p_trace_outside must be >= 0");
        if (n >= (1ll << 31)) return fail("contact tracing supports populations below 2^31 agents");
    }
    return 0;
}

static bool tracing_can_start(const seir_params &p) { return p.tracing_enabled && p.t_tracing_start >= 1 && p.t_tracing_start < p.T; }

static int alloc_state(seir_handle *h)
{
    const size_t R = (size_t)h->n_rep, np = (size_t)h->n_pad, T = (size_t)h->prm.T;
    if (tracing_can_start(h->prm)) {
        // every outside infection is logged once per hitter (:387); same-day multiple hits make a few % more
        // entries than infections
        h->log_cap = (uint32_t)(np + np / 4 + 1024);
        CK(h->log_ent.alloc(R * h->log_cap));
        CK(h->log_head.alloc(R * np));
        CK(h->log_cnt.alloc(R));
        CK(h->rootbits.alloc(R * np / 32));
        CK(h->rootsum.alloc(R * np / 1024));
        CK(h->root_cnt.alloc(R * (T + 2)));
        CK(h->tmark.alloc(R * np));
        CK(cudaMemset(h->tmark.p, 0, R * np * 8));
        CK(h->trc.alloc(R * np));
        CK(h->tstack.alloc(R * np));
        CK(h->rootsum2.alloc(R * ((np + 32767) / 32768)));
        CK(h->tcomp.alloc(R * np));
        CK(h->tlead.alloc(R * np));
        CK(cudaMemset(h->tcomp.p, 0, R * np * 8));
        CK(cudaMemset(h->tlead.p, 0, R * np * 8));
        CK(h->treached.alloc(R * np));
        CK(h->troots.alloc(R * np));
        CK(h->treps.alloc(R * np));
        CK(h->tcsz.alloc(R * np));
        CK(h->tctl.alloc(R));
        CK(h->tcta.alloc(R * 2048));  // (grids of up to 2048 CTAs)
        CK(h->tbig.alloc(R * kBigCap));
        // the day's edge cache: room for (household size - 1 + log entries) of every agent the frontier reaches; a day
        // that needs more makes the searches evaluate the remaining agents' edges from the log again
        h->tedge_cap = (uint32_t)(2 * np + 1024);
        if (const char *dc = getenv("SEIR_DEBUG_TEDGE_CAP")) h->tedge_cap = (uint32_t)atoi(dc);  // (tests: exercise the fall-back)
        CK(h->tedge_off.alloc(R * np));
        CK(h->tedge_cnt.alloc(R * np));
        CK(h->tedge.alloc(R * h->tedge_cap));
        CK(h->ttotal.alloc(2));
        CK(cudaMemset(h->ttotal.p, 0, 16));
        h->trace_alloc = true;
    }
    if (h->history_mode == SEIR_HISTORY_EAGER) CK(h->hist.alloc(R * T * np));
    CK(h->tally.alloc(R * T * TAL_ROWS * kAges));
    CK(h->list_cnt.alloc(R * 3 * (T + 2)));
    CK(h->day_ns.alloc(T + 2));
    CK(cudaMemset(h->day_ns.p, 0, (T + 2) * 8));
    {
        const size_t tb = R * T * TAL_ROWS * kAges * 4;
        h->host_results = false;
        if (h->tally_host) { cudaFreeHost(h->tally_host); h->tally_host = nullptr; }  // (re-allocation for a longer T)
        if (h->counters_host) { cudaFreeHost(h->counters_host); h->counters_host = nullptr; }
        if (h->tally64_host) { cudaFreeHost(h->tally64_host); h->tally64_host = nullptr; }
        if (tb <= ((size_t)256 << 20)) {  // (beyond that the getters copy on demand)
            CK(cudaHostAlloc((void **)&h->tally_host, tb, cudaHostAllocDefault));
            CK(cudaHostAlloc((void **)&h->counters_host, R * 4 * 8, cudaHostAllocDefault));
            const size_t cells64 = (size_t)R * T * SEIR_N_TALLY * kAges;
            CK(h->tally64.alloc(cells64));
            CK(cudaHostAlloc((void **)&h->tally64_host, cells64 * 8, cudaHostAllocDefault));
        }
    }
    CK(h->fine_ns.alloc((T + 2) * 16));
    CK(cudaMemset(h->fine_ns.p, 0, (T + 2) * 16 * 8));
    CK(h->stage_ns.alloc((T + 2) * 8));
    CK(cudaMemset(h->stage_ns.p, 0, (T + 2) * 8 * 8));
    h->T_alloc = h->prm.T;
    return 0;
}

static void fill_ctx(seir_handle *h)
{
    DevCtx &c = h->cx;
    const seir_params &p = h->prm;
    c.n = h->n; c.n_pad = h->n_pad; c.n_rep = h->n_rep; c.T = p.T;
    int bpr = h->grid / h->n_rep;
    if (bpr < 1) bpr = 1;
    c.blocks_per_rep = bpr;
    c.list_old = h->list_old.p; c.list_new = h->list_new.p; c.list_cnt = h->list_cnt.p;
    c.pk_pool = h->pk_pool.p;
    c.ctl = h->ctl.p;
    c.rp = h->rp.p; c.bank = h->bank.p; c.home_rw = h->home.p; c.T_loop = p.T;
    c.stat = h->stat.p; c.grp_off = h->grp_off.p; c.grp_members = h->grp_members.p;
    c.p_hh_vec = h->p_hh_uniform ? nullptr : h->p_hh.p; c.p_hh_scalar = h->p_hh_scalar;
    c.p_ms = h->tables.p + kTabPms; c.p_sc = h->tables.p + kTabPsc; c.p_cd = h->tables.p + kTabPcd;
    c.iso_mean = h->tables.p + kTabIso; c.iso_asym = h->tables.p + kTabIsoA;
    c.nat_enlam = h->tables.p + kTabNatEn; c.nat_alias_prob = h->tables.p + kTabAliasP;
    c.nat_alias_idx = h->alias_idx.p;
    c.enlam = h->replay_tables ? h->enlam.p : nullptr;
    c.hist = h->hist.p; c.inbox = h->inbox.p; c.winner = h->winner.p; c.xday = h->xday.p; c.rec = h->rec.p; c.tally = h->tally.p;
    c.home = h->home.p; c.seeds = h->seeds.p; c.counters = h->counters.p; c.barrier = h->barrier.p;
    c.day_ns = h->day_ns.p; c.stage_ns = h->stage_ns.p; c.fine_ns = h->fine_ns.p;
    { const char *dp = getenv("SEIR_DEBUG_PASS"); c.dbg_pass = dp ? atoi(dp) : -1; }
    { const char *sr = getenv("SEIR_SPARSE_ROUNDS"); c.sparse_rounds = sr ? (uint32_t)atoi(sr) : 2u; }  // 0: staged passes only
    // (2: a day with up to two entries per warp is still cheaper as two warp-cooperative rounds, 14 us, than as a staged pass, 18 us:
    //  profiles/r02/experiments.md)
    { const char *sx = getenv("SEIR_SPARSE_MAX"); c.sparse_max = sx ? (uint32_t)atoi(sx) : 0xFFFFFFFFu; }
    {   // the one-cluster kernel takes the days with at most one entry per warp of its CTAs (SEIR_CLUSTER_CAP=0: never)
        const char *cc = getenv("SEIR_CLUSTER_CAP");
        h->cluster_cap = cc ? (uint32_t)atoi(cc) : (uint32_t)(kClusterCtas * (kSparseThreads / 32));
    }
    {   // single-replica runs: the sparse-day kernel (one CTA per SM) takes the days with at most one entry per warp and round
        const unsigned long long cap = (unsigned long long)c.sparse_rounds * (unsigned long long)(h->sms * (kSparseThreads / 32));
        c.sparse_cap = (uint32_t)(cap < c.sparse_max ? cap : c.sparse_max);
        if (h->n_shards > 1 && c.sparse_cap > h->shard_threshold) c.sparse_cap = h->shard_threshold;  // (a sharded run's replicated phase ends there)
    }
    c.n_shards = h->n_shards; c.my_shard = h->my_shard; c.shard_threshold = h->shard_threshold;
    for (int g = 0; g <= kMaxShards; ++g) c.shard_lo[g] = (uint32_t)h->shard_lo[g < h->n_shards ? g : (h->n_shards > 0 ? h->n_shards : 0)];
    for (int g = 0; g < kMaxShards; ++g) {
        const bool mine = g == h->my_shard || h->n_shards <= 1 || g >= h->n_shards;
        c.peer_winner[g] = mine ? h->winner.p : (unsigned long long *)h->peer_ptr[g][0];
        c.peer_inbox[g] = mine ? h->inbox.p : (uint32_t *)h->peer_ptr[g][1];
        c.peer_list_new[g] = mine ? h->list_new.p : (uint32_t *)h->peer_ptr[g][2];
        c.peer_list_cnt[g] = mine ? h->list_cnt.p : (uint32_t *)h->peer_ptr[g][3];
        c.peer_flags[g] = mine ? h->flags.p : (unsigned int *)h->peer_ptr[g][4];
        c.peer_xday[g] = mine ? h->xday.p : (uint16_t *)h->peer_ptr[g][5];
    }
    c.epoch_base = h->epoch_base;
    const bool tr = tracing_can_start(p) && h->trace_alloc;
    c.trace_from = tr ? p.t_tracing_start : 0;
    c.p_trace_hh = p.p_trace_household; c.p_trace_out = p.p_trace_outside;
    c.log_ent = tr ? h->log_ent.p : nullptr; c.log_head = tr ? h->log_head.p : nullptr; c.log_cnt = tr ? h->log_cnt.p : nullptr;
    c.log_cap = tr ? h->log_cap : 0u;
    c.rootbits = tr ? h->rootbits.p : nullptr; c.rootsum = tr ? h->rootsum.p : nullptr; c.root_cnt = tr ? h->root_cnt.p : nullptr;
    c.tmark = tr ? h->tmark.p : nullptr; c.trc = tr ? h->trc.p : nullptr; c.tstack = tr ? h->tstack.p : nullptr;
    c.rootsum2 = tr ? h->rootsum2.p : nullptr;
    const bool par = tr && p.launch == SEIR_LAUNCH_PERSISTENT && h->n_rep <= kMaxRepsPerCta * (h->grid / (h->grid / h->n_rep > 0 ? h->grid / h->n_rep : 1));
    c.tcomp = par ? h->tcomp.p : nullptr; c.tlead = par ? h->tlead.p : nullptr;
    c.treached = par ? h->treached.p : nullptr; c.troots = par ? h->troots.p : nullptr;
    c.treps = par ? h->treps.p : nullptr; c.tcsz = par ? h->tcsz.p : nullptr;
    c.tctl = par ? h->tctl.p : nullptr; c.ttotal = par ? h->ttotal.p : nullptr;
    c.tcta = par ? h->tcta.p : nullptr; c.tbig = par ? h->tbig.p : nullptr;
    c.tedge_off = par ? h->tedge_off.p : nullptr; c.tedge_cnt = par ? h->tedge_cnt.p : nullptr;
    c.tedge = par ? h->tedge.p : nullptr; c.tedge_cap = par ? h->tedge_cap : 0u;
    c.t_lockdown = p.t_lockdown; c.t_stayhome = p.t_stayhome_start; c.mode = p.mode;
    c.p_doc = p.p_documented_in_mild; c.p_contact = p.p_infect_given_contact; c.asym = p.asymptomatic_transmissibility;
    c.m_sev = p.mean_time_to_severe; c.m_mild_rec = p.mean_time_mild_recovery; c.m_crit = p.mean_time_to_critical;
    c.m_sev_rec = p.mean_time_severe_recovery; c.m_crit_rec = p.mean_time_critical_recovery; c.m_death = p.mean_time_to_death;
    c.act_mu = p.time_to_activation_mean; c.act_sigma = p.time_to_activation_std;
}

extern "C" {

const char *seir_last_error(void) { return g_err.c_str(); }

}  // extern "C"
#include "seir_prologue_host.cuh"   // seir_prologue_mt: the reference's MT19937 picks of one seed, host only
#include "seir_tsv_host.cuh"        // seir_format_tsv: the drivers' per-job TSV files, host only
extern "C" {

int seir_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

const char *seir_build_info(void)
{
    return "libseir_b200: CUDA kernels for sm_100a (compute_100a), nvcc " SEIR_STR(__CUDACC_VER_MAJOR__) "."
           SEIR_STR(__CUDACC_VER_MINOR__) ", -lineinfo, event-driven active lists, " SEIR_STR(SEIR_THREADS) " threads/CTA";
}

void seir_abi_sizes(int64_t out5[5])
{
    out5[0] = sizeof(seir_params); out5[1] = sizeof(seir_population); out5[2] = sizeof(seir_tables);
    out5[3] = sizeof(seir_replica); out5[4] = sizeof(seir_overrides);
}

int seir_create(const seir_params *params, const seir_population *pop, const seir_tables *tables, int n_replicas,
                int device, seir_handle **out)
{
    if (!out) return fail("out handle pointer is NULL");
    *out = nullptr;
    if (!pop || !pop->households || !pop->age || !pop->diabetes || !pop->hypertension || !pop->group_offsets ||
        !pop->group_members || !pop->p_infect_household)
        return fail("seir_population: a required pointer is NULL");
    if (n_replicas < 1 || n_replicas > 4096) return fail("n_replicas must be in [1, 4096]");
    const int64_t n = pop->n;
    if (n < 1 || n > 2000000000ll) return fail("population size must be in [1, 2e9]");
    if (pop->W < 1 || pop->W > 7) return fail("households.shape[1] must be in [1, 7]");
    if (check_params(params, n)) return 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail("no CUDA device: libseir_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("libseir_b200 is built for sm_100a (B200); this device is older");

    seir_handle *h = new seir_handle();
    h->device = device; h->sms = prop.multiProcessorCount; h->n_rep = n_replicas; h->prm = *params; h->n = n;
    h->history_mode = params->history_mode == SEIR_HISTORY_ON_DEMAND ? SEIR_HISTORY_ON_DEMAND : SEIR_HISTORY_EAGER;
    h->n_pad = (n + kPad - 1) / kPad * kPad;
    int rc = 1;
    do {
        // ---- static population: verify the contiguous all-pairs household layout, pack 2 B/agent
        const int W = pop->W;
        std::vector<uint16_t> stat((size_t)h->n_pad, 0);
        bool ok = true;
        std::string why;
        for (int64_t i = 0; i < n && ok;) {
            // household of i: i is its first member (pos 0); size = 1 + number of row entries
            int sz = 1;
            const int64_t *row = pop->households + i * W;
            while (sz - 1 < W && row[sz - 1] != -1) ++sz;
            if (i + sz > n) { ok = false; why = "household runs past n"; break; }
            for (int pos = 0; pos < sz && ok; ++pos) {
                const int64_t a = i + pos;
                const int64_t *r = pop->households + a * W;
                for (int j = 0; j < W; ++j) {
                    const int64_t expect = j < sz - 1 ? i + j + (j >= pos ? 1 : 0) : -1;
                    if (r[j] != expect) { ok = false; why = "row is not the contiguous all-pairs layout"; break; }
                }
                const int64_t ag = pop->age[a], db = pop->diabetes[a], hy = pop->hypertension[a];
                if (ag < 0 || ag >= kAges || (db | hy) < 0 || db > 1 || hy > 1) { ok = false; why = "age/comorbidity out of range"; break; }
                stat[(size_t)a] = pack_static((int)ag, (int)db, (int)hy, pos, sz);
            }
            i += sz;
        }
        if (!ok) { fail("seir_population.households: " + why + " (sample_households.py:344-349 emits contiguous households)"); break; }
        h->age8.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) h->age8[(size_t)i] = (uint8_t)pop->age[i];
        std::vector<int32_t> goff(kAges + 1), gmem((size_t)n);
        h->group_sizes.resize(kAges);
        bool gok = pop->group_offsets[0] == 0 && pop->group_offsets[kAges] == n;
        for (int a = 0; a < kAges && gok; ++a) {
            goff[a] = (int32_t)pop->group_offsets[a];
            h->group_sizes[a] = pop->group_offsets[a + 1] - pop->group_offsets[a];
            if (h->group_sizes[a] < 0) gok = false;
        }
        goff[kAges] = (int32_t)n;
        for (int a = 0; a < kAges && gok; ++a)
            for (int64_t k = pop->group_offsets[a]; k < pop->group_offsets[a + 1]; ++k) {
                const int64_t m = pop->group_members[k];
                if (m < 0 || m >= n || pop->age[m] != a) { gok = false; break; }
                gmem[(size_t)k] = (int32_t)m;
            }
        if (!gok) { fail("seir_population.group_offsets/group_members do not partition the agents by age"); break; }
        {   // group sizes for the device-side accumulation of the tallies (S = group size - infections so far)
            long long gs64[kAges];
            for (int a = 0; a < kAges; ++a) gs64[a] = (long long)h->group_sizes[a];
            if (h->gs_dev.alloc(kAges) != cudaSuccess ||
                cudaMemcpy(h->gs_dev.p, gs64, sizeof gs64, cudaMemcpyHostToDevice) != cudaSuccess) { fail("group sizes: device allocation failed"); break; }
        }
        h->p_hh_scalar = pop->p_infect_household[0];
        h->p_hh_uniform = true;
        for (int64_t i = 1; i < n; ++i)
            if (pop->p_infect_household[i] != h->p_hh_scalar) { h->p_hh_uniform = false; break; }

        cudaError_t e;
#define CKB(call) if ((e = (call)) != cudaSuccess) { fail(std::string(#call " failed: ") + cudaGetErrorString(e)); break; }
        CKB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CKB(cudaEventCreate(&h->ev0)); CKB(cudaEventCreate(&h->ev1)); CKB(cudaEventCreate(&h->ev2));
        CKB(cudaEventCreate(&h->ev3));
        {
            int least = 0, greatest = 0;
            CKB(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CKB(cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, greatest));
        }
        CKB(h->stat.alloc((size_t)h->n_pad));
        CKB(cudaMemcpy(h->stat.p, stat.data(), (size_t)h->n_pad * 2, cudaMemcpyHostToDevice));
        CKB(h->grp_off.alloc(kAges + 1));
        CKB(cudaMemcpy(h->grp_off.p, goff.data(), (kAges + 1) * 4, cudaMemcpyHostToDevice));
        CKB(h->grp_members.alloc((size_t)n));
        CKB(cudaMemcpy(h->grp_members.p, gmem.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
        if (!h->p_hh_uniform) {
            CKB(h->p_hh.alloc((size_t)n));
            CKB(cudaMemcpy(h->p_hh.p, pop->p_infect_household, (size_t)n * 8, cudaMemcpyHostToDevice));
        }
        const size_t R = (size_t)n_replicas, np = (size_t)h->n_pad;
        CKB(h->inbox.alloc(R * np / 16));
        CKB(h->winner.alloc(R * np));
        CKB(h->rec.alloc(R * np));
        CKB(h->list_old.alloc(R * 2 * np));
        CKB(h->list_new.alloc(R * 2 * np));
        CKB(h->pk_pool.alloc(R * kParkRing * np));  // day buckets of the parked agents (64 B per agent and replica)
        CKB(h->home.alloc(R * np / 32));
        CKB(cudaMemset(h->home.p, 0, R * np / 32 * 4));
        CKB(h->seeds.alloc(R));
        CKB(h->counters.alloc(R * 4));
        CKB(h->ctl.alloc(kMaxLaunches + 1));
        CKB(h->barrier.alloc(kBarStride * (2 + 64)));  // flat word, global word, up to 64 group counters (1024 CTAs)
        CKB(h->init_off.alloc(R + 1));
        CKB(h->rp.alloc(R));
        CKB(h->kcum.alloc(R * (kAges + 2)));
        CKB(h->flags.alloc(kMaxShards + 2));
        CKB(cudaMemset(h->flags.p, 0, (kMaxShards + 2) * 4));
#undef CKB
        if (alloc_state(h)) break;
        if (upload_tables(h, tables)) break;
        if (upload_bank(h, 1, tables, &params->p_infect_given_contact)) break;
        // persistent grid: every CTA must be co-resident
        int occ = 0;
        if (cudaFuncSetAttribute(seir_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(seir_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(seir_dense_solo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(seir_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(seir_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(seir_day_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)) != cudaSuccess ||
            cudaFuncSetAttribute(seir_day_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem)) != cudaSuccess) {
            fail("cannot reserve shared memory for the day-step kernels");
            break;
        }
        int occ_t = 0, occ_s = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, seir_dense_solo_kernel, kThreads, sizeof(PassSmem)) != cudaSuccess || occ_s < 1 ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seir_persistent_kernel<false>, kThreads, sizeof(PassSmem)) != cudaSuccess || occ < 1 ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, seir_persistent_kernel<true>, kThreads, sizeof(PassSmem)) != cudaSuccess || occ_t < 1) {
            fail("occupancy query failed for seir_persistent_kernel");
            break;
        }
        if (occ_t < occ) occ = occ_t;  // one grid size for all instantiations
        if (occ_s < occ) occ = occ_s;
        h->occ = occ;
        h->grid = h->sms * occ;
        fill_ctx(h);
        rc = 0;
    } while (0);
    if (rc) { seir_destroy(h); return 1; }
    *out = h;
    return 0;
}

int seir_shard_export(seir_handle *h, uint8_t *handles_out)
{
    if (!h || !handles_out) return fail("seir_shard_export: NULL argument");
    CK(cudaSetDevice(h->device));
    CK(h->xday.alloc((size_t)h->n_rep * h->n_pad));  // (only a sharded run keeps the infection days: 2 B per agent)
    void *ptrs[SEIR_SHARD_HANDLES] = {h->winner.p, h->inbox.p, h->list_new.p, h->list_cnt.p, h->flags.p, h->xday.p};
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
    for (int k = 0; k < SEIR_SHARD_HANDLES; ++k) {
        cudaIpcMemHandle_t mh;
        CK(cudaIpcGetMemHandle(&mh, ptrs[k]));
        memcpy(handles_out + 64 * k, &mh, 64);
    }
    return 0;
}

int seir_shard_attach(seir_handle *h, int rank, int world, const uint8_t *all_handles, const int64_t *bounds)
{
    if (!h || !all_handles || !bounds) return fail("seir_shard_attach: NULL argument");
    if (world < 2 || world > kMaxShards || rank < 0 || rank >= world) return fail("seir_shard_attach: need 2 <= world <= 8 and 0 <= rank < world");
    if (h->n_rep != 1) return fail("a sharded run takes one replica");
    if (h->n_shards > 1) return fail("seir_shard_attach: already attached");
    if (bounds[0] != 0 || bounds[world] != h->n) return fail("shard bounds must run from 0 to n");
    CK(cudaSetDevice(h->device));
    for (int g = 0; g < world; ++g) {
        if (bounds[g + 1] <= bounds[g]) return fail("shard bounds must be increasing");
        // a cut must fall on a household boundary: household transmission never leaves the shard
        if (g > 0) {
            std::vector<uint16_t> st(1);
            CK(cudaMemcpy(st.data(), h->stat.p + bounds[g], 2, cudaMemcpyDeviceToHost));
            if ((st[0] >> 9) & 7) return fail("a shard bound cuts a household (bounds must be household starts)");
        }
    }
    for (int g = 0; g < world; ++g) {
        if (g == rank) continue;
        for (int k = 0; k < SEIR_SHARD_HANDLES; ++k) {
            cudaIpcMemHandle_t mh;
            memcpy(&mh, all_handles + ((size_t)g * SEIR_SHARD_HANDLES + k) * 64, 64);
            CK(cudaIpcOpenMemHandle(&h->peer_ptr[g][k], mh, cudaIpcMemLazyEnablePeerAccess));
        }
    }
    h->n_shards = world; h->my_shard = rank;
    for (int g = 0; g <= world; ++g) h->shard_lo[g] = bounds[g];
    h->shard_group_sizes.assign(kAges, 0);
    for (int64_t i = bounds[rank]; i < bounds[rank + 1]; ++i) h->shard_group_sizes[h->age8[(size_t)i]] += 1;
    {   // (a shard counts its own agents: the tallies of the ranks add up to the population's)
        long long gs64[kAges];
        for (int a = 0; a < kAges; ++a) gs64[a] = (long long)h->shard_group_sizes[(size_t)a];
        CK(h->gs_dev.alloc(kAges));
        CK(cudaMemcpy(h->gs_dev.p, gs64, sizeof gs64, cudaMemcpyHostToDevice));
    }
    fill_ctx(h);
    return 0;
}

int seir_shard_set_threshold(seir_handle *h, int64_t entries)
{
    if (!h) return fail("handle is NULL");
    if (entries < 0 || entries > 0xFFFFFFFFll) return fail("seir_shard_set_threshold: entries must be in [0, 2^32)");
    h->shard_threshold = (uint32_t)entries;
    fill_ctx(h);
    return 0;
}

int seir_set_params(seir_handle *h, const seir_params *params, const seir_tables *tables)
{
    if (!h) return fail("handle is NULL");
    CK(cudaSetDevice(h->device));
    if (params) {
        if (check_params(params, h->n)) return 1;
        h->prm = *params;
        if ((params->T > h->T_alloc || (tracing_can_start(*params) && !h->trace_alloc)) && alloc_state(h)) return 1;
    }
    if (tables && upload_tables(h, tables)) return 1;
    if (tables && upload_bank(h, 1, tables, &h->prm.p_infect_given_contact)) return 1;  // (a new parameter set: back to one table set)
    fill_ctx(h);
    return 0;
}

int seir_set_table_bank(seir_handle *h, int n_sets, const seir_tables *sets, const double *p_infect_given_contact)
{
    if (!h || !sets || !p_infect_given_contact) return fail("seir_set_table_bank: NULL argument");
    if (n_sets < 1 || n_sets > 4096) return fail("seir_set_table_bank: n_sets must be in [1, 4096]");
    CK(cudaSetDevice(h->device));
    if (upload_bank(h, n_sets, sets, p_infect_given_contact)) return 1;
    fill_ctx(h);
    return 0;
}

int seir_run(seir_handle *h, const seir_replica *reps) { return seir_run_ex(h, reps, nullptr); }

int seir_run_ex(seir_handle *h, const seir_replica *reps, const seir_overrides *ovr)
{
    if (!h || !reps) return fail("seir_run: NULL argument");
    CK(cudaSetDevice(h->device));
    if (h->prm.mode == SEIR_MODE_REPLAY && !h->replay_tables) return fail("replay mode needs enlam_pre/enlam_post tables");
    if (h->n_shards > 1) {
        if (h->n_rep != 1) return fail("a sharded run takes one replica");
        if (h->prm.launch != SEIR_LAUNCH_PERSISTENT) return fail("a sharded run needs the persistent launch (the GPUs meet at a device-side barrier every day)");
        if (tracing_can_start(h->prm)) return fail("contact tracing is not available in sharded runs (pass t_tracing_start >= T or tracing_enabled = 0)");
        if (ovr) return fail("per-replica overrides are not available in sharded runs");
    }
    const int R = h->n_rep;
    const int64_t n = h->n;
    // ---- per-replica overrides: day numbers and table set (defaults: the handle's parameters)
    h->rp_host.assign((size_t)R, RepParams{});
    int T_loop = 2, t_home_min = 0x7FFFFFFF, trace_min = 0;
    for (int r = 0; r < R; ++r) {
        RepParams &q = h->rp_host[(size_t)r];
        q.T = ovr ? ovr[r].T : h->prm.T;
        q.t_lockdown = ovr ? ovr[r].t_lockdown : h->prm.t_lockdown;
        q.t_stayhome = ovr ? ovr[r].t_stayhome_start : h->prm.t_stayhome_start;
        const int tts = ovr ? ovr[r].t_tracing_start : h->prm.t_tracing_start;
        q.trace_from = (h->prm.tracing_enabled && tts >= 1 && tts < q.T && h->trace_alloc) ? tts : 0;
        q.bank = ovr ? ovr[r].table_set : 0;
        if (q.T < 2 || q.T > h->prm.T) return fail("seir_overrides.T must be in [2, the handle's T]");
        if (q.bank < 0 || q.bank >= h->n_sets) return fail("seir_overrides.table_set is not a registered set (seir_set_table_bank)");
        if (ovr && h->prm.tracing_enabled && tts >= 1 && tts < q.T && !h->trace_alloc)
            return fail("seir_overrides.t_tracing_start switches tracing on, but the handle was created without tracing state "
                        "(create it with a t_tracing_start below T)");
        if (q.T > T_loop) T_loop = q.T;
        if (q.t_stayhome >= 1 && q.t_stayhome < q.T && q.t_stayhome < t_home_min) t_home_min = q.t_stayhome;
        if (q.trace_from && (!trace_min || q.trace_from < trace_min)) trace_min = q.trace_from;
    }
    if (h->n_shards > 1) {  // every rank runs the same sequence of seir_run calls: the barrier epochs stay in step
        ++h->run_count;
        h->epoch_base = h->epoch_next;
        h->epoch_next += (unsigned int)(h->prm.T + 2);
        CK(cudaMemsetAsync(h->flags.p + kMaxShards + 1, 0, sizeof(unsigned int), h->stream));  // (the error word of an earlier time-out)
    }
    // ---- per-run host inputs -> device
    // the run's small inputs (seeds, offsets, the ids of the initial cases as int32, the pick counts of the keyed
    // prologue, the overrides) are assembled in ONE pinned host buffer and copied from there: truly asynchronous
    // copies, no pageable staging inside the driver
    int64_t n_ids = 0;
    bool any_keyed = false;
    h->n_init_run.assign((size_t)R, 0);
    for (int r = 0; r < R; ++r) {
        if (reps[r].keyed_prologue) {
            if (!(reps[r].initial_infected_fraction >= 0.0) || reps[r].initial_infected_fraction > 1.0)
                return fail("seir_replica.initial_infected_fraction must be in [0, 1]");
            h->n_init_run[(size_t)r] = (int64_t)(reps[r].initial_infected_fraction * (double)n);  // int(frac * n) (:186)
            any_keyed = true;
        } else {
            if (reps[r].n_initial < 0 || (reps[r].n_initial > 0 && !reps[r].initial_infected)) return fail("seir_replica.initial_infected is invalid");
            h->n_init_run[(size_t)r] = reps[r].n_initial;
        }
        n_ids += h->n_init_run[(size_t)r];
    }
    const size_t kc = (size_t)R * (kAges + 2);
    const size_t in_bytes = (size_t)R * 8 + (size_t)(R + 1) * 8 + kc * 8 + (size_t)R * sizeof(RepParams) + (size_t)n_ids * 4;
    if (in_bytes > h->in_host_bytes) {
        if (h->in_host) cudaFreeHost(h->in_host);
        h->in_host = nullptr; h->in_host_bytes = 0;
        CK(cudaHostAlloc((void **)&h->in_host, in_bytes + 4096, cudaHostAllocDefault));
        h->in_host_bytes = in_bytes + 4096;
    }
    unsigned char *cur = h->in_host;
    uint64_t *seeds = reinterpret_cast<uint64_t *>(cur); cur += (size_t)R * 8;
    int64_t *off = reinterpret_cast<int64_t *>(cur); cur += (size_t)(R + 1) * 8;
    int64_t *kcum = reinterpret_cast<int64_t *>(cur); cur += kc * 8;
    RepParams *rp = reinterpret_cast<RepParams *>(cur); cur += (size_t)R * sizeof(RepParams);
    int32_t *ids = reinterpret_cast<int32_t *>(cur);
    const bool home_used = t_home_min != 0x7FFFFFFF;  // Home = Home_real only from t_stayhome_start on (:243-244): a run that never gets there never reads the bitmap
    off[0] = 0;
    for (int r = 0; r < R; ++r) {
        seeds[r] = reps[r].seed;
        rp[r] = h->rp_host[(size_t)r];
        off[r + 1] = off[r] + h->n_init_run[(size_t)r];
        int64_t *kc_r = kcum + (size_t)r * (kAges + 2);
        kc_r[0] = 0;
        if (reps[r].keyed_prologue) {
            for (int a = 0; a < kAges; ++a) {
                const double f = reps[r].fraction_stay_home ? reps[r].fraction_stay_home[a] : 0.0;
                if (!(f >= 0.0) || f > 1.0) return fail("seir_replica.fraction_stay_home must be in [0, 1]");
                kc_r[a + 1] = kc_r[a] + (home_used ? (int64_t)(f * (double)h->group_sizes[(size_t)a]) : 0);  // int(frac * len(matches)) (:181)
            }
            kc_r[kAges + 1] = kc_r[kAges] + h->n_init_run[(size_t)r];
        } else {
            for (int a = 0; a <= kAges; ++a) kc_r[a + 1] = 0;
            for (int64_t k = 0; k < reps[r].n_initial; ++k) {
                const int64_t a = reps[r].initial_infected[k];
                if (a < 0 || a >= n) return fail("seir_replica.initial_infected holds an id outside [0, n)");
                ids[(size_t)(off[r] + k)] = (int32_t)a;
            }
        }
    }
    CK(h->init_ids.alloc((size_t)n_ids));
    CK(cudaMemcpyAsync(h->seeds.p, seeds, R * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->init_off.p, off, (R + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->rp.p, rp, (size_t)R * sizeof(RepParams), cudaMemcpyHostToDevice, h->stream));
    if (any_keyed) CK(cudaMemcpyAsync(h->kcum.p, kcum, kc * 8, cudaMemcpyHostToDevice, h->stream));
    if (n_ids) CK(cudaMemcpyAsync(h->init_ids.p, ids, (size_t)n_ids * 4, cudaMemcpyHostToDevice, h->stream));  // (the keyed replicas' slots are overwritten by the device)
    const int64_t n_words = h->n_pad / 32;
    for (int r = 0; r < R && home_used; ++r) {
        uint32_t *bm = h->home.p + (size_t)r * n_words;
        if (!reps[r].keyed_prologue && reps[r].home_real) {
            CK(h->stage_u8.alloc((size_t)n));
            CK(cudaMemcpyAsync(h->stage_u8.p, reps[r].home_real, (size_t)n, cudaMemcpyHostToDevice, h->stream));
            // (the one staging buffer is reused by the next replica: stream order keeps its copy behind this kernel;
            // the caller's buffers are borrowed until seir_run returns, and it synchronises before it does)
            seir_pack_home_kernel<<<h->sms * 4, 256, 0, h->stream>>>(h->stage_u8.p, bm, n, n_words);
        } else {
            CK(cudaMemsetAsync(bm, 0, (size_t)n_words * 4, h->stream));
        }
    }
    h->home_valid = home_used;
    fill_ctx(h);
    h->cx.T_loop = T_loop;
    {   // a nonce no other run of this process has used (record memory is recycled between handles, and it is not cleared)
        static std::atomic<uint32_t> g_serial{0};
        h->run_serial = (g_serial.fetch_add(1u) + 1u) & 0x7FFFFFFFu;
        h->cx.run_nonce = h->run_serial | 0x80000000u;
    }
    h->cx.trace_from = trace_min;  // the first day on which ANY replica traces (each replica records roots from its own day on)
    h->launches = 0;
    // ---- init + day loop, timed on the launching stream
    CK(cudaEventRecord(h->ev0, h->stream));
    const bool tracing = h->cx.trace_from >= 1;
    if (tracing) {
        CK(cudaMemsetAsync(h->log_cnt.p, 0, (size_t)R * 4, h->stream));
        CK(cudaMemsetAsync(h->root_cnt.p, 0, (size_t)R * (h->prm.T + 2) * 4, h->stream));
        CK(cudaMemsetAsync(h->rootbits.p, 0, (size_t)R * h->n_pad / 32 * 4, h->stream));
        CK(cudaMemsetAsync(h->rootsum.p, 0, (size_t)R * h->n_pad / 1024 * 4, h->stream));
        CK(cudaMemsetAsync(h->rootsum2.p, 0, (size_t)R * ((h->n_pad + 32767) / 32768) * 4, h->stream));
        if ((++h->trace_runs & 0xFFFFFu) == 0u) {  // the stamps wrap after 2^20 runs: start over with clean marks
            CK(cudaMemsetAsync(h->tmark.p, 0, (size_t)R * h->n_pad * 8, h->stream));
            CK(cudaMemsetAsync(h->tcomp.p, 0, (size_t)R * h->n_pad * 8, h->stream));
            CK(cudaMemsetAsync(h->tlead.p, 0, (size_t)R * h->n_pad * 8, h->stream));
            ++h->trace_runs;
        }
        h->cx.tday0 = (h->trace_runs & 0xFFFFFu) << 12;
    }
    seir_init_state_kernel<<<h->sms * 8, 256, 0, h->stream>>>(h->cx, h->barrier.p, (int)(kBarStride * (2 + 64)));
    ++h->launches;
    if (any_keyed) {  // the prologue's picks (:176-187), made on the device
        dim3 g(h->sms * 2 > 64 ? 64 : h->sms * 2, R);
        seir_prologue_kernel<<<g, 256, 0, h->stream>>>(h->cx, h->kcum.p, h->init_ids.p, h->init_off.p, home_used ? 1 : 0);
        ++h->launches;
    }
    if (n_ids) {
        dim3 g(32, R);
        seir_init_infected_kernel<<<g, 128, 0, h->stream>>>(h->cx, h->init_ids.p, h->init_off.p);
        ++h->launches;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    if (h->prm.launch == SEIR_LAUNCH_PERSISTENT) {
        void *args[] = {&h->cx};
        // single-replica runs without tracing: sparse-day and dense-day kernels in turn (see seir_sparse_kernel); the last
        // dense launch never hands a day back, so four launches finish the run, and one that finds nothing left to do exits
        // at once.  A sharded run's REPLICATED phase is such a run on every GPU; its dense launches also stop at the first
        // day whose lists exceed the shard threshold, and the generic kernel (sharded passes, cross-GPU barrier) goes on.
        const bool solo = !tracing && R == 1 && h->prm.mode == SEIR_MODE_NATIVE && SEIR_FLAT_BARRIER && !getenv("SEIR_NO_SOLO") &&
                          !(h->n_shards > 1 && h->shard_threshold == 0u);
        h->cx.start_word = -1;
        if (solo) {
            // launch 0: the first days in ONE thread-block cluster (seir_cluster_kernel); then sparse, dense, sparse, dense
            {
                int k0 = 0;
                uint32_t cap = h->cluster_cap < h->cx.sparse_cap ? h->cluster_cap : h->cx.sparse_cap;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(kClusterCtas); cfg.blockDim = dim3(kSparseThreads);
                cfg.dynamicSmemBytes = sizeof(PassSmem); cfg.stream = h->stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = kClusterCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                if (cap == 0u || h->cluster_failed) { cap = 0u; cfg.numAttrs = 0; cfg.gridDim = dim3(1); }  // (switched off: a one-CTA launch that hands day 1 on)
                cudaError_t ce = cudaLaunchKernelEx(&cfg, seir_cluster_kernel, h->cx, k0, cap);
                if (ce != cudaSuccess && cfg.numAttrs) {  // no room for the cluster on this device (partitioned, busy): the 148-CTA kernels take day 1 on
                    (void)cudaGetLastError();
                    h->cluster_failed = true;
                    cap = 0u; cfg.numAttrs = 0; cfg.gridDim = dim3(1);
                    ce = cudaLaunchKernelEx(&cfg, seir_cluster_kernel, h->cx, k0, cap);
                }
                CK(ce);
                ++h->launches;
            }
            for (int k = 1; k < 5; ++k) {
                uint32_t yield_below = k == 4 ? 0u : h->cx.sparse_cap / 2u + 1u;  // (0: this launch never hands over)
                uint32_t yield_above = h->n_shards > 1 ? h->shard_threshold : 0xFFFFFFFFu;
                void *args[] = {&h->cx, &k, &yield_below, &yield_above};
                if (k % 2 == 1) CK(cudaLaunchCooperativeKernel((const void *)seir_sparse_kernel, dim3(h->sms), dim3(kSparseThreads), args, sizeof(PassSmem), h->stream));
                else CK(cudaLaunchCooperativeKernel((const void *)seir_dense_solo_kernel, dim3(h->grid), dim3(kThreads), args, sizeof(PassSmem), h->stream));
                ++h->launches;
                if (getenv("SEIR_DEBUG_SYNC")) {  // (debugging: which launch faults)
                    const cudaError_t e = cudaStreamSynchronize(h->stream);
                    if (e != cudaSuccess) { char b[128]; snprintf(b, sizeof b, "day-loop launch %d failed: %s", k, cudaGetErrorString(e)); return fail(b); }
                }
            }
            if (h->n_shards > 1) {
                h->cx.start_word = 5;
                h->cx.barrier = h->barrier.p + 6 * kBarStride;  // (its own barrier words: launches 1..4 counted in words 2..5)
                void *args[] = {&h->cx};
                CK(cudaLaunchCooperativeKernel((const void *)seir_persistent_kernel<false>, dim3(h->grid), dim3(kThreads), args, sizeof(PassSmem), h->stream));
                h->cx.barrier = h->barrier.p;
                ++h->launches;
            }
        } else {
        const void *fn = tracing ? (const void *)seir_persistent_kernel<true> : (const void *)seir_persistent_kernel<false>;
        CK(cudaLaunchCooperativeKernel(fn, dim3(h->grid), dim3(kThreads), args, sizeof(PassSmem), h->stream));
        ++h->launches;
        }
    } else {
#ifdef SEIR_SORT_PROBE
        const int sort_mask = getenv("SEIR_DEBUG_SORT") ? atoi(getenv("SEIR_DEBUG_SORT")) : 0;  // 1 transmitting, 2 new, 4 quarantined
        const bool sort_lists = sort_mask != 0;
        cudaEvent_t pa, pb;
        CK(cudaEventCreate(&pa)); CK(cudaEventCreate(&pb));
        h->probe_ms = 0.0f;
#endif
        for (int t = 1; t <= T_loop; ++t) {
#ifdef SEIR_SORT_PROBE
            CK(cudaEventRecord(pa, h->stream));
#endif
            if (tracing) seir_day_kernel<true><<<h->grid, kThreads, sizeof(PassSmem), h->stream>>>(h->cx, t);
            else seir_day_kernel<false><<<h->grid, kThreads, sizeof(PassSmem), h->stream>>>(h->cx, t);
            ++h->launches;
#ifdef SEIR_SORT_PROBE
            {   // the day kernels alone are timed; with SEIR_DEBUG_SORT=1 tomorrow's three list segments are sorted by agent id
                CK(cudaEventRecord(pb, h->stream));
                CK(cudaEventSynchronize(pb));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, pa, pb));
                h->probe_ms += ms;
                if (sort_lists && R == 1 && t < h->prm.T) {
                    uint32_t c[3];
                    const size_t T2 = (size_t)h->prm.T + 2;
                    CK(cudaMemcpy(&c[0], h->list_cnt.p + 0 * T2 + (t + 1), 4, cudaMemcpyDeviceToHost));  // transmitting
                    CK(cudaMemcpy(&c[1], h->list_cnt.p + 1 * T2 + (t + 1), 4, cudaMemcpyDeviceToHost));  // new
                    CK(cudaMemcpy(&c[2], h->list_cnt.p + 2 * T2 + (t + 1), 4, cudaMemcpyDeviceToHost));  // quarantined
                    uint32_t *lo = h->list_old.p + (size_t)((t + 1) & 1) * h->n_pad, *ln = h->list_new.p + (size_t)((t + 1) & 1) * h->n_pad;
                    auto pol = thrust::cuda::par.on(h->stream);
                    if (sort_mask & 1) thrust::sort(pol, lo, lo + c[0]);
                    if (sort_mask & 4) thrust::sort(pol, lo + (h->n_pad - c[2]), lo + h->n_pad);
                    if (sort_mask & 2) thrust::sort(pol, ln, ln + c[1]);
                }
            }
#endif
            if (tracing && t >= h->cx.trace_from && t < T_loop) {
                seir_trace_kernel<<<R < h->grid ? R : h->grid, 32, sizeof(TraceSmem), h->stream>>>(h->cx, t);
                ++h->launches;
            }
        }
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(h->ev2, h->stream));
    h->host_results = false;
    if (h->tally_host) {  // tallies and counters are final here: bring them home while the history is written
        CK(cudaStreamWaitEvent(h->copy_stream, h->ev2, 0));
        // (the copy stream has the higher priority and its kernel is queued first: a CTA that becomes ready together with
        // the materialisation kernel's gets its SM before that kernel fills every thread slot of the device)
        if (h->tally64_host && h->gs_dev.p) {
            seir_tally_accumulate_kernel<<<R, 1024, 0, h->copy_stream>>>(h->tally.p, h->tally64.p, h->gs_dev.p, h->prm.T);
            ++h->launches;
            CK(cudaMemcpyAsync(h->tally64_host, h->tally64.p, (size_t)R * h->prm.T * SEIR_N_TALLY * kAges * 8, cudaMemcpyDeviceToHost, h->copy_stream));
        }
        CK(cudaMemcpyAsync(h->counters_host, h->counters.p, (size_t)R * 4 * 8, cudaMemcpyDeviceToHost, h->copy_stream));
        h->raw_home = !(h->tally64_host && h->gs_dev.p);   // (the raw per-day changes travel only when nothing accumulated them here)
        if (h->raw_home)
            CK(cudaMemcpyAsync(h->tally_host, h->tally.p, (size_t)R * h->prm.T * TAL_ROWS * kAges * 4, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    if (h->history_mode == SEIR_HISTORY_EAGER) {
        seir_materialize_kernel<<<h->sms * 8, kMatThreads, 0, h->stream>>>(h->cx);
        ++h->launches;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev3, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->tally_host) {
        CK(cudaStreamSynchronize(h->copy_stream));
        h->host_results = true;
    }
    CK(cudaEventElapsedTime(&h->loop_ms, h->ev1, h->ev2));
#ifdef SEIR_SORT_PROBE
    if (h->prm.launch == SEIR_LAUNCH_PER_DAY) h->loop_ms = h->probe_ms;  // (the day kernels alone)
#endif
    CK(cudaEventElapsedTime(&h->mat_ms, h->ev2, h->ev3));
    CK(cudaEventElapsedTime(&h->total_ms, h->ev0, h->ev3));
    if (h->n_shards > 1) {
        unsigned int fl[kMaxShards + 2];
        CK(cudaMemcpy(fl, h->flags.p, sizeof fl, cudaMemcpyDeviceToHost));
        if (fl[kMaxShards + 1]) return fail("sharded run: a peer GPU did not reach the day barrier within 4 s");
    }
#ifdef SEIR_CHECKS
    {
        std::vector<unsigned long long> c((size_t)R * 4);
        CK(cudaMemcpy(c.data(), h->counters.p, c.size() * 8, cudaMemcpyDeviceToHost));
        for (int r = 0; r < R; ++r)
            if (c[(size_t)r * 4] & 4ull) return fail("SEIR_CHECKS: an index left its array (see SEIR_CHECK in seir_kernels.cuh)");
    }
#endif
    if (tracing) {
        std::vector<unsigned long long> c((size_t)R * 4);
        if (h->host_results) memcpy(c.data(), h->counters_host, c.size() * 8);
        else CK(cudaMemcpy(c.data(), h->counters.p, c.size() * 8, cudaMemcpyDeviceToHost));
        for (int r = 0; r < R; ++r) {
            if (c[(size_t)r * 4] & 1ull) return fail("contact tracing: the outside-infection log overflowed (more hits than 1.25 n)");
            if (c[(size_t)r * 4] & 2ull) {
                char b[160];
                snprintf(b, sizeof b, "contact tracing: the tracer's search list overflowed (replica %d, flags %llu)", r, c[(size_t)r * 4]);
                return fail(b);
            }
            if (c[(size_t)r * 4] & 8ull) return fail("contact tracing: an agent infected more than 504 others outside its household (the tracer stages at most that many log entries; the reference keeps 100)");
        }
    }
    return 0;
}

int seir_last_run_ms(seir_handle *h, float *day_loop_ms, float *total_ms)
{
    if (!h) return fail("handle is NULL");
    if (day_loop_ms) *day_loop_ms = h->loop_ms;
    if (total_ms) *total_ms = h->total_ms;
    return 0;
}

int seir_last_run_ms3(seir_handle *h, float *init_ms, float *day_loop_ms, float *materialize_ms)
{
    if (!h) return fail("handle is NULL");
    if (init_ms) *init_ms = h->total_ms - h->loop_ms - h->mat_ms;
    if (day_loop_ms) *day_loop_ms = h->loop_ms;
    if (materialize_ms) *materialize_ms = h->mat_ms;
    return 0;
}

// the device's tally rows of one replica: from the pinned copy the run left behind, else copied now
static int raw_tallies(seir_handle *h, int replica, std::vector<uint32_t> &own, const uint32_t **raw_p)
{
    const size_t per = (size_t)h->prm.T * TAL_ROWS * kAges;
    if (h->host_results && h->raw_home) {
        *raw_p = h->tally_host + (size_t)replica * per;
    } else {
        own.resize(per);
        CK(cudaMemcpy(own.data(), h->tally.p + (size_t)replica * per, per * 4, cudaMemcpyDeviceToHost));
        *raw_p = own.data();
    }
    return 0;
}

int seir_get_tallies(seir_handle *h, int replica, int64_t *out)
{
    if (!h || !out) return fail("seir_get_tallies: NULL argument");
    if (replica < 0 || replica >= h->n_rep) return fail("replica out of range");
    CK(cudaSetDevice(h->device));
    const int T = h->prm.T;
    if (h->host_results && h->tally64_host && h->gs_dev.p) {  // accumulated on the device, brought home by the run itself
        const size_t per64 = (size_t)T * SEIR_N_TALLY * kAges;
        memcpy(out, h->tally64_host + (size_t)replica * per64, per64 * 8);
        return 0;
    }
    std::vector<uint32_t> own;
    const uint32_t *raw_p;
    if (raw_tallies(h, replica, own, &raw_p)) return 1;
    // planes in the order of seir_individual.py:392: S,E,Mild,Documented,Severe,Critical,R,D,Q.  The device rows are
    // per-day changes (the day loop visits an agent only when something can happen to it): accumulate them.
    std::vector<int64_t> cum((size_t)TAL_ROWS * kAges, 0);
    for (int t = 0; t < T; ++t) {
        const uint32_t *d = raw_p + (size_t)t * TAL_ROWS * kAges;
        int64_t *o = out + (size_t)t * SEIR_N_TALLY * kAges;
        for (int k = 0; k < TAL_ROWS * kAges; ++k) cum[(size_t)k] += (int64_t)(int32_t)d[k];  // (signed rows wrap around 2^32)
        for (int a = 0; a < kAges; ++a) {
            auto c = [&](int row) { return cum[(size_t)row * kAges + a]; };
            o[0 * kAges + a] = (h->n_shards > 1 ? h->shard_group_sizes[a] : h->group_sizes[a]) - c(TAL_NEW_E);
            o[1 * kAges + a] = c(TAL_NEW_E) + c(TAL_D_E);
            o[2 * kAges + a] = c(TAL_D_MILD);
            o[3 * kAges + a] = c(TAL_NEW_DOC);
            o[4 * kAges + a] = c(TAL_D_SEV);
            o[5 * kAges + a] = c(TAL_D_CRIT);
            o[6 * kAges + a] = c(TAL_NEW_R);
            o[7 * kAges + a] = c(TAL_NEW_D);
            o[8 * kAges + a] = c(TAL_NEW_QE) + c(TAL_D_Q) + c(TAL_NEW_Q2);
        }
    }
    return 0;
}

int seir_get_agent_vector(seir_handle *h, int replica, int which, double *out)
{
    if (!h || !out) return fail("seir_get_agent_vector: NULL argument");
    if (replica < 0 || replica >= h->n_rep) return fail("replica out of range");
    const bool previous = (which & SEIR_VEC_PREVIOUS) != 0;
    which &= ~SEIR_VEC_PREVIOUS;
    if (which < 0 || which >= SEIR_VEC_COUNT) return fail("vector id out of range");
    if (previous && !h->prev_vectors) return fail("no run has been set aside (seir_history_preserve)");
    CK(cudaSetDevice(h->device));
    CK(h->stage_d.alloc((size_t)h->n));
    DevCtx cx = h->cx;
    if (previous) {
        cx.rec = h->rec_prev.p;
        cx.inbox = h->inbox_prev.p;
        cx.trc = h->prev_trc ? h->trc_prev.p : nullptr;
    }
    seir_agent_vector_kernel<<<h->sms * 4, 256, 0, h->stream>>>(cx, replica, which, h->stage_d.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->stage_d.p, (size_t)h->n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int seir_get_cohorts(seir_handle *h, int replica, int64_t *out)
{
    if (!h || !out) return fail("seir_get_cohorts: NULL argument");
    if (replica < 0 || replica >= h->n_rep) return fail("replica out of range");
    CK(cudaSetDevice(h->device));
    const size_t cells = (size_t)h->prm.T * 4 * kAges;
    DevBuf<unsigned long long> buf;
    CK(buf.alloc(cells));
    CK(cudaMemsetAsync(buf.p, 0, cells * 8, h->stream));
    seir_cohort_kernel<<<h->sms * 4, 256, 0, h->stream>>>(h->cx, replica, buf.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, buf.p, cells * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    buf.release();
    return 0;
}

int seir_history_preserve(seir_handle *h)
{
    if (!h) return fail("handle is NULL");
    if (h->history_mode != SEIR_HISTORY_EAGER) return fail("SEIR_HISTORY_ON_DEMAND keeps no history to set aside");
    CK(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->n_rep * h->prm.T * h->n_pad;
    CK(h->hist_prev.alloc(bytes));
    CK(cudaMemcpyAsync(h->hist_prev.p, h->hist.p, bytes, cudaMemcpyDeviceToDevice, h->stream));
    // the per-agent vectors (:392) are functions of the records, the ever-infected bits and the tracer's words
    const size_t cells = (size_t)h->n_rep * h->n_pad;
    CK(h->rec_prev.alloc(cells));
    CK(h->inbox_prev.alloc(cells / 16));
    CK(cudaMemcpyAsync(h->rec_prev.p, h->rec.p, cells * sizeof(AgentRec), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->inbox_prev.p, h->inbox.p, cells / 16 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
    h->prev_trc = h->trace_alloc && h->trc.p != nullptr;
    if (h->prev_trc) {
        CK(h->trc_prev.alloc(cells));
        CK(cudaMemcpyAsync(h->trc_prev.p, h->trc.p, cells * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    h->prev_T = h->prm.T;
    h->prev_vectors = true;
    return 0;
}

int seir_get_history(seir_handle *h, int replica, int which, int t0, int t1, uint8_t *out)
{
    if (!h || !out) return fail("seir_get_history: NULL argument");
    if (replica < 0 || replica >= h->n_rep) return fail("replica out of range");
    const bool previous = (which & SEIR_H_PREVIOUS) != 0;
    which &= ~SEIR_H_PREVIOUS;
    if (which < 0 || which > SEIR_H_PACKED) return fail("history plane out of range");
    if (previous && !h->prev_T) return fail("no history has been set aside (seir_history_preserve)");
    const int T_rows = previous ? h->prev_T : h->prm.T;
    const int T_rep = previous || h->rp_host.empty() ? T_rows : h->rp_host[(size_t)replica].T;
    if (t0 < 0 || t1 > T_rep || t0 > t1) return fail("day range out of bounds");
    if (h->history_mode == SEIR_HISTORY_ON_DEMAND) {
        if (previous) return fail("SEIR_HISTORY_ON_DEMAND keeps no earlier history");
        if (h->rp_host.empty()) return fail("no run yet");
        CK(cudaSetDevice(h->device));
        const int64_t n = h->n, np = h->n_pad;
        int rows_per = (int)((int64_t)(256ll << 20) / np);
        if (rows_per < 1) rows_per = 1;
        DevBuf<uint8_t> rows;
        CK(rows.alloc((size_t)rows_per * np));
        CK(h->stage_u8.alloc((size_t)rows_per * n));
        for (int t = t0; t < t1; t += rows_per) {
            const int nr = t1 - t < rows_per ? t1 - t : rows_per;
            seir_rows_kernel<<<h->sms * 8, 256, 0, h->stream>>>(h->cx, replica, t, t + nr, rows.p);
            seir_history_plane_kernel<<<h->sms * 8, 256, 0, h->stream>>>(rows.p, n, np, nr, which, h->stage_u8.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(out + (size_t)(t - t0) * n, h->stage_u8.p, (size_t)nr * n, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        rows.release();
        return 0;
    }
    CK(cudaSetDevice(h->device));
    const int64_t n = h->n;
    int rows_per = (int)((int64_t)(256ll << 20) / (n > 0 ? n : 1));
    if (rows_per < 1) rows_per = 1;
    CK(h->stage_u8.alloc((size_t)rows_per * n > (size_t)n ? (size_t)rows_per * n : (size_t)n));
    const uint8_t *base = (previous ? h->hist_prev.p : h->hist.p) + (size_t)replica * T_rows * h->n_pad;
    for (int t = t0; t < t1; t += rows_per) {
        const int nr = t1 - t < rows_per ? t1 - t : rows_per;
        seir_history_plane_kernel<<<h->sms * 8, 256, 0, h->stream>>>(base + (size_t)t * h->n_pad, n, h->n_pad, nr, which, h->stage_u8.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out + (size_t)(t - t0) * n, h->stage_u8.p, (size_t)nr * n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

int seir_get_prologue(seir_handle *h, int replica, uint8_t *home_out, int64_t *init_out, int64_t cap, int64_t *n_init)
{
    if (!h || !n_init) return fail("seir_get_prologue: NULL argument");
    if (replica < 0 || replica >= h->n_rep) return fail("replica out of range");
    if (h->n_init_run.empty()) return fail("no run yet");
    CK(cudaSetDevice(h->device));
    const int64_t n = h->n;
    if (home_out) {
        std::vector<uint32_t> bm((size_t)(h->n_pad / 32), 0u);
        if (h->home_valid) CK(cudaMemcpy(bm.data(), h->home.p + (size_t)replica * (h->n_pad / 32), bm.size() * 4, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; ++i) home_out[i] = (uint8_t)((bm[(size_t)(i >> 5)] >> (i & 31)) & 1u);
    }
    int64_t off = 0;
    for (int r = 0; r < replica; ++r) off += h->n_init_run[(size_t)r];
    const int64_t k = h->n_init_run[(size_t)replica];
    *n_init = k;
    if (init_out) {
        if (cap < k) return fail("seir_get_prologue: init_out is too small");
        std::vector<int32_t> ids((size_t)k);
        if (k) CK(cudaMemcpy(ids.data(), h->init_ids.p + off, (size_t)k * 4, cudaMemcpyDeviceToHost));
        for (int64_t j = 0; j < k; ++j) init_out[j] = ids[(size_t)j];
    }
    return 0;
}

int seir_get_day_times(seir_handle *h, int64_t *out)
{
    if (!h || !out) return fail("seir_get_day_times: NULL argument");
    CK(cudaSetDevice(h->device));
    std::vector<unsigned long long> v((size_t)h->prm.T + 2);
    CK(cudaMemcpy(v.data(), h->day_ns.p, v.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < v.size(); ++k) out[k] = (int64_t)v[k];
    return 0;
}

int seir_get_stage_times(seir_handle *h, int64_t *out)
{
    if (!h || !out) return fail("seir_get_stage_times: NULL argument");
    CK(cudaSetDevice(h->device));
    std::vector<unsigned long long> v(((size_t)h->prm.T + 2) * 8);
    CK(cudaMemcpy(v.data(), h->stage_ns.p, v.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < v.size(); ++k) out[k] = (int64_t)v[k];
    return 0;
}

// (debug) the hand-over words of the last single-replica run: launch k started at day out[k] + 1
int seir_get_handover(seir_handle *h, int32_t *out, int n)
{
    if (!h || !out) return fail("seir_get_handover: NULL argument");
    CK(cudaSetDevice(h->device));
    if (n > kMaxLaunches + 1) n = kMaxLaunches + 1;
    CK(cudaMemcpy(out, h->ctl.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int seir_get_fine_times(seir_handle *h, int64_t *out)
{
    if (!h || !out) return fail("seir_get_fine_times: NULL argument");
    CK(cudaSetDevice(h->device));
    std::vector<unsigned long long> v(((size_t)h->prm.T + 2) * 16);
    CK(cudaMemcpy(v.data(), h->fine_ns.p, v.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < v.size(); ++k) out[k] = (int64_t)v[k];
    return 0;
}

int seir_get_counters(seir_handle *h, int replica, int64_t *out4)
{
    if (!h || !out4) return fail("seir_get_counters: NULL argument");
    if (replica < 0 || replica >= h->n_rep) return fail("replica out of range");
    CK(cudaSetDevice(h->device));
    unsigned long long c[4];
    if (h->host_results) memcpy(c, h->counters_host + (size_t)replica * 4, sizeof c);
    else CK(cudaMemcpy(c, h->counters.p + (size_t)replica * 4, sizeof c, cudaMemcpyDeviceToHost));
    out4[0] = h->launches; out4[2] = (int64_t)c[2]; out4[3] = (int64_t)c[3];
    // sum over the days of A(t-1), the agents in E/Mild/Severe/Critical: from the tallies (the day loop visits an
    // active agent only on the days on which something can happen to it; c[1] counts those visits)
    std::vector<uint32_t> own;
    const uint32_t *raw_p;
    if (raw_tallies(h, replica, own, &raw_p)) return 1;
    int64_t active = 0, sum = 0;
    const int T_rep = h->rp_host.empty() ? h->prm.T : h->rp_host[(size_t)replica].T;
    for (int t = 0; t + 1 < T_rep; ++t) {
        const uint32_t *d = raw_p + (size_t)t * TAL_ROWS * kAges;
        for (int a = 0; a < kAges; ++a)
            active += (int64_t)(int32_t)d[TAL_NEW_E * kAges + a] + (int32_t)d[TAL_D_E * kAges + a] + (int32_t)d[TAL_D_MILD * kAges + a] +
                      (int32_t)d[TAL_D_SEV * kAges + a] + (int32_t)d[TAL_D_CRIT * kAges + a];
        sum += active;
    }
    out4[1] = sum;
    h->visits = (long long)c[1];
    return 0;
}

int seir_get_visits(seir_handle *h, int replica, int64_t *out)
{
    int64_t c4[4];
    if (!out) return fail("seir_get_visits: NULL argument");
    if (seir_get_counters(h, replica, c4)) return 1;
    *out = h->visits;
    return 0;
}

/* debug: raw copy of one of the tracer's per-replica arrays (0 ctl, 1 roots, 2 reps, 3 csz, 4 reached, 5 lead, 6 comp) */
int seir_debug_trace_array(seir_handle *h, int which, int replica, void *out, uint64_t bytes)
{
    if (!h || !out || !h->trace_alloc) return fail("seir_debug_trace_array: no tracer state");
    CK(cudaSetDevice(h->device));
    const size_t np = (size_t)h->n_pad, r = (size_t)replica;
    const void *src = nullptr;
    switch (which) {
        case 0: src = h->tctl.p + r; break;
        case 1: src = h->troots.p + r * np; break;
        case 2: src = h->treps.p + r * np; break;
        case 3: src = h->tcsz.p + r * np; break;
        case 4: src = h->treached.p + r * np; break;
        case 5: src = h->tlead.p + r * np; break;
        case 6: src = h->tcomp.p + r * np; break;
        default: return fail("seir_debug_trace_array: unknown array");
    }
    CK(cudaMemcpy(out, src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int seir_host_register(void *ptr, uint64_t bytes)
{
    if (!ptr || !bytes) return fail("seir_host_register: NULL / empty buffer");
    CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return 0;
}

int seir_host_unregister(void *ptr)
{
    if (!ptr) return fail("seir_host_unregister: NULL buffer");
    CK(cudaHostUnregister(ptr));
    return 0;
}

void seir_destroy(seir_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    h->stat.release(); h->grp_off.release(); h->grp_members.release(); h->init_ids.release(); h->init_off.release();
    h->p_hh.release(); h->tables.release(); h->enlam.release(); h->stage_d.release(); h->hist.release(); h->rec_prev.release(); h->inbox_prev.release(); h->trc_prev.release();
    h->stage_u8.release(); h->inbox.release(); h->tally.release(); h->home.release(); h->winner.release(); h->xday.release();
    h->counters.release(); h->rec.release(); h->seeds.release(); h->barrier.release(); h->ctl.release();
    for (int g = 0; g < kMaxShards; ++g)
        for (int k = 0; k < SEIR_SHARD_HANDLES; ++k)
            if (h->peer_ptr[g][k]) cudaIpcCloseMemHandle(h->peer_ptr[g][k]);
    h->flags.release();
    h->list_old.release(); h->list_new.release(); h->list_cnt.release(); h->day_ns.release();
    h->log_ent.release(); h->log_head.release(); h->log_cnt.release(); h->rootbits.release(); h->rootsum.release();
    h->root_cnt.release(); h->tmark.release(); h->trc.release(); h->tstack.release();
    h->rootsum2.release(); h->tcomp.release(); h->tlead.release(); h->treached.release(); h->troots.release();
    h->treps.release(); h->tcsz.release(); h->tctl.release(); h->ttotal.release(); h->tcta.release(); h->tbig.release();
    h->tedge_off.release(); h->tedge_cnt.release(); h->tedge.release();
    h->alias_idx.release(); h->stage_ns.release(); h->fine_ns.release();
    if (h->ev3) cudaEventDestroy(h->ev3);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->in_host) cudaFreeHost(h->in_host);
    if (h->tally_host) cudaFreeHost(h->tally_host);
    if (h->tally64_host) cudaFreeHost(h->tally64_host);
    h->tally64.release(); h->gs_dev.release();
    if (h->counters_host) cudaFreeHost(h->counters_host);
    delete h;
}

int seir_probe_draws(int device, uint64_t seed, int64_t m, double mean, double mu, double sigma, double enlam,
                     double *exp_days_out, double *lognormal_out, int64_t *poisson_out, double *log_out,
                     const double *log_in)
{
    if (m < 1) return fail("m must be positive");
    CK(cudaSetDevice(device));
    DevBuf<double> a, b, c, d;
    DevBuf<long long> p;
    CK(a.alloc(m)); CK(b.alloc(m)); CK(c.alloc(m)); CK(d.alloc(m)); CK(p.alloc(m));
    CK(cudaMemcpy(d.p, log_in, m * 8, cudaMemcpyHostToDevice));
    seir_probe_kernel<<<256, 256>>>(seed, m, mean, mu, sigma, enlam, a.p, b.p, p.p, c.p, d.p);
    CK(cudaGetLastError());
    CK(cudaMemcpy(exp_days_out, a.p, m * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lognormal_out, b.p, m * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(poisson_out, p.p, m * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(log_out, c.p, m * 8, cudaMemcpyDeviceToHost));
    a.release(); b.release(); c.release(); d.release(); p.release();
    return 0;
}

}  // extern "C"
