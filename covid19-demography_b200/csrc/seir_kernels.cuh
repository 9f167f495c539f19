// seir_kernels.cuh -- device side of the B200 SEIR day step (sm_100a).
//
// Layout in HBM (per replica unless marked shared):
//   stat[n]        u16, shared   age(7) | diabetes | hypertension | household position(3) | household size-1 (3)
//   grp_off/members i32, shared  CSR of the reference's age_groups tuple (seir_individual.py:36)
//   winner[n]      u64           0 = never hit; else day<<48 | infector<<16 | event ordinal of the LAST hit
//                                in the reference's agent order (atomicMax).  winner is also the
//                                susceptibility oracle: S[t-1,c]  <=>  winner[c]==0 || day(winner[c])==t.
//   xday[n]        u16           sharded runs only: day of infection + 1 (0 = never).  EVERY GPU holds a full copy, so the susceptibility
//                                test of a community contact, S[t-1,c] <=> xday[c]==0 || xday[c]==t+1, never leaves the GPU: a first
//                                hitter writes its own copy, and every morning each GPU completes its copy from the owners' lists of
//                                the day's new infectees (xday_refresh)
//   inbox[n/16]    u32           2 bits per agent: "ever infected", "an isolation draw of a hit was 0 => Q" (:353-354)
//   rec[n]         32 B          one sector per EVER-INFECTED agent: state byte, static word, every event day and
//                                drawn time-to-event; the whole T-day history of the agent is a function of it.
//   list_old/new[2][n] u32       daily lists: agents that may TRANSMIT today (E / Mild, not quarantined), and agents
//                                infected yesterday (record not built yet)
//   pk_pool[16][n] (day buckets) quarantined agents (Q: they neither transmit nor draw) are PARKED in the bucket of the
//                                next day on which one of their `t - t0 == draw` tests can fire, and visited only then
//   hist[T][n]     u8            packed state rows (bits 0-2 compartment, bit 3 Q, bit 4 Documented): the
//                                reference's nine bool[T,n] outputs, written ONCE by the materialise kernel.
//
// Day pass t (1 <= t <= T; pass T only finalises day T-1): one thread per entry of the day's lists, in
//   ROUNDS of kThreads entries that move through four convergent stages with shared-memory work queues:
//   A own progression (:256-337) + which household members / how many contacts may be infected,
//   C one candidate contact per thread (:373-375), H one hit per thread (:347-356,:376-385),
//   W write-back of the infector's counters (:357-359,:386-390) and re-append to tomorrow's list.
//   The FIRST sender that hits a susceptible c on day t (atomicMax returns a record of an earlier day)
//   appends c to tomorrow's "new" list.
//   S, R and D agents are never touched by a day pass (the reference's iteration is a no-op for them).
// All cross-agent effects of a day go through atomics on the infectee (winner, inbox), so ONE grid-wide
// barrier per day suffices and the kernel runs persistently over the whole horizon.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seir_b200.h"
#include "../../include/seir_draws.h"

namespace seir {

constexpr int kAges = SEIR_N_AGES;
#ifndef SEIR_THREADS
#define SEIR_THREADS 512
#endif
#ifndef SEIR_MIN_BLOCKS
#define SEIR_MIN_BLOCKS 2
#endif
#ifndef SEIR_EVE_PARK
#define SEIR_EVE_PARK 1   // staged passes: agents with a progression step due tomorrow go through tomorrow's day bucket (agent_day)
#endif
constexpr int kThreads = SEIR_THREADS;   // threads per CTA of the day-step kernels
constexpr int kMinBlocks = SEIR_MIN_BLOCKS;
constexpr int kVec = 16;                 // agents per 16 B vector of state bytes
constexpr int kPad = 4096;               // row stride is a multiple of this many agents
constexpr int kOutCap = 3 * kThreads;    // per-CTA staging of list appends
constexpr int kParkHorizon = 15;         // an agent is parked at most this many days ahead (it is looked at again then),
constexpr int kParkRing = kParkHorizon + 1;  // so the buckets of days t .. t + 15 are the live ones: day d uses slab d mod 16
constexpr int kDocScan = 8;              // days of documentation draws (:280-289) evaluated ahead when a Mild agent is parked

// device tally rows (per day, per age)
// Every row is a per-day CHANGE (an agent is only visited on days on which something may happen to it, so nothing
// is recounted): seir_tally_accumulate_kernel accumulates them into the reference's daily counts (read by seir_get_tallies).
enum { TAL_NEW_E = 0,  // infections of the day (S -> E)
       TAL_NEW_QE,     // of them quarantined on infection (:353-354, or reached by a tracer the same day)
       TAL_D_E, TAL_D_MILD, TAL_D_SEV, TAL_D_CRIT,  // signed: change of the E / Mild / Severe / Critical head counts by the day's transitions
       TAL_D_Q,        // signed: change of the number of quarantined E/Mild/Severe/Critical agents
       TAL_NEW_R, TAL_NEW_D, TAL_NEW_DOC,
       TAL_NEW_Q2,     // recovered / dead agents quarantined by a tracer (:72-76 traces every non-S contact)
       TAL_ROWS };
constexpr int kTalLagRows = TAL_NEW_QE + 1;  // rows [0, kTalLagRows) of a pass describe day t-1 (the records built today), the rest day t

// compartments (bits 0-2 of the state byte)
enum { C_S = 0, C_E = 1, C_MILD = 2, C_SEV = 3, C_CRIT = 4, C_R = 5, C_D = 6 };
constexpr uint8_t ST_Q = 0x08, ST_DOC = 0x10;
constexpr uint8_t F_MILD = 1, F_SEV = 2, F_CRIT = 4;  // stages the agent has reached
constexpr uint8_t F_TRX = 8;                          // trc[agent] is valid (the agent was reached by a tracer)
// winner word: bit 63 = "reached by a tracer on its infection day" (the agent has no record yet that day)
constexpr unsigned long long W_TRACED = 1ull << 63;
__host__ __device__ __forceinline__ int wday(unsigned long long w) { return (int)((w >> 48) & 0x7FFFu); }
constexpr int kLogWidth = 100;  // width of the reference's infected_by log (:197)
constexpr int kMaxShards = 8;   // GPUs of one box
constexpr int kMaxLaunches = 64; // day-loop launches of one single-replica run (sparse-day and dense-day kernels in turn)
constexpr uint16_t NEVER = 0xFFFFu;
constexpr uint32_t kHole = 0xFFFFFFFFu;  // an entry of the transmitting list whose agent left it during a warp-cooperative pass

// 32 B per-agent record.  Derived days: t_infected = t_exposed + tt_activation (F_MILD),
// t_severe = t_infected + tt_severe (F_SEV), t_critical = t_severe + tt_critical (F_CRIT) -- the reference
// fires those transitions exactly when `t - t0 == tt` (:257,:291,:312).
// Kept as eight 32-bit words with constant-index accessors so that it lives in REGISTERS (a struct of
// sixteen uint16 fields ends up in local memory as soon as its address is taken).
struct __align__(16) AgentRec {
    uint32_t w[8];
#define SEIR_FIELD(name, idx)                                                                                          \
    __host__ __device__ __forceinline__ uint16_t name() const { return (uint16_t)(w[(idx) >> 1] >> (((idx)&1) * 16)); } \
    __host__ __device__ __forceinline__ void set_##name(uint32_t v)                                                    \
    {                                                                                                                  \
        w[(idx) >> 1] = ((idx)&1) ? ((w[(idx) >> 1] & 0x0000FFFFu) | (v << 16)) : ((w[(idx) >> 1] & 0xFFFF0000u) | (v & 0xFFFFu)); \
    }
    SEIR_FIELD(t_exposed, 0)      // :200,:355
    SEIR_FIELD(tt_activation, 1)  // :218,:356
    SEIR_FIELD(tt_severe, 2)      // :213,:263  (0xFFFF == inf)
    SEIR_FIELD(tt_recovery, 3)    // :214
    SEIR_FIELD(tt_critical, 4)    // :215
    SEIR_FIELD(tt_death, 5)       // :216
    SEIR_FIELD(tt_isolate, 6)     // :217
    SEIR_FIELD(t_documented, 7)   // :204
    SEIR_FIELD(n_inf_by, 8)       // :207
    SEIR_FIELD(n_inf_asympt, 9)   // :209
    SEIR_FIELD(n_inf_outside, 10) // :208
    SEIR_FIELD(t_end, 11)         // day of R or D (0xFFFF: none yet)
    SEIR_FIELD(t_q_on, 12)        // first day with Q set (0xFFFF: never)
    SEIR_FIELD(stat, 13)          // copy of the static word
#undef SEIR_FIELD
    __host__ __device__ __forceinline__ uint8_t state() const { return (uint8_t)(w[7] & 0xFFu); }         // packed state byte
    __host__ __device__ __forceinline__ uint8_t flags() const { return (uint8_t)((w[7] >> 8) & 0xFFu); }  // F_*
    __host__ __device__ __forceinline__ void set_state(uint32_t v) { w[7] = (w[7] & ~0xFFu) | (v & 0xFFu); }
    __host__ __device__ __forceinline__ void set_flags(uint32_t v) { w[7] = (w[7] & ~0xFF00u) | ((v & 0xFFu) << 8); }
    __host__ __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 0;
    }
};
static_assert(sizeof(AgentRec) == 32, "AgentRec must be one 32 B sector");

// static per-agent word: age(7) | diabetes<<7 | hypertension<<8 | household position(3)<<9 | (household size-1)(3)<<12
__host__ __device__ inline uint16_t pack_static(int age, int diab, int hyp, int pos, int size)
{
    return (uint16_t)(age | (diab << 7) | (hyp << 8) | (pos << 9) | ((size - 1) << 12));
}

// inbox word (16 agents): byte b holds agents b, b+4, b+8, b+12 -- "infected" in bits 0-3,
// "Q on infection" in bits 4-7 -- so that (word >> w) & 0x01010101 lines the infected bits of agents
// 4w..4w+3 up with the bytes of a 32-bit word of four state bytes (used by the materialise kernel).
__host__ __device__ inline uint32_t inbox_infected_bit(int k) { return 1u << (8 * (k & 3) + (k >> 2)); }
__host__ __device__ inline uint32_t inbox_q_bit(int k) { return 1u << (8 * (k & 3) + 4 + (k >> 2)); }

// per-replica overrides (seir_overrides): day numbers and the table set of the replica's grid point
struct RepParams { int T, t_lockdown, t_stayhome, trace_from, bank, pad[3]; };
// one table set of the bank: what parameter_sweep.py's grid changes (:54-56) -- the three severity tables
// (mortality_multiplier), the kept-contact Poisson table and the contact probability (p_infect_given_contact)
constexpr int kBankPms = 0, kBankPsc = 404, kBankPcd = 808, kBankNatEn = 1212, kBankPcontact = 1616, kBankStride = 1624;
struct TraceCtl;
struct DevCtx {
    // population, shared by all replicas
    int64_t n;            // agents
    int64_t n_pad;        // row stride, multiple of kPad
    int n_rep;
    int blocks_per_rep;   // CTAs that share one replica's list in a pass (>= 1)
    int T;                // rows of the per-day arrays (the handle's horizon)
    int T_loop;           // the longest horizon of this launch's replicas (<= T)
    int mode;
    const RepParams *rp;  // [rep]
    const double *bank;   // [sets][kBankStride]
    const uint16_t *stat;
    const int32_t *grp_off;      // [102]
    const int32_t *grp_members;  // [n]
    const double *p_hh_vec;      // [n] or nullptr
    double p_hh_scalar;
    // tables
    const double *p_ms, *p_sc, *p_cd;  // [101*4]
    const double *iso_mean, *iso_asym; // [101]
    const double *nat_enlam;           // [2][101][2]
    const double *nat_alias_prob;      // [101][101]
    const int32_t *nat_alias_idx;      // [101][101]
    const double *enlam;               // [2][101][101] (replay) or nullptr
    // per-replica state
    uint8_t *hist;                 // [rep][T][n_pad]
    uint32_t *inbox;               // [rep][n_pad/16]
    unsigned long long *winner;    // [rep][n_pad]
    uint16_t *xday;                // [rep][n_pad] day of infection + 1 (0: never infected)
    AgentRec *rec;                 // [rep][n_pad]
    uint32_t *list_old;            // [rep][2][n_pad]  agents that transmitted yesterday and may transmit today
    uint32_t *list_new;            // [rep][2][n_pad]  agents infected yesterday (record not built yet)
    uint32_t *list_cnt;            // [rep][3][T+2]    entries pass t consumes: transmitting list / new list / parked agents due on day t
    // day buckets of the parked (quarantined) agents: the agents due on day d are pk_pool[d mod 16][0 .. list_cnt[2][d])
    uint32_t *pk_pool;             // [rep][kParkRing][n_pad]
    uint32_t *tally;               // [rep][T][TAL_ROWS][101]
    const uint32_t *home;          // [rep][n_pad/32] bitmap of Home_real
    uint32_t *home_rw;             // (the same, for the keyed prologue kernel)
    const uint64_t *seeds;         // [rep]
    unsigned long long *counters;  // [rep][4]
    unsigned int *barrier;         // grid barrier counter
    unsigned long long *day_ns;    // [T+2] %globaltimer at the start of pass t (CTA 0), [T+1] = end of the loop
    unsigned long long *fine_ns;   // [T+2][16] debug stamps inside stage A (only with -DSEIR_FINE_STAMPS)
    unsigned long long *stage_ns;  // [T+2][8] CTA 0, first round of pass t: stamps after init / A / C / H / W / flush / end
    // contact tracing (:60-88); all null / 0 when tracing can never switch on
    int trace_from;                // first day with contact_tracing == True (:241-242), 0 = never
    double p_trace_hh, p_trace_out;
    uint4 *log_ent;                // [rep][log_cap] outside-infection log: infectee, next entry + 1, day<<16 | ordinal, -
    uint32_t *log_head;            // [rep][n_pad]   most recent entry + 1 of the infector (0 = none)
    uint32_t *log_cnt;             // [rep]
    uint32_t log_cap;
    uint32_t *rootbits;            // [rep][n_pad/32]   agents documented today by their own progression (tracing roots)
    uint32_t *rootsum;             // [rep][n_pad/1024] bit w set = rootbits word w non-zero
    uint32_t *root_cnt;            // [rep][T+2] roots of day t
    unsigned long long *tmark;     // [rep][n_pad] day<<32 | root + 1 of the last tracer visit
    uint32_t *trc;                 // [rep][n_pad] first post-end traced-Q day | last traced day<<16 (valid with F_TRX)
    uint32_t *tstack;              // [rep][n_pad] DFS stacks of the tracing warps
    uint32_t *rootsum2;            // [rep][ceil(n_pad/32768)] bit s set = rootsum word s non-zero
    unsigned long long *tcomp, *tlead;  // [rep][n_pad] parallel tracer: union-find parents, component leaders (day-stamped)
    uint32_t *treached, *troots, *treps, *tcsz;  // [rep][n_pad]
    struct TraceCtl *tctl;         // [rep]
    uint32_t *tcta;                // [rep][grid] roots found by each CTA in its slice of the root bitmap
    uint32_t *tedge_off, *tedge_cnt;  // [rep][n_pad] the day's successful edges of a reached agent: first slot in tedge / how many (kEdgeNotCached: none kept)
    uint32_t *tedge;               // [rep][tedge_cap] target | kEdgeHH | kEdgeToday, written by the frontier step, read by every root's search
    uint32_t tedge_cap;
    uint32_t *tbig;                // [rep][kBigCap] leaders (root slots) of the components that get a whole CTA
    uint32_t sparse_rounds;        // a day whose lists hold at most this many entries PER WARP of the replica's CTAs runs as a
                                   // warp-cooperative pass (sparse_pass: one agent per warp, its draws and gathers spread over
                                   // the lanes); 0: never.  env SEIR_SPARSE_ROUNDS
    uint32_t sparse_cap;           // single-replica runs: a day with at most this many list entries is run by seir_sparse_kernel ...
    int start_word;                // the generic persistent kernel starts at day ctl[start_word] + 1 (-1: day 1)
    uint32_t run_nonce;            // tags the timers a hitter leaves in its infectee's record-to-be (never 0, new for every run)
    int *ctl;                      // [kMaxLaunches + 1] launch k of a run starts at day ctl[k] + 1 and leaves the next day to run in ctl[k + 1]
    uint32_t sparse_max;           // ... and at most this many entries in all (tests: SEIR_SPARSE_MAX makes a small run switch back and forth)
    int dbg_pass;                  // (debug builds with -DSEIR_CTA_STAMPS) the pass whose per-CTA stage stamps go to fine_ns; env SEIR_DEBUG_PASS
    uint32_t tday0;                // run number << 12: the tracer's day stamps (tmark, tcomp, tlead) are tday0 + t and never need clearing
    unsigned long long *ttotal;    // nodes ever appended to a frontier (grows monotonically: the BFS loop's stop test)
    // one population across the GPUs of a box (household-aligned shards, peer memory over NVLink)
    int n_shards, my_shard;        // n_shards <= 1: not sharded
    uint32_t shard_threshold;      // a sharded run steps the WHOLE population on every GPU (replicated, no exchange, no
                                   // cross-GPU barrier) until a day's lists exceed this many entries; 0: sharded from day 1
    uint32_t shard_lo[kMaxShards + 1];             // shard g owns agents [shard_lo[g], shard_lo[g+1])
    unsigned long long *peer_winner[kMaxShards];   // every GPU keeps full-size arrays in GLOBAL agent indexing;
    uint32_t *peer_inbox[kMaxShards];              // an agent's winner / inbox / list entries live on its owner
    uint32_t *peer_list_new[kMaxShards];
    uint32_t *peer_list_cnt[kMaxShards];
    uint16_t *peer_xday[kMaxShards];               // every GPU's copy of xday: the first hitter of an agent stores into all of them
    unsigned int *peer_flags[kMaxShards];          // [kMaxShards + 2] per GPU: day epochs of the peers, release word, error word
    unsigned int epoch_base;                       // run number * (T + 2): epochs never repeat, flags need no reset
    // scalars
    int t_lockdown, t_stayhome;
    double p_doc, p_contact, asym;
    double m_sev, m_mild_rec, m_crit, m_sev_rec, m_crit_rec, m_death, act_mu, act_sigma;
};

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t add_sat16(uint16_t a, uint16_t b)
{
    uint32_t s = (uint32_t)a + (uint32_t)b;
    return (a == NEVER || b == NEVER || s >= 0xFFFFu) ? NEVER : (uint16_t)s;
}
__device__ __forceinline__ uint16_t inc_sat16(uint16_t a, uint32_t by)
{
    uint32_t s = (uint32_t)a + by;
    return s >= 0xFFFEu ? (uint16_t)0xFFFEu : (uint16_t)s;
}
// `t - t0 == tt` of the reference on integer-valued doubles; tt == NEVER encodes inf/nan
__device__ __forceinline__ bool due(int t, int t0, uint16_t tt) { return tt != NEVER && t0 + (int)tt == t; }

__device__ __forceinline__ AgentRec load_rec(const AgentRec *p)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldcg(q), b = __ldcg(q + 1);
    AgentRec r;
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
    r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
    return r;
}
__device__ __forceinline__ void store_rec(AgentRec *p, const AgentRec &r)
{
    uint4 *q = reinterpret_cast<uint4 *>(p);
    __stcg(q, make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]));
    __stcg(q + 1, make_uint4(r.w[4], r.w[5], r.w[6], r.w[7]));
}
__device__ __forceinline__ int t_infected_of(const AgentRec &r) { return (int)r.t_exposed() + (int)r.tt_activation(); }
__device__ __forceinline__ int t_severe_of(const AgentRec &r) { return t_infected_of(r) + (int)r.tt_severe(); }
__device__ __forceinline__ int t_critical_of(const AgentRec &r) { return t_severe_of(r) + (int)r.tt_critical(); }

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v) : : "memory");  // (memory: the stamps stay where they are written)
    return v;
}

#ifdef SEIR_FINE_STAMPS
#define SEIR_FINE(k) do { if (blockIdx.x == 0 && slot == 0) cx.fine_ns[(size_t)t * 16 + (k)] = global_timer_ns(); } while (0)
#else
#define SEIR_FINE(k) do { } while (0)
#endif
// Bounds checks of our own (compute-sanitizer is closed on this pool): with -DSEIR_CHECKS every index that reaches a
// list, a record or a queue is tested, a violation sets bit 2 of the replica's error word (seir_run then fails)
// and the access is skipped.  Off in the product build.
#ifdef SEIR_CHECKS
#define SEIR_CHECK(cond, cx, rep) (!(cond) ? (atomicOr(&(cx).counters[(rep) * 4 + 0], 4ull), false) : true)
#else
#define SEIR_CHECK(cond, cx, rep) true
#endif

// ---- out-of-line draw helpers ----------------------------------------------------------------------------
// The day step is latency-bound and its code is executed by few warps per pass, so instruction fetch is
// on the critical path: every heavy piece of the draw contract exists ONCE in the kernel image.
__device__ __noinline__ uint4 draw4(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub)
{
    const sd_block b = sd_draw(seed, agent, day, site, sub);
    return make_uint4(b.w[0], b.w[1], b.w[2], b.w[3]);
}
__device__ __forceinline__ sd_block as_block(uint4 v)
{
    sd_block b;
    b.w[0] = v.x; b.w[1] = v.y; b.w[2] = v.z; b.w[3] = v.w;
    return b;
}
__device__ __forceinline__ sd_block draw_block(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub)
{
    return as_block(draw4(seed, agent, day, site, sub));
}
__device__ __noinline__ double exp_days_ni(double u, double mean) { return sd_exp_days(u, mean); }
__device__ __noinline__ double lognormal_days_ni(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub_base,
                                                 double mu, double sigma)
{
    return sd_lognormal_days(seed, agent, day, site, sub_base, mu, sigma);
}
// continuation of sd_poisson_day after the first uniform (lane 1 of the day block) did not end the product
// (general form: the product so far, X contacts counted, the next block is `call`)
__device__ __noinline__ uint32_t poisson_day_from(uint64_t seed, uint32_t agent, uint32_t day, double prod, double enlam, uint32_t cap,
                                                  uint32_t X, uint32_t call)
{
    for (;; ++call) {
        const sd_block b = sd_draw(seed, agent, day, SD_SITE_NCNT, call);
        prod = SD_MUL(prod, sd_lane(b, 0));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
        prod = SD_MUL(prod, sd_lane(b, 1));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
    }
}
__device__ __noinline__ double log_ni(double x) { return sd_log(x); }
__device__ __noinline__ double exp_ni(double x) { return sd_exp(x); }
__device__ __noinline__ uint32_t poisson_day_more(uint64_t seed, uint32_t agent, uint32_t day, double prod, double enlam, uint32_t cap)
{
    uint32_t X = 1;
    for (uint32_t call = 0;; ++call) {
        const sd_block b = sd_draw(seed, agent, day, SD_SITE_NCNT, call);
        prod = SD_MUL(prod, sd_lane(b, 0));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
        prod = SD_MUL(prod, sd_lane(b, 1));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
    }
}
__device__ __noinline__ uint32_t poisson_ni(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub_base,
                                            double enlam, uint32_t cap)
{
    return sd_poisson(seed, agent, day, site, sub_base, enlam, cap);
}

// per-CTA shared state of a day pass.  A ROUND takes up to kSlots list entries through six convergent
// stages separated by __syncthreads(), with shared-memory work queues between them:
//   A  every entry: record build/load, tallies of day t-1, the draw-free transitions (recovery, death,
//      isolation entry), the documentation draw, the household S-mask, the first community uniform.
//      Everything rare or heavy is QUEUED, not executed in place: a warp pays for every path any of
//      its lanes takes, so stage A keeps to the path (almost) every lane takes
//   E  agents with a progression event due today (E->Mild, Mild->Severe, Severe->Critical), compacted
//   D  draw items: one Philox block of household draws / one sender's kept-contact count
//   C  candidate contacts -> H hits -> W infector counters.
// Sparse lists and short queues are dealt ONE ITEM PER WARP (then 2, 4, ... per warp as they grow).
// No per-thread state survives a stage: stages talk through the queues below and the records in L2.
#ifndef SEIR_ENTRIES_PER_THREAD
#define SEIR_ENTRIES_PER_THREAD 4
#endif
constexpr int kPerThread = SEIR_ENTRIES_PER_THREAD;  // list entries a thread takes through stage A per round
constexpr int kSlots = kPerThread * kThreads;        // agents per round and CTA
#ifndef SEIR_QUEUE_DIV
#define SEIR_QUEUE_DIV 1   // (tests: shrink the queues so that their overflow paths run on small populations)
#endif
constexpr int kConCap = kSlots / SEIR_QUEUE_DIV;                      // contact queue entries per round
constexpr int kHitCap = kSlots / 2 / SEIR_QUEUE_DIV;                  // hit queue entries per round
constexpr int kDrawCap = 2 * kSlots / SEIR_QUEUE_DIV;                 // draw-item queue entries per round
#ifndef SEIR_DYNAMIC_MAP
#define SEIR_DYNAMIC_MAP 1   // replica ensembles: CTAs are dealt to the replicas anew in every pass, in proportion to their lists
#endif
#ifndef SEIR_PREFETCH
#define SEIR_PREFETCH 0       // 1: stage A asks the L2 for the records of the group kWarps grabs ahead
#endif
#ifndef SEIR_QUEUE_DRAWS
#define SEIR_QUEUE_DRAWS 1    // 1: household draw blocks / kept-contact counts run in stage D; 0: in place in stage A
#endif
constexpr int kWarps = kThreads / 32;
#ifndef SEIR_SPARSE_THREADS
#define SEIR_SPARSE_THREADS 512
#endif
constexpr int kSparseThreads = SEIR_SPARSE_THREADS;  // threads per CTA of seir_sparse_kernel (one CTA per SM: 128 registers per thread)
constexpr int kOutOld = kSlots;                      // staged survivors that may transmit tomorrow
constexpr int kOutNew = kSlots;                      // staged new infectees
static_assert(kSlots <= 65536, "slot ids travel in 16 bits");
struct RepView {
    uint32_t *inbox;
    unsigned long long *winner;
    uint16_t *xday;
    AgentRec *rec;
    const uint32_t *home;
    uint32_t *next_old, *next_new;          // tomorrow's lists
    uint32_t *next_old_cnt, *next_new_cnt;
    uint32_t *pk_cnt;                       // [T+2] entries of the day buckets (row 2 of list_cnt)
    uint32_t *pk_pool;
    uint32_t n_pad_m1;
    uint64_t seed;
    bool home_on;
    int phase;
    int rep;
    bool tracing;                           // contact_tracing is True today (:241-242)
    int T;                                  // the replica's horizon (seir_overrides.T)
    bool sharded;                           // this pass exchanges with the peer GPUs (each steps the agents it owns)
    bool replicated;                        // a sharded run before its switch: every GPU steps everybody, counts its own
    const double *p_ms, *p_sc, *p_cd;       // the replica's severity tables (its table set)
    const double *p_contact;                // -> p_infect_given_contact of the set (:146; read by the replay mode only)
};

// draw-queue items: slot<<16 | kind<<14 | payload
enum { DQ_HH = 0,      // payload = jb<<2 | two S-mask bits: household members 2jb, 2jb+1 (draw block jb of site HH)
       DQ_NATIVE = 1,  // the sender's first Knuth uniform did not end the product: finish the kept-contact count
       DQ_REPLAY = 2 };// the sender's 101 Poisson draws of the replay mode

struct __align__(16) PassSmem {
    RepView v;                     // the replica this CTA serves in the current pass (uniform per CTA)
    uint32_t out_old[kOutOld];
    uint32_t pk_agent[kSlots];     // agents parked this round ...
    uint16_t pk_meta[kSlots];      // ... (wake day - t - 1) << 11 | rank among the CTA's parks for that day
    uint32_t out_new[kOutNew];
    uint32_t hit_c[kHitCap];       // infectee
    uint32_t hit_meta[kHitCap];    // slot<<16 | ord
    uint8_t hit_age[kHitCap];      // age of the infectee, 127 = look it up
    uint32_t conq[kConCap];        // slot<<16 | payload (native: k; replay: age<<8 | j)
    uint32_t drawq[kDrawCap];
    uint32_t slot_agent[kSlots];
    uint32_t hit_n[kSlots];        // hits of the slot's agent this round: total | outside<<16
    uint16_t slot_info[kSlots];    // age | E<<7 | household position<<8
    uint32_t tally[TAL_ROWS * kAges];
    // small per-age tables staged once per kernel (the grid barrier's acquire loads keep flushing L1)
    double iso_asym[kAges], iso_mean[kAges];
    double nat_enlam[2 * kAges * 2];
    int32_t grp_off[kAges + 1];
    uint32_t cnt[4];
    RepParams rp;                  // the replica's overrides, read once per CTA (and again only when the CTA changes replica)
    int rp_rep;                    // the replica sm.rp belongs to (-1: none yet)
    uint32_t list_n[3];            // today's list lengths (new, transmitting, quarantined), read once per CTA
    uint32_t run_cnt[3];           // this CTA's share of the run counters (persistent kernel: flushed once, at the end)
    int dyn[3];                    // replica ensembles: the replica this CTA serves in the current pass, its part, the replica's CTAs
    uint32_t pk_n[32], pk_base[32]; // parks of the round per wake-day offset; their first slot in the day's bucket
    uint32_t n_old, n_pk, n_new, n_hit, n_con, n_draw, base_old, base_new, next_group;
    uint32_t own_n[kMaxShards], own_base[kMaxShards];  // a sharded flush: staged new infectees per owner, their first slot in the owner's list
};

__device__ __forceinline__ void stage_tables(const DevCtx &cx, PassSmem &sm)
{
    for (int k = threadIdx.x; k < kAges; k += kThreads) {
        sm.iso_asym[k] = cx.iso_asym[k];
        sm.iso_mean[k] = cx.iso_mean[k];
    }
    for (int k = threadIdx.x; k <= kAges; k += kThreads) sm.grp_off[k] = cx.grp_off[k];
    // (nat_enlam belongs to the replica's table set: staged by the pass)
    if (threadIdx.x == 0) sm.rp_rep = -1;
    __syncthreads();
}

__device__ __forceinline__ bool at_home(const RepView &v, int64_t c)
{
    return v.home_on && ((__ldg(v.home + (c >> 5)) >> (c & 31)) & 1u);
}

// threads per item when n items are dealt to a CTA: one item per warp while they fit, then 2, 4, ... per warp
__device__ __forceinline__ uint32_t spread_of(uint32_t n, uint32_t threads)
{
    uint32_t s = 32;
    while (s > 1 && n * s > threads) s >>= 1;
    return s;
}
// f(q) for q in [0, n), item q on thread (q mod (kThreads/S)) * S
template <class F> __device__ __forceinline__ void for_items(uint32_t n, F f)
{
    const uint32_t S = spread_of(n, kThreads), per = kThreads / S;
    const bool mine = (threadIdx.x & (S - 1u)) == 0u;
    for (uint32_t q = threadIdx.x / S; q < n; q += per)
        if (mine) f(q);
}

// Everything below exists twice: Pass<false> for runs in which contact tracing can never switch on (no
// outside-infection log, no root bitmap -- and none of their registers), Pass<true> for the others.
template <bool TR> struct Pass {
// owner of agent c in a sharded run
static __device__ __forceinline__ int owner_of(const DevCtx &cx, uint32_t c)
{
    int g = 0;
    while (g + 1 < cx.n_shards && c >= cx.shard_lo[g + 1]) ++g;
    return g;
}

// ---- stage H: one successful transmission slot's agent -> c on day t (:347-359 / :376-390) ----------
static __device__ __noinline__ void process_hit(const DevCtx &cx, PassSmem &sm, uint32_t c, uint32_t meta,
                                            int age_c, int t)
{
    const RepView &v = sm.v;
    const uint32_t slot = meta >> 16, ord = meta & 0xFFFFu;
    if (!SEIR_CHECK(slot < (uint32_t)kSlots && (int64_t)c < cx.n, cx, v.rep)) return;
    const uint32_t i = sm.slot_agent[slot];
    const unsigned long long key = ((unsigned long long)t << 48) | ((unsigned long long)i << 16) | ord;
    if (age_c == 127) age_c = __ldg(cx.stat + c) & 127;  // (independent of the atomic below: one round trip)
    const bool sh = v.sharded;
    const int g = sh ? owner_of(cx, c) : 0;
    const bool remote = sh && g != cx.my_shard;
    // (sharded: every atomic on an infectee is system-scope -- it is performed at the owner's L2 over NVLink)
    const unsigned long long old = sh ? atomicMax_system(cx.peer_winner[g] + c, key) : atomicMax(v.winner + c, key);
    const double tiso = exp_days_ni(sd_lane(draw_block(v.seed, i, (uint32_t)t, SD_SITE_INF, ord << 8), 0), sm.iso_asym[age_c]);
    uint32_t bits = 0;
    if (wday(old) != t) {  // first hit of c today: announce it, queue it for tomorrow
        // (a remote infectee's "infected" bit is set by its owner when it builds the record: one NVLink atomic less per hit)
        if (!remote) bits = inbox_infected_bit((int)(c & 15));
        // c is susceptible for the last day: from tomorrow on nobody's contact test sees it as S (this GPU's copy; a sharded
        // run brings every GPU's copy up to date from the owners' lists of new infectees in the morning, xday_refresh)
        if (cx.n_shards > 1) __stcg(v.xday + c, (uint16_t)(t + 1));
        // staged in shared memory whoever owns c: a sharded run's flush reserves the room in each owner's list of tomorrow
        // with ONE system-scope atomic per CTA and owner (flush_new_sharded)
        const uint32_t pos = atomicAdd(&sm.n_new, 1u);
        if (pos < (uint32_t)kOutNew) {
            sm.out_new[pos] = c;
        } else if (remote) {  // staging full: append to the OWNER's list of tomorrow, in its memory
            const size_t lrow = (size_t)((t + 1) & 1) * cx.n_pad;
            const uint32_t p2 = atomicAdd_system(cx.peer_list_cnt[g] + (cx.T + 2) + (t + 1), 1u);
            if (SEIR_CHECK((int64_t)p2 < cx.n_pad, cx, v.rep)) __stcg(cx.peer_list_new[g] + lrow + p2, c);
        } else {
            v.next_new[atomicAdd_system(v.next_new_cnt, 1u)] = c;
        }
    }
    if (tiso == 0.0) bits |= inbox_q_bit((int)(c & 15));
    if (bits) {
        if (sh) atomicOr_system(cx.peer_inbox[g] + (c >> 4), bits);
        else atomicOr(v.inbox + (c >> 4), bits);
    }
    atomicAdd(&sm.hit_n[slot], ord >= 16u ? 0x10001u : 1u);  // ord >= 16: community (:386-388)
    if (TR && cx.log_ent && ord >= 16u) {  // infected_by[i, k] = c (:387); k follows from (day, ordinal) when the log is read
        const uint32_t e = atomicAdd(cx.log_cnt + v.rep, 1u);
        if (e < cx.log_cap) {
            const uint32_t prev = atomicExch(cx.log_head + (size_t)v.rep * cx.n_pad + i, e + 1u);
            __stcg(cx.log_ent + (size_t)v.rep * cx.log_cap + e, make_uint4(c, prev, ((uint32_t)t << 16) | ord, 0u));
        } else {
            atomicOr(&cx.counters[v.rep * 4 + 0], 1ull);  // log overflow: seir_run reports it
        }
    }
}

static __device__ __forceinline__ void push_hit(const DevCtx &cx, PassSmem &sm, uint32_t c, uint32_t meta,
                                         int age_c, int t)
{
    const uint32_t pos = atomicAdd(&sm.n_hit, 1u);
    if (pos < (uint32_t)kHitCap) {
        sm.hit_c[pos] = c;
        sm.hit_meta[pos] = meta;
        sm.hit_age[pos] = (uint8_t)age_c;
    } else {
        process_hit(cx, sm, c, meta, age_c, t);  // queue full: handle in place
    }
}

// ---- stage C: one candidate contact of slot's agent (:373-375) ---------------------------------------
static __device__ __noinline__ void process_contact(const DevCtx &cx, PassSmem &sm, uint32_t item, int t)
{
    const RepView &v = sm.v;
    const uint32_t slot = item >> 16, payload = item & 0xFFFFu;
    const uint32_t i = sm.slot_agent[slot];
    const uint32_t info = sm.slot_info[slot];
    int a;
    uint32_t ord;
    sd_block b;
    if (cx.mode == SEIR_MODE_NATIVE) {
        b = draw_block(v.seed, i, (uint32_t)t, SD_SITE_NPICK, payload);
        // contact age: one alias-table look-up in the sender's row (same arithmetic as sd_alias_pick)
        const double x = SD_MUL(sd_lane(b, 0), (double)kAges);
        int col = (int)x;
        if (col > kAges - 1) col = kAges - 1;
        const int row = (int)(info & 127u) * kAges + col;
        const double pr = __ldg(cx.nat_alias_prob + row);
        const int al = __ldg(cx.nat_alias_idx + row);
        a = (SD_SUB(x, (double)col) < pr) ? col : al;
        ord = SD_ORD_COMM_NATIVE(payload);
    } else {
        a = (int)(payload >> 8);
        const uint32_t j = payload & 255u;
        b = draw_block(v.seed, i, (uint32_t)t, SD_SITE_KEEP, payload);
        double p_c = __ldg(v.p_contact);
        if (info & 128u) p_c = SD_MUL(p_c, cx.asym);
        if (!(sd_lane(b, 0) < p_c)) return;  // :373
        ord = SD_ORD_COMM_REPLAY(a, j);
    }
    const int g0 = sm.grp_off[a], glen = sm.grp_off[a + 1] - g0;
    if (glen == 0) return;
    const uint32_t c = (uint32_t)__ldg(cx.grp_members + g0 + (int64_t)sd_index(b, (uint64_t)glen));  // :374
    if (!SEIR_CHECK((int64_t)c < cx.n && sd_index(b, (uint64_t)glen) < (uint64_t)glen, cx, v.rep)) return;
    bool s_c;  // S[t-1, c]
    if (cx.n_shards > 1) {  // a sharded run (replicated days included): this GPU's copy of the infection days
        const uint32_t xd = __ldcg(v.xday + c);
        s_c = xd == 0u || xd == (uint32_t)t + 1u;
    } else {
        const unsigned long long w = __ldcg(v.winner + c);
        s_c = w == 0ull || wday(w) == t;
    }
    if (s_c && !at_home(v, c))  // :375
        push_hit(cx, sm, c, (slot << 16) | ord, a, t);
}

static __device__ __forceinline__ void push_contact(const DevCtx &cx, PassSmem &sm, uint32_t item, int t)
{
    const uint32_t pos = atomicAdd(&sm.n_con, 1u);
    if (pos < (uint32_t)kConCap) sm.conq[pos] = item;
    else process_contact(cx, sm, item, t);
}

// ---- stage D: one draw item ---------------------------------------------------------------------------
static __device__ __noinline__ void process_draw_item(const DevCtx &cx, PassSmem &sm, uint32_t item, int t)
{
    const RepView &v = sm.v;
    const uint32_t slot = item >> 16, kind = (item >> 14) & 3u, payload = item & 0x3FFFu;
    const uint32_t i = sm.slot_agent[slot];
    const uint32_t info = sm.slot_info[slot];
    const bool Ep = (info & 128u) != 0u;
    const int age = (int)(info & 127u);
    if (kind == DQ_HH) {  // :346 for household members 2jb, 2jb+1
        const uint32_t jb = payload >> 2, bits = payload & 3u;
        const int pos = (int)((info >> 8) & 7u);
        const uint32_t base = i - (uint32_t)pos;
        double p_h = cx.p_hh_vec ? __ldg(cx.p_hh_vec + i) : cx.p_hh_scalar;
        if (Ep) p_h = SD_MUL(p_h, cx.asym);
        const sd_block hb = draw_block(v.seed, i, (uint32_t)t, SD_SITE_HH, jb);
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int j = 2 * (int)jb + l;
            if (((bits >> l) & 1u) && sd_lane(hb, l) < p_h)
                push_hit(cx, sm, base + j + (j >= pos ? 1 : 0), (slot << 16) | SD_ORD_HH(j), 127, t);
        }
    } else if (kind == DQ_NATIVE) {  // sd_poisson_day past its first uniform (lane 1 of the day block)
        const double en = sm.nat_enlam[(v.phase * kAges + age) * 2 + (Ep ? 1 : 0)];
        const sd_block db = draw_block(v.seed, i, (uint32_t)t, SD_SITE_DAY, 0);
        const uint32_t kept = poisson_day_more(v.seed, i, (uint32_t)t, sd_lane(db, 1), en, SD_MAX_CONTACTS_NATIVE);
        for (uint32_t k = 0; k < kept; ++k) push_contact(cx, sm, (slot << 16) | k, t);
    } else {  // :367-372, one Poisson per non-empty contact age
        const double *en = cx.enlam + ((size_t)v.phase * kAges + age) * kAges;
#pragma unroll 1
        for (int a = 0; a < kAges; ++a) {
            if (sm.grp_off[a + 1] == sm.grp_off[a]) continue;  // :368-369
            const uint32_t nc = poisson_ni(v.seed, i, (uint32_t)t, SD_SITE_POIS, (uint32_t)a << 8, __ldg(en + a),
                                           SD_MAX_CONTACTS_PER_AGE);
            for (uint32_t j = 0; j < nc; ++j) push_contact(cx, sm, (slot << 16) | ((uint32_t)a << 8) | j, t);
        }
    }
}

static __device__ __forceinline__ void push_draw(const DevCtx &cx, PassSmem &sm, uint32_t item, int t)
{
#if SEIR_QUEUE_DRAWS
    const uint32_t pos = atomicAdd(&sm.n_draw, 1u);
    if (pos < (uint32_t)kDrawCap) sm.drawq[pos] = item;
    else process_draw_item(cx, sm, item, t);
#else
    process_draw_item(cx, sm, item, t);
#endif
}

// (tw: 1, or 0 for an agent another GPU owns while a sharded run is still replicated -- it is stepped here too, but
// counted by its owner only)
static __device__ __forceinline__ void build_new_record(const DevCtx &cx, PassSmem &sm, uint32_t i, int t, AgentRec &r, uint32_t tw)
{
    const RepView &v = sm.v;
    // infected on day t-1 by the winner of that day's hits (:347-356)
    const unsigned long long w = __ldcg(v.winner + i);
    const uint32_t ib = __ldcg(v.inbox + (i >> 4));
    const uint16_t stt = __ldg(cx.stat + i);
    if (v.sharded) atomicOr(v.inbox + (i >> 4), inbox_infected_bit((int)(i & 15)));  // (a hitter on another GPU left it to the owner)
    const uint4 hint = __ldcg(reinterpret_cast<const uint4 *>(v.rec + i));  // the timers a hitter of a warp-cooperative pass left (sparse_entry)
    const uint32_t inf = (uint32_t)(w >> 16), ord = (uint32_t)(w & 0xFFFFu);
    const int tx = t - 1;
    uint32_t tiso16, tact16;
    if (hint.x == (uint32_t)w && hint.y == (uint32_t)((w & ~W_TRACED) >> 32) && hint.w == cx.run_nonce) {
        tiso16 = hint.z & 0xFFFFu;
        tact16 = hint.z >> 16;
    } else {
        tiso16 = sd_enc16(exp_days_ni(sd_lane(draw_block(v.seed, inf, (uint32_t)tx, SD_SITE_INF, ord << 8), 0), sm.iso_asym[stt & 127]));
        tact16 = sd_enc16(lognormal_days_ni(v.seed, inf, (uint32_t)tx, SD_SITE_INF, ord << 8, cx.act_mu, cx.act_sigma));
    }
    r.clear();
    r.set_t_exposed((uint16_t)tx);
    r.set_tt_isolate(tiso16);
    r.set_tt_activation(tact16);
    r.set_t_end(NEVER);
    r.set_t_q_on(NEVER);
    r.set_stat(stt);
    r.set_state(C_E);
    if (ib & inbox_q_bit((int)(i & 15))) {
        r.set_state(r.state() | ST_Q);
        r.set_t_q_on((uint16_t)tx);
        atomicAdd(&sm.tally[TAL_NEW_QE * kAges + (stt & 127)], tw);
    }
    if (w & W_TRACED) {  // a tracer reached the agent on the day it was infected (:83-87): Documented, Q (inbox bit), traced
        r.set_state(r.state() | ST_DOC);
        r.set_t_documented((uint16_t)tx);
    }
    atomicAdd(&sm.tally[TAL_NEW_E * kAges + (stt & 127)], tw);
}

// the progression events of the agent's own day (:256-272, :291-311, :312-321)
static __device__ __forceinline__ void progression_events(const DevCtx &cx, const PassSmem &sm, uint64_t seed, uint32_t i, int t,
                                                AgentRec &r, int comp, int &ncomp, bool &nq, bool &ndoc, bool &root)
{
    const uint16_t stt = r.stat();
    const int age = stt & 127;
    const int k3 = age * 4 + ((stt >> 7) & 1) * 2 + ((stt >> 8) & 1);
    if (comp == C_E) {  // :256-272 activation
        ncomp = C_MILD;
        r.set_flags(r.flags() | F_MILD);
        sd_block b = draw_block(seed, i, (uint32_t)t, SD_SITE_ACT, 0);
        if (sd_lane(b, 0) < __ldg(sm.v.p_ms + k3)) {
            r.set_tt_severe(sd_enc16(exp_days_ni(sd_lane(b, 1), cx.m_sev)));
            r.set_tt_recovery(NEVER);
        } else {
            r.set_tt_recovery(sd_enc16(exp_days_ni(sd_lane(b, 1), cx.m_mild_rec)));
            r.set_tt_severe(NEVER);
        }
        const double tiso = exp_days_ni(sd_lane(draw_block(seed, i, (uint32_t)t, SD_SITE_ACT, 1), 0), sm.iso_mean[age]);
        r.set_tt_isolate(sd_enc16(tiso));
        if (tiso == 0.0) nq = true;
    } else if (comp == C_MILD) {  // :291-311 Mild -> Severe
        ncomp = C_SEV;
        r.set_flags(r.flags() | F_SEV);
        if (!ndoc) {  // :295-303 severe cases are always documented (and start a trace)
            ndoc = true;
            root = true;
            r.set_t_documented((uint16_t)t);
        }
        nq = true;
        sd_block b = draw_block(seed, i, (uint32_t)t, SD_SITE_SEV, 0);
        if (sd_lane(b, 0) < __ldg(sm.v.p_sc + k3)) {
            r.set_tt_critical(sd_enc16(exp_days_ni(sd_lane(b, 1), cx.m_crit)));
            r.set_tt_recovery(NEVER);
        } else {
            r.set_tt_recovery(add_sat16(sd_enc16(exp_days_ni(sd_lane(b, 1), cx.m_sev_rec)), r.tt_severe()));
            r.set_tt_critical(NEVER);
        }
    } else {  // :312-321 Severe -> Critical
        ncomp = C_CRIT;
        r.set_flags(r.flags() | F_CRIT);
        sd_block b = draw_block(seed, i, (uint32_t)t, SD_SITE_CRIT, 0);
        if (sd_lane(b, 0) < __ldg(sm.v.p_cd + k3)) {
            r.set_tt_death(sd_enc16(exp_days_ni(sd_lane(b, 1), cx.m_death)));
            r.set_tt_recovery(NEVER);
        } else {
            r.set_tt_recovery(add_sat16(add_sat16(sd_enc16(exp_days_ni(sd_lane(b, 1), cx.m_crit_rec)), r.tt_severe()), r.tt_critical()));
            r.set_tt_death(NEVER);
        }
    }
}

// The next day after t on which something can happen to a QUARANTINED agent (it neither transmits nor enters
// isolation, :328-337 are guarded by not Q[t-1]): one of its `t - t0 == draw` tests (:257, :276, :291, :312, :324)
// fires, or -- Mild and not Documented -- its daily documentation draw (:280-289) succeeds.  The draws are keyed on
// (agent, day), so the days ahead can be evaluated now: at most kDocScan of them, after which the agent is looked at
// again.  0x7FFFFFFF: nothing can ever happen (a draw of 0 froze the agent).
// (out of line, the record's six day fields by value: the scan loop must not cost stage A its registers)
static __device__ __noinline__ int next_wake_day(const DevCtx &cx, uint64_t seed, uint32_t i, uint32_t w0, uint32_t w1, uint32_t w2,
                                                 int comp, bool doc, int t)
{
    const int t_exposed = (int)(w0 & 0xFFFFu);
    const uint16_t tt_activation = (uint16_t)(w0 >> 16), tt_severe = (uint16_t)(w1 & 0xFFFFu), tt_recovery = (uint16_t)(w1 >> 16),
                   tt_critical = (uint16_t)(w2 & 0xFFFFu), tt_death = (uint16_t)(w2 >> 16);
    int w = 0x7FFFFFFF;
#define SEIR_CAND(t0, tt) do { if ((tt) != NEVER) { const int d_ = (int)(t0) + (int)(tt); if (d_ > t && d_ < w) w = d_; } } while (0)
    if (comp == C_E) {
        SEIR_CAND(t_exposed, tt_activation);
    } else {
        const int t_inf = t_exposed + (int)tt_activation;
        SEIR_CAND(t_inf, tt_recovery);
        if (comp == C_MILD) SEIR_CAND(t_inf, tt_severe);
        else if (comp == C_SEV) SEIR_CAND(t_inf + (int)tt_severe, tt_critical);
        else SEIR_CAND(t_inf + (int)tt_severe + (int)tt_critical, tt_death);
    }
#undef SEIR_CAND
    if (comp == C_MILD && !doc && cx.p_doc > 0.0) {
        const int hi = w < t + kDocScan + 1 ? w : t + kDocScan + 1;  // the draws of days t+1 .. hi-1; day hi is visited anyway
        int first = hi;
#pragma unroll
        for (int k = kDocScan; k >= 1; --k) {  // (all blocks, no early exit: the Philox rounds of the independent days interleave)
            const sd_block b = sd_draw(seed, i, (uint32_t)(t + k), SD_SITE_DAY, 0);
            if (t + k < hi && sd_lane(b, 0) < cx.p_doc) first = t + k;
        }
        w = first;
    }
    return w;
}

// ---- the agent's own day (:256-337) and the set-up of its transmission (:339-390) ---------------------
// Returns DAY_KEEP if the agent may transmit tomorrow, DAY_KEEP_Q if it stays active and quarantined (then `wake`
// is the next day it has to be looked at), DAY_DROP if it left the active compartments or is frozen for good.
// (DAY_KEEP_Q carries the wake day's offset: DAY_KEEP_Q | (wake - t - 1) << 2)
enum { DAY_DROP = 0, DAY_KEEP = 1, DAY_KEEP_Q = 3 };
static __device__ __forceinline__ int agent_day(const DevCtx &cx, PassSmem &sm, uint32_t slot, uint32_t i, AgentRec &r,
                                         bool dirty, int t, uint32_t tw)
{
    const RepView &v = sm.v;
    const uint16_t stt = r.stat();
    const int age = stt & 127;
    const uint8_t st = r.state();
    const int comp = st & 7;
    const bool Ep = comp == C_E, Mp = comp == C_MILD, Sevp = comp == C_SEV, Cp = comp == C_CRIT;
    const bool Qp = (st & ST_Q) != 0, Dp = (st & ST_DOC) != 0;
    int ncomp = comp;
    bool nq = Qp, ndoc = Dp;
    bool transmit = !Qp;
    bool root = false;  // newly documented today by its own progression: the root of a trace (:284-289, :298-303)
    // the agent's day block (lane 0: documentation draw, lane 1: first uniform of the kept-contact count), drawn ONCE, here,
    // for every lane of the warp that may need it -- the two users sit in different branches below, and a warp would
    // otherwise run the Philox rounds once per branch
    const bool want_doc = Mp && !Dp && cx.p_doc > 0.0;
    const bool have_db = want_doc || (!Qp && cx.mode == SEIR_MODE_NATIVE);
    sd_block db;
    db.w[0] = db.w[1] = db.w[2] = db.w[3] = 0u;
    SEIR_FINE(3);
    if (have_db) db = draw_block(v.seed, i, (uint32_t)t, SD_SITE_DAY, 0);
    SEIR_FINE(4);
    const int t_inf = t_infected_of(r);
    if (!Ep && due(t, t_inf, r.tt_recovery())) {  // :276-279
        ncomp = C_R;
        nq = false;
        transmit = false;
        atomicAdd(&sm.tally[TAL_NEW_R * kAges + age], tw);
    } else {
        const bool ev = (Ep && due(t, r.t_exposed(), r.tt_activation())) || (Mp && due(t, t_inf, r.tt_severe())) ||
                        (Sevp && due(t, t_severe_of(r), r.tt_critical()));
        if (want_doc) {  // :280-289 (lane 0 of the agent's day block)
            if (sd_lane(db, 0) < cx.p_doc) {
                ndoc = true;
                root = true;
                if (TR && v.tracing) nq = true;  // :287-288
                r.set_t_documented((uint16_t)t);
            }
        }
        if (ev) {
            progression_events(cx, sm, v.seed, i, t, r, comp, ncomp, nq, ndoc, root);
            dirty = true;
        } else if (Cp && due(t, t_critical_of(r), r.tt_death())) {  // :323-327
            ncomp = C_D;
            nq = false;
            atomicAdd(&sm.tally[TAL_NEW_D * kAges + age], tw);
        }
        if (ndoc && !Dp) atomicAdd(&sm.tally[TAL_NEW_DOC * kAges + age], tw);
        if (TR && root && v.tracing) {
            atomicOr(cx.rootbits + (size_t)v.rep * (cx.n_pad / 32) + (i >> 5), 1u << (i & 31));
            atomicOr(cx.rootsum + (size_t)v.rep * (cx.n_pad / 1024) + (i >> 10), 1u << ((i >> 5) & 31));
            atomicOr(cx.rootsum2 + (size_t)v.rep * ((cx.n_pad + 32767) / 32768) + (i >> 15), 1u << ((i >> 10) & 31));
            atomicAdd(cx.root_cnt + (size_t)v.rep * (cx.T + 2) + t, 1u);
        }
    }
    // :328-337 isolation entry (time_to_isolate may just have been redrawn at :270)
    if (transmit) {
        if (!Ep && due(t, t_infected_of(r), r.tt_isolate())) { nq = true; transmit = false; }
        else if (Ep && due(t, r.t_exposed(), r.tt_isolate())) { nq = true; transmit = false; }
    }
    // the day's changes of the head counts (every tally row is a change: nobody recounts the agents that were not visited)
    const bool active = ncomp <= C_CRIT;  // (ncomp >= C_E always)
    if (ncomp != comp) {
        atomicAdd(&sm.tally[(TAL_D_E + comp - C_E) * kAges + age], 0u - tw);
        if (active) atomicAdd(&sm.tally[(TAL_D_E + ncomp - C_E) * kAges + age], tw);
    }
    if ((nq && active) != Qp) atomicAdd(&sm.tally[TAL_D_Q * kAges + age], Qp ? 0u - tw : tw);
    // the agent's own record is final for today except for the infector counters (stage W)
    const uint8_t nst = (uint8_t)(ncomp | (nq ? ST_Q : 0) | (ndoc ? ST_DOC : 0));
    if (nst != st) {
        r.set_state(nst);
        if (nq && !Qp && r.t_q_on() == NEVER) r.set_t_q_on((uint16_t)t);
        if (ncomp == C_R || ncomp == C_D) r.set_t_end((uint16_t)t);
        dirty = true;
    }
    if (dirty) store_rec(v.rec + i, r);
    SEIR_FINE(5);

    if (transmit) {
        // ---- household push :339-359; households are contiguous ranges [i-pos, i-pos+size)
        const int pos = (stt >> 9) & 7, size = ((stt >> 12) & 7) + 1;
        const uint32_t base = i - (uint32_t)pos;
        uint32_t smask = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) {  // independent loads: one round trip for the household
            if (j < size - 1) {
                const unsigned long long w = __ldcg(v.winner + base + j + (j >= pos ? 1 : 0));
                if (w == 0ull || wday(w) == t) smask |= 1u << j;  // S[t-1,c] (:346)
            }
        }
        SEIR_FINE(6);
        // ---- community push :361-390: the first Knuth uniform of the kept-contact count is lane 1 of the day block
        const bool out = !at_home(v, i);
        if (out && cx.mode == SEIR_MODE_NATIVE) {
            const double en = sm.nat_enlam[(v.phase * kAges + age) * 2 + (Ep ? 1 : 0)];
            if (sd_lane(db, 1) > en) push_draw(cx, sm, (slot << 16) | (DQ_NATIVE << 14), t);
        } else if (out) {
            push_draw(cx, sm, (slot << 16) | (DQ_REPLAY << 14), t);
        }
#pragma unroll
        for (uint32_t jb = 0; jb < 4; ++jb) {
            const uint32_t bits = (smask >> (2 * jb)) & 3u;
            if (bits) push_draw(cx, sm, (slot << 16) | (DQ_HH << 14) | (jb << 2) | bits, t);
        }
    }
    SEIR_FINE(7);
    if (!active) return DAY_DROP;
#if SEIR_EVE_PARK
    // An agent whose activation (:256-272) or Mild -> Severe step (:291-311) is due TOMORROW goes through tomorrow's day
    // bucket instead of the transmitting list: there every lane of its warp has an event, while among 32 transmitting
    // agents two or three have one and the whole warp walks through their draws (two Philox blocks, three exponential
    // times) at a tenth of its lanes.  Where an agent is listed changes nothing about what happens to it.
    if (!nq && t + 1 < v.T &&
        ((ncomp == C_E && due(t + 1, r.t_exposed(), r.tt_activation())) || (ncomp == C_MILD && due(t + 1, t_infected_of(r), r.tt_severe()))))
        return DAY_KEEP_Q;  // (wake day offset 0)
#endif
    if (!nq) return DAY_KEEP;
    // An agent that turned quarantined TODAY sits in a warp of transmitting agents: it is looked at again tomorrow, among
    // the parked agents due that day -- warps in which every lane works out its wake day (the scan of the documentation
    // draws is the longest path of this function, and one lane taking it would hold its whole warp back).
    // (Without documentation draws the wake day is a handful of compares: computed at once.)
    int wake = t + 1;
    if (Qp || !(ncomp == C_MILD && !ndoc && cx.p_doc > 0.0)) wake = next_wake_day(cx, v.seed, i, r.w[0], r.w[1], r.w[2], ncomp, ndoc, t);
    if (wake >= v.T) return DAY_DROP;  // nothing more can happen to it within the horizon: its record is final
    if (wake > t + kParkHorizon) wake = t + kParkHorizon;  // (the bucket offset travels in 5 bits: look again then)
    return DAY_KEEP_Q | ((wake - t - 1) << 2);
}

// ---- stage A: one list entry -----------------------------------------------------------------------------
static __device__ __forceinline__ int stage_a_entry(const DevCtx &cx, PassSmem &sm, uint32_t slot, uint32_t i,
                                             bool is_new, int t, bool finalize, bool own)
{
    const RepView &v = sm.v;
    AgentRec r;
    bool dirty = false;
    const uint32_t tw = own ? 1u : 0u;
    if (is_new) {
        build_new_record(cx, sm, i, t, r, tw);
        dirty = true;
    } else {
        r = load_rec(v.rec + i);
    }
    SEIR_FINE(2);
    const uint16_t stt = r.stat();
    const int age = stt & 127;
    const int comp = r.state() & 7;
    sm.slot_agent[slot] = i;
    sm.slot_info[slot] = (uint16_t)(age | (comp == C_E ? 128 : 0) | (((stt >> 9) & 7) << 8) | (own ? 0 : 0x8000));
    if (finalize) {  // pass T: nothing moves (only the agents infected on the last day are listed: their records)
        if (dirty) store_rec(v.rec + i, r);
        return DAY_DROP;
    }
    return agent_day(cx, sm, slot, i, r, dirty, t, tw);
}

// Survivors: the ones that may transmit tomorrow re-append themselves to tomorrow's list (one shared-memory atomic
// per warp); the quarantined ones are staged for the bucket of their wake day -- `rank` numbers the CTA's parks of a day.
static __device__ __forceinline__ void append_survivors(PassSmem &sm, int keep, uint32_t i)
{
    const RepView &v = sm.v;
    const int lane = threadIdx.x & 31;
    const unsigned mt = __ballot_sync(0xFFFFFFFFu, keep == DAY_KEEP);
    if (mt) {
        uint32_t base = 0;
        if (lane == __ffs(mt) - 1) base = atomicAdd(&sm.n_old, (uint32_t)__popc(mt));
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(mt) - 1);
        if (keep == DAY_KEEP) {
            const uint32_t pos = base + __popc(mt & ((1u << lane) - 1u));
            if (pos < (uint32_t)kOutOld) sm.out_old[pos] = i;
            else v.next_old[atomicAdd(v.next_old_cnt, 1u)] = i;
        }
    }
    const unsigned mq = __ballot_sync(0xFFFFFFFFu, (keep & 3) == DAY_KEEP_Q);
    if (mq) {  // one shared-memory atomic per warp for the staging slot, one per warp and wake day for the rank
        uint32_t base = 0;
        if (lane == __ffs(mq) - 1) base = atomicAdd(&sm.n_pk, (uint32_t)__popc(mq));  // (at most one park per slot and round: < kSlots)
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(mq) - 1);
        if ((keep & 3) == DAY_KEEP_Q) {
            const uint32_t off = (uint32_t)keep >> 2;                // wake day - t - 1: 0 .. kParkHorizon - 1
            const unsigned same = __match_any_sync(mq, off);
            uint32_t r0 = 0;
            if (lane == __ffs(same) - 1) r0 = atomicAdd(&sm.pk_n[off], (uint32_t)__popc(same));
            r0 = __shfl_sync(same, r0, __ffs(same) - 1);
            const uint32_t k = base + __popc(mq & ((1u << lane) - 1u));
            sm.pk_agent[k] = i;
            sm.pk_meta[k] = (uint16_t)((off << 11) | (r0 + __popc(same & ((1u << lane) - 1u))));
        }
    }
}

// right after stage A, when the round's survivors are staged: thread 0 reserves their room in tomorrow's list.  The
// returned offset stays in a register until the flush, so the round trip to L2 overlaps the stages in between
// instead of standing in front of the grid barrier.
static __device__ __forceinline__ uint32_t reserve_survivors(PassSmem &sm)
{
    const RepView &v = sm.v;
    uint32_t base = 0;
    if (threadIdx.x == 0) {
        const uint32_t so = sm.n_old < (uint32_t)kOutOld ? sm.n_old : (uint32_t)kOutOld;
        if (so) base = atomicAdd(v.next_old_cnt, so);
    }
    return base;
}
// (the same for the round's parks: threads 128 .. 142, one wake day each)
static __device__ __forceinline__ uint32_t reserve_parks(PassSmem &sm, int t) { return sm.n_pk ? park_reserve(sm, t) : 0u; }

// The round's parks go to their day buckets.  Threads 128 .. 142 own one wake-day offset each and reserve the room
// (one atomic per CTA and day); afterwards every staged agent is stored at its rank behind that base.
static __device__ __forceinline__ uint32_t park_reserve(PassSmem &sm, int t)
{
    const int o = (int)threadIdx.x - 128;
    if (o < 0 || o >= kParkHorizon) return 0u;
    const uint32_t cnt = sm.pk_n[o];
    return cnt ? atomicAdd(sm.v.pk_cnt + (t + 1 + o), cnt) : 0u;
}
static __device__ __forceinline__ void park_store(const DevCtx &cx, PassSmem &sm, int t, uint32_t n_pk)
{
    const RepView &v = sm.v;
    for (uint32_t k = threadIdx.x; k < n_pk; k += kThreads) {
        const uint32_t meta = sm.pk_meta[k], o = meta >> 11, pos = sm.pk_base[o] + (meta & 2047u);
        if (SEIR_CHECK((int64_t)pos < cx.n_pad, cx, v.rep))
            __stcg(v.pk_pool + (size_t)((t + 1 + (int)o) & (kParkRing - 1)) * cx.n_pad + pos, sm.pk_agent[k]);
    }
}

// the staged survivors, parks and (flush_new) new infectees to tomorrow's lists / the day buckets (the whole CTA;
// ends with the staging counters reset)
static __device__ __forceinline__ void flush_survivors(const DevCtx &cx, PassSmem &sm, int t, bool flush_new, uint32_t reserved,
                                                       uint32_t park_reserved, bool new_reserved = false)
{
    const RepView &v = sm.v;
    const uint32_t so = sm.n_old < (uint32_t)kOutOld ? sm.n_old : (uint32_t)kOutOld;
    const uint32_t sp = sm.n_pk;
    const uint32_t sn = flush_new ? (sm.n_new < (uint32_t)kOutNew ? sm.n_new : (uint32_t)kOutNew) : 0u;
    __syncthreads();
    if (so | sp | sn) {
        if (threadIdx.x == 0 && so) { sm.base_old = reserved; sm.n_old = 0; }
        if (threadIdx.x == 64 && sn && !v.sharded) { sm.base_new = new_reserved ? reserved : atomicAdd_system(v.next_new_cnt, sn); sm.n_new = 0; }
        if (v.sharded && sn) {  // (CTA-uniform) the staged infectees belong to different GPUs: count them per owner, ...
            if (threadIdx.x < kMaxShards) sm.own_n[threadIdx.x] = 0;
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < sn; k += kThreads) atomicAdd(&sm.own_n[owner_of(cx, sm.out_new[k])], 1u);
            __syncthreads();
            // ... reserve the room in every owner's list of tomorrow (one system-scope atomic per CTA and owner), ...
            if (threadIdx.x < (unsigned)cx.n_shards) {
                const uint32_t cnt = sm.own_n[threadIdx.x];
                sm.own_base[threadIdx.x] = cnt ? atomicAdd_system(cx.peer_list_cnt[threadIdx.x] + (cx.T + 2) + (t + 1), cnt) : 0u;
                sm.own_n[threadIdx.x] = 0;
            }
            if (threadIdx.x == 64) sm.n_new = 0;
        }
        if (sp && threadIdx.x >= 128 && threadIdx.x < 128 + kParkHorizon) sm.pk_base[threadIdx.x - 128] = park_reserved;
        __syncthreads();
#ifdef SEIR_CHECKS
        if (threadIdx.x == 0 && ((so && sm.base_old + so > v.n_pad_m1 + 1u) || (sn && !v.sharded && sm.base_new + sn > v.n_pad_m1 + 1u)))
            atomicOr(&sm.cnt[3], 1u);
        __syncthreads();
        if (sm.cnt[3]) return;
#endif
        for (uint32_t k = threadIdx.x; k < so; k += kThreads) v.next_old[sm.base_old + k] = sm.out_old[k];
        if (!v.sharded) {
            for (uint32_t k = threadIdx.x; k < sn; k += kThreads) v.next_new[sm.base_new + k] = sm.out_new[k];
        } else {  // ... and store each infectee behind its owner's base, in the owner's memory
            const size_t lrow = (size_t)((t + 1) & 1) * cx.n_pad;
            for (uint32_t k = threadIdx.x; k < sn; k += kThreads) {
                const uint32_t c = sm.out_new[k];
                const int g = owner_of(cx, c);
                const uint32_t pos = sm.own_base[g] + atomicAdd(&sm.own_n[g], 1u);
                if (SEIR_CHECK((int64_t)pos < cx.n_pad, cx, v.rep)) __stcg(cx.peer_list_new[g] + lrow + pos, c);
            }
        }
        if (sp) park_store(cx, sm, t, sp);
        __syncthreads();
        if (sp) {
            if (threadIdx.x < 32) sm.pk_n[threadIdx.x] = 0;
            if (threadIdx.x == 32) sm.n_pk = 0;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------
// The warp-cooperative pass of a SPARSE day (sparse_entry / sparse_rounds).
// While a day's lists are shorter than the warps that serve them, the staged pass above runs at its latency floor:
// ONE agent's chain through five stages, queues and __syncthreads(), every link a draw or a round trip issued by a
// single lane.  Here a WARP takes one list entry through its whole day, and the 32 lanes hold everything of that day
// that is independent by the keyed-draw contract: the agent's day block, its four household blocks, the blocks of the
// three progression events, four more blocks of the kept-contact count and sixteen contact picks are ONE SIMD Philox
// evaluation issued before the agent's record has arrived; the household members' words, the contacts' alias / member /
// winner gathers and the hits' atomics run side by side on their own lanes; the tries of an infectee's log-normal
// activation time are evaluated at once.  The agent's logic (:256-337) is executed uniformly by all lanes.
// No queue and no CTA barrier; tallies go straight to the day's rows (fire-and-forget reductions).
// Tomorrow's list of transmitting agents keeps the SLOT of every entry (a dropped agent leaves kHole), so a survivor
// costs one store and no atomic; holes disappear whenever a staged pass runs (it re-appends densely).
// Same statements, same keys, same atomics on the infectee as the staged pass: the two are interchangeable day by day
// (tests/test_gpu_parity.py runs every fixture through both).
static __device__ __forceinline__ void tally_add(const DevCtx &cx, const RepView &v, int t, int row, int age, uint32_t val)
{
    const int day = row < kTalLagRows ? t - 1 : t;
    if (val && day < v.T) atomicAdd(cx.tally + ((size_t)v.rep * cx.T + day) * (TAL_ROWS * kAges) + row * kAges + age, val);
}

// next_wake_day with the documentation draws of the days ahead on lanes 1 .. kDocScan (every lane returns the day)
static __device__ __forceinline__ int next_wake_day_warp(const DevCtx &cx, uint64_t seed, uint32_t i, const AgentRec &r, int comp, bool doc, int t)
{
    const int lane = threadIdx.x & 31;
    const int t_exposed = (int)r.t_exposed();
    int w = 0x7FFFFFFF;
#define SEIR_CAND(t0, tt) do { if ((tt) != NEVER) { const int d_ = (int)(t0) + (int)(tt); if (d_ > t && d_ < w) w = d_; } } while (0)
    if (comp == C_E) {
        SEIR_CAND(t_exposed, r.tt_activation());
    } else {
        const int t_inf = t_exposed + (int)r.tt_activation();
        SEIR_CAND(t_inf, r.tt_recovery());
        if (comp == C_MILD) SEIR_CAND(t_inf, r.tt_severe());
        else if (comp == C_SEV) SEIR_CAND(t_inf + (int)r.tt_severe(), r.tt_critical());
        else SEIR_CAND(t_inf + (int)r.tt_severe() + (int)r.tt_critical(), r.tt_death());
    }
#undef SEIR_CAND
    if (comp == C_MILD && !doc && cx.p_doc > 0.0) {
        const int hi = w < t + kDocScan + 1 ? w : t + kDocScan + 1;
        const int k = lane >= 1 && lane <= kDocScan ? lane : 1;
        const sd_block b = draw_block(seed, i, (uint32_t)(t + k), SD_SITE_DAY, 0);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, lane >= 1 && lane <= kDocScan && t + k < hi && sd_lane(b, 0) < cx.p_doc);
        w = m ? t + (__ffs(m) - 1) : hi;
    }
    return w;
}

// One list entry by one warp (all 32 lanes arrive and leave together).  kind: 0 parked agent due today, 1 infected
// yesterday (no record yet), 2 may transmit today.  Returns DAY_DROP / DAY_KEEP / DAY_KEEP_Q | offset like agent_day.
#ifdef SEIR_FINE_STAMPS
#define SEIR_SFINE(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) cx.fine_ns[(size_t)t * 16 + (k)] = (unsigned long long)clock64(); } while (0)  // (SM cycles: deltas within CTA 0)
#else
#define SEIR_SFINE(k) do { } while (0)
#endif
static __device__ __noinline__ int sparse_entry(const DevCtx &cx, PassSmem &sm, uint32_t i, int kind, int t, bool finalize, bool own,
                                                uint32_t site, uint32_t sub)
{
    const RepView &v = sm.v;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tw = own ? 1u : 0u;
    const bool is_new = kind == 1;
    // ---- everything that only needs (i, t) is in flight before the record arrives
    AgentRec r;
    unsigned long long w_self = 0;
    uint32_t ib = 0, stt_new = 0;
    uint4 hint = make_uint4(0u, 0u, 0u, 0u);
    if (is_new) {
        w_self = __ldcg(v.winner + i);
        ib = __ldcg(v.inbox + (i >> 4));
        stt_new = __ldg(cx.stat + i);
        hint = __ldcg(reinterpret_cast<const uint4 *>(v.rec + i));  // the timers its infector may have left (see the hits below)
    } else {
        r = load_rec(v.rec + i);
    }
    SEIR_SFINE(0);
    double p_h = cx.p_hh_vec ? __ldg(cx.p_hh_vec + i) : cx.p_hh_scalar;
    const bool home_i = at_home(v, i);
    const sd_block blk = finalize ? sd_block{{0u, 0u, 0u, 0u}} : draw_block(v.seed, i, (uint32_t)t, site, sub);
    const double u0 = sd_lane(blk, 0), u1 = sd_lane(blk, 1);
    SEIR_SFINE(1);
    bool dirty = false;
    if (is_new) {  // infected on day t-1 by the winner of that day's hits (:347-356); try `lane` of the polar method on lane `lane`
        const uint32_t inf = (uint32_t)(w_self >> 16), ord = (uint32_t)(w_self & 0xFFFFu);
        const int tx = t - 1;
        const int age = (int)(stt_new & 127u);
        // the winner of the day's hits left the timers in the record's first words (tagged with its key and the run's
        // nonce): nothing to draw.  Otherwise (a hit by a staged pass, or a later store of an earlier hitter won the race
        // for the words) they are drawn here, the tries of the polar method side by side on the lanes.
        uint32_t tiso16, tact16;
        if (hint.x == (uint32_t)w_self && hint.y == (uint32_t)((w_self & ~W_TRACED) >> 32) && hint.w == cx.run_nonce) {
            tiso16 = hint.z & 0xFFFFu;
            tact16 = hint.z >> 16;
        } else {
            const sd_block b = draw_block(v.seed, inf, (uint32_t)tx, SD_SITE_INF, (ord << 8) + lane);
            const double x1 = SD_SUB(SD_MUL(2.0, sd_lane(b, 0)), 1.0), x2 = SD_SUB(SD_MUL(2.0, sd_lane(b, 1)), 1.0);
            const double r2 = SD_ADD(SD_MUL(x1, x1), SD_MUL(x2, x2));
            const bool ok = lane >= 1u && r2 < 1.0 && r2 != 0.0;
            const double lg = log_ni(lane == 0u ? SD_SUB(1.0, sd_lane(b, 0)) : (ok ? r2 : 1.0));   // one log for all: lane 0 the exponential, the others their try
            double tiso = SD_RINT(SD_MUL(-lg, sm.iso_asym[age]));
            double z = ok ? SD_MUL(SD_SQRT(SD_DIV(SD_MUL(-2.0, lg), r2)), x2) : 0.0;
            const unsigned acc = __ballot_sync(0xFFFFFFFFu, ok);
            tiso = __shfl_sync(0xFFFFFFFFu, tiso, 0);
            double tact;
            if (acc) {
                z = __shfl_sync(0xFFFFFFFFu, z, __ffs(acc) - 1);
                const double x = exp_ni(SD_FMA(cx.act_sigma, z, cx.act_mu));
                tact = x <= 0.0 ? 1.0 : SD_RINT(x);
            } else {  // (31 rejections in a row: the plain loop goes on from there)
                tact = lognormal_days_ni(v.seed, inf, (uint32_t)tx, SD_SITE_INF, ord << 8, cx.act_mu, cx.act_sigma);
            }
            tiso16 = sd_enc16(tiso);
            tact16 = sd_enc16(tact);
        }
        r.clear();
        r.set_t_exposed((uint16_t)tx);
        r.set_tt_isolate(tiso16);
        r.set_tt_activation(tact16);
        r.set_t_end(NEVER);
        r.set_t_q_on(NEVER);
        r.set_stat(stt_new);
        r.set_state(C_E);
        if (ib & inbox_q_bit((int)(i & 15))) {
            r.set_state(r.state() | ST_Q);
            r.set_t_q_on((uint16_t)tx);
            if (lane == 0u) tally_add(cx, v, t, TAL_NEW_QE, age, tw);
        }
        if (w_self & W_TRACED) {
            r.set_state(r.state() | ST_DOC);
            r.set_t_documented((uint16_t)tx);
        }
        if (lane == 0u) tally_add(cx, v, t, TAL_NEW_E, age, tw);
        dirty = true;
    }
    SEIR_SFINE(2);
    if (finalize) {  // pass T: the agents infected on the last day get their records, nothing moves
        if (dirty && lane == 0u) store_rec(v.rec + i, r);
        return DAY_DROP;
    }
    // ---- the agent's own day (:256-337), uniform over the lanes; the draws come from the lanes that hold them
    const uint16_t stt = r.stat();
    const int age = stt & 127;
    const uint8_t st = r.state();
    const int comp = st & 7;
    const bool Ep = comp == C_E, Mp = comp == C_MILD, Sevp = comp == C_SEV, Cp = comp == C_CRIT;
    const bool Qp = (st & ST_Q) != 0, Dp = (st & ST_DOC) != 0;
    int ncomp = comp;
    bool nq = Qp, ndoc = Dp;
    bool transmit = !Qp;
    const double doc_u = __shfl_sync(0xFFFFFFFFu, u0, 0), knuth_u = __shfl_sync(0xFFFFFFFFu, u1, 0);
    const double act0 = __shfl_sync(0xFFFFFFFFu, u0, 5), act1 = __shfl_sync(0xFFFFFFFFu, u1, 5), act2 = __shfl_sync(0xFFFFFFFFu, u0, 6);
    const double sev0 = __shfl_sync(0xFFFFFFFFu, u0, 7), sev1 = __shfl_sync(0xFFFFFFFFu, u1, 7);
    const double cr0 = __shfl_sync(0xFFFFFFFFu, u0, 8), cr1 = __shfl_sync(0xFFFFFFFFu, u1, 8);
    const bool want_doc = Mp && !Dp && cx.p_doc > 0.0;
    const int t_inf = t_infected_of(r);
    if (!Ep && due(t, t_inf, r.tt_recovery())) {  // :276-279
        ncomp = C_R;
        nq = false;
        transmit = false;
        if (lane == 0u) tally_add(cx, v, t, TAL_NEW_R, age, tw);
    } else {
        const bool ev = (Ep && due(t, r.t_exposed(), r.tt_activation())) || (Mp && due(t, t_inf, r.tt_severe())) ||
                        (Sevp && due(t, t_severe_of(r), r.tt_critical()));
        if (want_doc && doc_u < cx.p_doc) {  // :280-289 (tracing days never run here: no Q from :287)
            ndoc = true;
            r.set_t_documented((uint16_t)t);
        }
        if (ev) {
            const int k3 = age * 4 + ((stt >> 7) & 1) * 2 + ((stt >> 8) & 1);
            dirty = true;
            if (Ep) {  // :256-272 activation: both exponentials at once (lane 0 the course, lane 1 the isolation time)
                ncomp = C_MILD;
                r.set_flags(r.flags() | F_MILD);
                const bool sev = act0 < __ldg(v.p_ms + k3);
                const double val = exp_days_ni(lane == 1u ? act2 : act1, lane == 1u ? sm.iso_mean[age] : (sev ? cx.m_sev : cx.m_mild_rec));
                const double tcourse = __shfl_sync(0xFFFFFFFFu, val, 0), tiso = __shfl_sync(0xFFFFFFFFu, val, 1);
                if (sev) { r.set_tt_severe(sd_enc16(tcourse)); r.set_tt_recovery(NEVER); }
                else { r.set_tt_recovery(sd_enc16(tcourse)); r.set_tt_severe(NEVER); }
                r.set_tt_isolate(sd_enc16(tiso));
                if (tiso == 0.0) nq = true;
            } else if (Mp) {  // :291-311 Mild -> Severe
                ncomp = C_SEV;
                r.set_flags(r.flags() | F_SEV);
                if (!ndoc) { ndoc = true; r.set_t_documented((uint16_t)t); }
                nq = true;
                const bool crit = sev0 < __ldg(v.p_sc + k3);
                const double val = exp_days_ni(sev1, crit ? cx.m_crit : cx.m_sev_rec);
                if (crit) { r.set_tt_critical(sd_enc16(val)); r.set_tt_recovery(NEVER); }
                else { r.set_tt_recovery(add_sat16(sd_enc16(val), r.tt_severe())); r.set_tt_critical(NEVER); }
            } else {  // :312-321 Severe -> Critical
                ncomp = C_CRIT;
                r.set_flags(r.flags() | F_CRIT);
                const bool dies = cr0 < __ldg(v.p_cd + k3);
                const double val = exp_days_ni(cr1, dies ? cx.m_death : cx.m_crit_rec);
                if (dies) { r.set_tt_death(sd_enc16(val)); r.set_tt_recovery(NEVER); }
                else { r.set_tt_recovery(add_sat16(add_sat16(sd_enc16(val), r.tt_severe()), r.tt_critical())); r.set_tt_death(NEVER); }
            }
        } else if (Cp && due(t, t_critical_of(r), r.tt_death())) {  // :323-327
            ncomp = C_D;
            nq = false;
            if (lane == 0u) tally_add(cx, v, t, TAL_NEW_D, age, tw);
        }
        if (ndoc && !Dp && lane == 0u) tally_add(cx, v, t, TAL_NEW_DOC, age, tw);
    }
    SEIR_SFINE(3);
    if (transmit) {  // :328-337 isolation entry
        if (!Ep && due(t, t_infected_of(r), r.tt_isolate())) { nq = true; transmit = false; }
        else if (Ep && due(t, r.t_exposed(), r.tt_isolate())) { nq = true; transmit = false; }
    }
    const bool active = ncomp <= C_CRIT;
    if (lane == 0u) {
        if (ncomp != comp) {
            tally_add(cx, v, t, TAL_D_E + comp - C_E, age, 0u - tw);
            if (active) tally_add(cx, v, t, TAL_D_E + ncomp - C_E, age, tw);
        }
        if ((nq && active) != Qp) tally_add(cx, v, t, TAL_D_Q, age, Qp ? 0u - tw : tw);
    }
    const uint8_t nst = (uint8_t)(ncomp | (nq ? ST_Q : 0) | (ndoc ? ST_DOC : 0));
    if (nst != st) {
        r.set_state(nst);
        if (nq && !Qp && r.t_q_on() == NEVER) r.set_t_q_on((uint16_t)t);
        if (ncomp == C_R || ncomp == C_D) r.set_t_end((uint16_t)t);
        dirty = true;
    }
    if (transmit) {
        // ---- household push :339-359 on lanes 0..6 (member j), community push :361-390 on lanes 16..31 (contact k)
        const int pos = (stt >> 9) & 7, size = ((stt >> 12) & 7) + 1;
        const uint32_t base = i - (uint32_t)pos;
        if (Ep) p_h = SD_MUL(p_h, cx.asym);
        const bool hh_lane = (int)lane < size - 1;
        const uint32_t member = base + lane + ((int)lane >= pos ? 1u : 0u);
        unsigned long long w_m = 1ull;
        uint32_t stat_m = 0;
        if (hh_lane) { w_m = __ldcg(v.winner + member); stat_m = __ldg(cx.stat + member); }
        // member j draws lane j&1 of household block j>>1, which lane 1 + (j>>1) holds
        const int src = 1 + (int)((lane >> 1) & 3u);
        const double h0 = __shfl_sync(0xFFFFFFFFu, u0, src), h1 = __shfl_sync(0xFFFFFFFFu, u1, src);
        // the kept-contact count: Knuth's product over lane 1 of the day block and the NCNT blocks of lanes 9..12
        // (every branch below is warp-uniform: the shuffles are executed by all lanes)
        uint32_t kept = 0;
        if (!home_i) {
            const double en = sm.nat_enlam[(v.phase * kAges + age) * 2 + (Ep ? 1 : 0)];
            double prod = knuth_u;
            if (prod > en) {
                kept = 1;
                bool done = false;
#pragma unroll 1
                for (int c = 0; c < 4 && !done; ++c) {
                    prod = SD_MUL(prod, __shfl_sync(0xFFFFFFFFu, u0, 9 + c));
                    if (prod <= en || kept >= SD_MAX_CONTACTS_NATIVE) done = true;
                    else {
                        kept += 1;
                        prod = SD_MUL(prod, __shfl_sync(0xFFFFFFFFu, u1, 9 + c));
                        if (prod <= en || kept >= SD_MAX_CONTACTS_NATIVE) done = true;
                        else kept += 1;
                    }
                }
                if (!done) kept = poisson_day_from(v.seed, i, (uint32_t)t, prod, en, SD_MAX_CONTACTS_NATIVE, kept, 4u);
            }
        }
    SEIR_SFINE(4);
        uint32_t hits = 0, hits_out = 0;
        const bool first_batch_hh = hh_lane && (w_m == 0ull || wday(w_m) == t) && ((lane & 1u) ? h1 : h0) < p_h;  // :346
        for (uint32_t k0 = 0; k0 == 0u || k0 < kept; k0 += 16u) {
            // This lane's candidate of the batch.  The gathers of a contact are three dependent round trips (alias row ->
            // member -> the member's winner word); the infectee's timers (:352-356, :381-385) depend on (i, t, ordinal)
            // only, so their draws are evaluated BETWEEN the gathers: the warp computes while the loads fly.
            bool cand = false, is_hit = false;
            uint32_t c = 0, ord = 0;
            int age_c = 0, col = 0, al = 0;
            double x = 0.0, pr = 0.0;
            bool s_c = true;  // S[t-1, c] (a household candidate passed the test above)
            sd_block b = blk;
            const uint32_t k = k0 + lane - 16u;
            const bool contact = lane >= 16u && k < kept;
            if (contact) {  // :373-375, native: contact age from the alias table of the sender's row, a uniform member
                if (k0 != 0u) b = draw_block(v.seed, i, (uint32_t)t, SD_SITE_NPICK, k);
                x = SD_MUL(sd_lane(b, 0), (double)kAges);
                col = (int)x;
                if (col > kAges - 1) col = kAges - 1;
                pr = __ldg(cx.nat_alias_prob + age * kAges + col);
                al = __ldg(cx.nat_alias_idx + age * kAges + col);
                cand = true;
                ord = SD_ORD_COMM_NATIVE(k);
            } else if (k0 == 0u && first_batch_hh) {
                cand = true;
                c = member;
                ord = SD_ORD_HH(lane);
                age_c = (int)(stat_m & 127u);
            }
            if (!__any_sync(0xFFFFFFFFu, cand)) continue;
    SEIR_SFINE(5);
            // the isolation draw's logarithm (in the shadow of the alias loads)
            double lg = 0.0;
            if (cand) lg = log_ni(SD_SUB(1.0, sd_lane(draw_block(v.seed, i, (uint32_t)t, SD_SITE_INF, ord << 8), 0)));
            if (contact) {
                age_c = (SD_SUB(x, (double)col) < pr) ? col : al;
                const int g0 = sm.grp_off[age_c], glen = sm.grp_off[age_c + 1] - g0;
                if (glen != 0) c = (uint32_t)__ldg(cx.grp_members + g0 + (int64_t)sd_index(b, (uint64_t)glen));
                else cand = false;
            }
            // the activation time the infectee would get from this hit (in the shadow of the member load)
            double tact = 0.0;
            if (cand) tact = lognormal_days_ni(v.seed, i, (uint32_t)t, SD_SITE_INF, ord << 8, cx.act_mu, cx.act_sigma);
            bool home_c = false;
            if (contact && cand) {
                if (SEIR_CHECK((int64_t)c < cx.n, cx, v.rep)) {
                    if (cx.n_shards > 1) {
                        const uint32_t xd = __ldcg(v.xday + c);
                        s_c = xd == 0u || xd == (uint32_t)t + 1u;
                    } else {
                        const unsigned long long w_c = __ldcg(v.winner + c);
                        s_c = w_c == 0ull || wday(w_c) == t;
                    }
                    home_c = at_home(v, c);  // (:375; the household push has no Home test, :346)
                } else {
                    cand = false;
                }
            }
            const double tiso = SD_RINT(SD_MUL(-lg, sm.iso_asym[age_c]));
    SEIR_SFINE(6);
            is_hit = cand && s_c && !home_c;
            if (is_hit) {
                const unsigned long long key = ((unsigned long long)t << 48) | ((unsigned long long)i << 16) | ord;
                // the hit and a slot of tomorrow's list of new infectees are requested together: the slot of a hit that
                // turns out not to be the first one on c today is left as a hole
                const unsigned long long old = atomicMax(v.winner + c, key);
                const uint32_t p = atomicAdd_system(v.next_new_cnt, 1u);
                const bool first = wday(old) != t;
                if (first && cx.n_shards > 1) __stcg(v.xday + c, (uint16_t)(t + 1));  // (a replicated day of a sharded run)
                if (SEIR_CHECK((int64_t)p < cx.n_pad, cx, v.rep)) __stcg(v.next_new + p, first ? c : kHole);
                uint32_t bits = first ? inbox_infected_bit((int)(c & 15)) : 0u;
                if (tiso == 0.0) bits |= inbox_q_bit((int)(c & 15));
                if (bits) atomicOr(v.inbox + (c >> 4), bits);
                // the winner so far leaves the infectee's timers where its record will be built tomorrow (rec[c] is not
                // in use yet): the builder takes them if the tag is the day's final winner, and draws them itself otherwise
                if ((old & ~W_TRACED) < key)
                    __stcg(reinterpret_cast<uint4 *>(v.rec + c),
                           make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)sd_enc16(tiso) | ((uint32_t)sd_enc16(tact) << 16), cx.run_nonce));
                if (TR && cx.log_ent && ord >= 16u) {  // infected_by[i, k] = c (:387)
                    const uint32_t e = atomicAdd(cx.log_cnt + v.rep, 1u);
                    if (e < cx.log_cap) {
                        const uint32_t prev = atomicExch(cx.log_head + (size_t)v.rep * cx.n_pad + i, e + 1u);
                        __stcg(cx.log_ent + (size_t)v.rep * cx.log_cap + e, make_uint4(c, prev, ((uint32_t)t << 16) | ord, 0u));
                    } else {
                        atomicOr(&cx.counters[v.rep * 4 + 0], 1ull);
                    }
                }
            }
    SEIR_SFINE(7);
            const unsigned hm = __ballot_sync(0xFFFFFFFFu, is_hit);
            hits += (uint32_t)__popc(hm);
            hits_out += (uint32_t)__popc(hm & 0xFFFF0000u);
        }
        if (hits) {  // :357-359, :386-390
            r.set_n_inf_by(inc_sat16(r.n_inf_by(), hits));
            if (Ep) r.set_n_inf_asympt(inc_sat16(r.n_inf_asympt(), hits));
            r.set_n_inf_outside(inc_sat16(r.n_inf_outside(), hits_out));
            dirty = true;
            if (own && lane == 0u) atomicAdd(&sm.cnt[2], hits);
        }
    }
    SEIR_SFINE(8);
    if (dirty && lane == 0u) store_rec(v.rec + i, r);
    if (!active) return DAY_DROP;
    if (!nq) return DAY_KEEP;
    int wake = t + 1;
    if (Qp || !(ncomp == C_MILD && !ndoc && cx.p_doc > 0.0)) wake = next_wake_day_warp(cx, v.seed, i, r, ncomp, ndoc, t);
    if (wake >= v.T) return DAY_DROP;
    if (wake > t + kParkHorizon) wake = t + kParkHorizon;
    return DAY_KEEP_Q | ((wake - t - 1) << 2);
}

// the lane's block of the agent's day: (site, sub) of lane l
static __device__ __forceinline__ void sparse_lane_role(uint32_t lane, uint32_t &site, uint32_t &sub)
{
    site = SD_SITE_DAY; sub = 0u;                                                 // lane 0 (and the spare lanes 13..15)
    if (lane >= 16u) { site = SD_SITE_NPICK; sub = lane - 16u; }                  // contact picks 0..15
    else if (lane >= 1u && lane <= 4u) { site = SD_SITE_HH; sub = lane - 1u; }    // household blocks 0..3
    else if (lane == 5u) { site = SD_SITE_ACT; sub = 0u; }
    else if (lane == 6u) { site = SD_SITE_ACT; sub = 1u; }
    else if (lane == 7u) { site = SD_SITE_SEV; sub = 0u; }
    else if (lane == 8u) { site = SD_SITE_CRIT; sub = 0u; }
    else if (lane >= 9u && lane <= 12u) { site = SD_SITE_NCNT; sub = lane - 9u; } // Knuth uniforms 2..9
}

// The replica's lists of day t by the warps of its CTAs: entry e goes to warp e / bpr of CTA part e mod bpr.
// (entries: [0, n_pk) parked agents due today, [n_pk, n_pk + n_new) infected yesterday, then the transmitting list)
static __device__ __noinline__ void sparse_rounds(const DevCtx &cx, PassSmem &sm, int t, uint32_t n_new, uint32_t n_old, uint32_t n_pk,
                                                     const uint32_t *cur_new, const uint32_t *cur_old, const uint32_t *cur_pk,
                                                     int my_part, int bpr, bool finalize, bool sharded_now)
{
    const RepView &v = sm.v;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t site, sub;
    sparse_lane_role(lane, site, sub);
    const uint32_t count = n_pk + n_new + n_old;
    for (uint32_t e = warp * (uint32_t)bpr + (uint32_t)my_part; e < count; e += (blockDim.x >> 5) * (uint32_t)bpr) {
        const int kind = e < n_pk ? 0 : e < n_pk + n_new ? 1 : 2;
        const uint32_t idx = kind == 0 ? e : kind == 1 ? e - n_pk : e - n_pk - n_new;
        const uint32_t i = __ldcg((kind == 0 ? cur_pk : kind == 1 ? cur_new : cur_old) + idx);
        int keep = DAY_DROP;
        SEIR_SFINE(11);
        if (i != kHole && SEIR_CHECK((int64_t)i < cx.n, cx, v.rep)) {
            const bool own = cx.n_shards <= 1 || (i >= cx.shard_lo[cx.my_shard] && i < cx.shard_lo[cx.my_shard + 1]);
            keep = sparse_entry(cx, sm, i, kind, t, finalize, own, site, sub);
            if (lane == 0u && own) {
                atomicAdd(&sm.cnt[0], 1u);
                if (kind == 1) atomicAdd(&sm.cnt[1], 1u);
            }
        }
        // (an agent a staged pass routed through today's bucket on the eve of its progression step -- agent_day -- and that
        // may transmit tomorrow: the bucket list has no slot in tomorrow's transmitting list, it goes on through the buckets)
        if (kind == 0 && keep == DAY_KEEP) keep = t + 1 < v.T ? DAY_KEEP_Q : DAY_DROP;
        if (lane == 0u && !finalize) {
            // tomorrow's transmitting list keeps today's slots: the list's entries first, then the new infectees
            if (kind != 0) v.next_old[kind == 2 ? idx : n_old + idx] = keep == DAY_KEEP ? i : kHole;
            if ((keep & 3) == DAY_KEEP_Q) {
                const int wake = t + 1 + (keep >> 2);
                const uint32_t p = atomicAdd(v.pk_cnt + wake, 1u);
                if (SEIR_CHECK((int64_t)p < cx.n_pad, cx, v.rep)) __stcg(v.pk_pool + (size_t)(wake & (kParkRing - 1)) * cx.n_pad + p, i);
            }
        }
    }
    SEIR_SFINE(13);
    if (my_part == 0 && threadIdx.x == 0 && !finalize) *v.next_old_cnt = n_old + n_new;  // (nobody else touches this counter today)
    (void)sharded_now;
}

// the replica's view of pass t (thread 0 of the CTA)
static __device__ __noinline__ void fill_view(const DevCtx &cx, PassSmem &sm, int t, int rep, bool sharded_now)
{
    const size_t lrow = ((size_t)rep * 2 + ((t + 1) & 1)) * cx.n_pad;
    uint32_t *cnt_old = cx.list_cnt + (size_t)rep * 3 * (cx.T + 2), *cnt_new = cnt_old + (cx.T + 2), *cnt_pk = cnt_new + (cx.T + 2);
    if (sm.rp_rep != rep) {  // (one request per CTA and replica: the overrides and the seed do not change during a launch)
        sm.rp = cx.rp[rep];
        sm.v.seed = cx.seeds[rep];
        sm.rp_rep = rep;
    }
    const RepParams rp = sm.rp;
    sm.v.T = rp.T;
    const double *bank = cx.bank + (size_t)rp.bank * kBankStride;
    sm.v.p_ms = bank + kBankPms; sm.v.p_sc = bank + kBankPsc; sm.v.p_cd = bank + kBankPcd;
    sm.v.p_contact = bank + kBankPcontact;
    sm.v.inbox = cx.inbox + (size_t)rep * (cx.n_pad / 16);
    sm.v.winner = cx.winner + (size_t)rep * cx.n_pad;
    sm.v.xday = cx.xday + (size_t)rep * cx.n_pad;
    sm.v.rec = cx.rec + (size_t)rep * cx.n_pad;
    sm.v.home = cx.home + (size_t)rep * (cx.n_pad / 32);
    sm.v.next_old = cx.list_old + lrow;
    sm.v.next_new = cx.list_new + lrow;
    sm.v.next_old_cnt = cnt_old + (t + 1);
    sm.v.next_new_cnt = cnt_new + (t + 1);
    sm.v.pk_cnt = cnt_pk;
    sm.v.pk_pool = cx.pk_pool + (size_t)rep * kParkRing * cx.n_pad;
    sm.v.n_pad_m1 = (uint32_t)(cx.n_pad - 1);
    sm.v.home_on = rp.t_stayhome >= 1 && t >= rp.t_stayhome;          // :243-244
    sm.v.phase = (rp.t_lockdown >= 1 && t >= rp.t_lockdown) ? 1 : 0;  // :234-240
    sm.v.rep = rep;
    sm.v.tracing = rp.trace_from >= 1 && t >= rp.trace_from;          // :241-242
    sm.v.sharded = sharded_now;
    sm.v.replicated = cx.n_shards > 1 && !sharded_now;
}
// today's list lengths of the replica (new, transmitting, parked due), by ONE thread of the CTA
static __device__ __forceinline__ void load_list_lengths(const DevCtx &cx, PassSmem &sm, int t, int rep)
{
    const uint32_t *c = cx.list_cnt + (size_t)rep * 3 * (cx.T + 2) + t;
    const uint32_t a = __ldcg(c + (cx.T + 2)), b = __ldcg(c), d = __ldcg(c + 2 * (cx.T + 2));
    sm.list_n[0] = a; sm.list_n[1] = b; sm.list_n[2] = d;
}
// this CTA's run counters of the pass (thread 0; the CTA has met since the last of them was counted)
static __device__ __noinline__ void flush_pass_counters(const DevCtx &cx, PassSmem &sm, int rep, bool defer_counters)
{
    if (defer_counters) {
        sm.run_cnt[0] += sm.cnt[0]; sm.run_cnt[1] += sm.cnt[1]; sm.run_cnt[2] += sm.cnt[2];
    } else {
        if (sm.cnt[0]) atomicAdd(&cx.counters[rep * 4 + 1], (unsigned long long)sm.cnt[0]);
        if (sm.cnt[1]) atomicAdd(&cx.counters[rep * 4 + 2], (unsigned long long)sm.cnt[1]);
        if (sm.cnt[2]) atomicAdd(&cx.counters[rep * 4 + 3], (unsigned long long)sm.cnt[2]);
    }
#ifdef SEIR_CHECKS
    if (sm.cnt[3]) atomicOr(&cx.counters[rep * 4 + 0], 4ull);
#endif
    sm.cnt[0] = 0; sm.cnt[1] = 0; sm.cnt[2] = 0; sm.cnt[3] = 0;
}

// ---------------------------------------------------------------------------------------
// Pass t over all replicas.  CTA b serves replica (b / blocks_per_rep) [+ k * reps_per_sweep]; the CTAs of a
// replica split its list: a list shorter than the replica's threads is dealt S threads apart (one entry
// per warp while it fits), a longer one in equal contiguous chunks, kPerThread entries per thread and round.
#ifdef SEIR_CTA_STAMPS  // (debug) the eight stage stamps of EVERY CTA (the first 2 (T + 2) of them) for pass cx.dbg_pass, into fine_ns
#define SEIR_STAMP(k) do { if (tid == 0 && !stamped) { \
        if (blockIdx.x == 0) cx.stage_ns[(size_t)t * 8 + (k)] = global_timer_ns(); \
        if (t == cx.dbg_pass && blockIdx.x < (unsigned)((cx.T + 2) * 2)) cx.fine_ns[blockIdx.x * 8 + (k)] = global_timer_ns(); } } while (0)
#else
#define SEIR_STAMP(k) do { if (blockIdx.x == 0 && tid == 0 && !stamped) cx.stage_ns[(size_t)t * 8 + (k)] = global_timer_ns(); } while (0)
#endif
// (inline_sparse: a sparse day runs as a warp-cooperative pass inside this kernel -- replica ensembles, tracing builds and
// sharded runs; the single-replica day loop hands its sparse days to seir_sparse_kernel instead and keeps this code out)
// (can_yield: the single-replica dense-day kernel; returns true WITHOUT touching the day when its lists have shrunk to
// fewer than yield_below entries (0: never) or, in a sharded run's replicated phase, more than yield_above (the generic
// kernel takes over: sharded passes) -- the sparse-day kernel takes over)
template <bool inline_sparse, bool can_yield = false>
static __device__ bool day_pass(const DevCtx &cx, PassSmem &sm, int t, int my_slot, int my_part, int reps_per_sweep,
                                int &bank_staged, bool defer_counters = false, bool sharded_now = false, uint32_t yield_below = 0u,
                                uint32_t yield_above = 0xFFFFFFFFu)
{
    const int tid = threadIdx.x;
    bool stamped = false;
    int bpr = cx.blocks_per_rep;
    int rep_first = my_slot < reps_per_sweep ? my_slot : cx.n_rep, rep_step = reps_per_sweep;
#if SEIR_DYNAMIC_MAP
    // Replica ensembles (one CTA never serves two replicas in a pass): the epidemics of an ensemble grow at different
    // times, and with a fixed share of the grid per replica a pass lasts as long as its largest replica.  So warp 0 of
    // every CTA reads the list lengths of ALL replicas and deals the grid anew: one CTA per replica, the others in
    // proportion to the lists; the result does not depend on who processes what.
    if (cx.n_rep > 1 && cx.n_rep <= (int)gridDim.x && cx.n_shards <= 1) {
        const int R = cx.n_rep;
        if (tid < 32) {
            uint32_t *scratch = sm.hit_n;  // (free until the first round starts)
            unsigned long long W = 0;
            for (int r = tid; r < R; r += 32) {
                const uint32_t *c0 = cx.list_cnt + (size_t)r * 3 * (cx.T + 2) + t;
                const uint32_t c = __ldcg(c0) + __ldcg(c0 + (cx.T + 2)) + __ldcg(c0 + 2 * (cx.T + 2));
                scratch[r] = c;
                W += c;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) W += __shfl_xor_sync(0xFFFFFFFFu, W, o);
            if (tid == 0) sm.dyn[0] = R;  // (idle unless an interval below holds this CTA)
            __syncwarp();
            const uint32_t spare = gridDim.x - (uint32_t)R, b = blockIdx.x;
            uint32_t carry = 0;
            for (int r0 = 0; r0 < R; r0 += 32) {
                const int r = r0 + tid;
                const uint32_t k = r < R ? 1u + (W ? (uint32_t)((unsigned long long)spare * scratch[r] / W) : 0u) : 0u;
                uint32_t incl = k;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (tid >= o) incl += up;
                }
                const uint32_t excl = carry + incl - k;
                if (r < R && b >= excl && b < excl + k) { sm.dyn[0] = r; sm.dyn[1] = (int)(b - excl); sm.dyn[2] = (int)k; }
                carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
        __syncthreads();
        rep_first = sm.dyn[0];
        rep_step = R;
        if (rep_first < R) { my_part = sm.dyn[1]; bpr = sm.dyn[2]; }
    }
#endif
    // (CTAs without a replica idle through the pass; they must still reach the grid barrier)
    for (int rep = rep_first; rep < cx.n_rep; rep += rep_step) {
        const size_t lrow = ((size_t)rep * 2 + ((t + 1) & 1)) * cx.n_pad, lcur = ((size_t)rep * 2 + (t & 1)) * cx.n_pad;
        uint32_t *cnt_old = cx.list_cnt + (size_t)rep * 3 * (cx.T + 2), *cnt_new = cnt_old + (cx.T + 2), *cnt_pk = cnt_new + (cx.T + 2);
        const RepView &v = sm.v;
        if (tid == 0) fill_view(cx, sm, t, rep, sharded_now);

        const uint32_t *cur_old = cx.list_old + lcur, *cur_new = cx.list_new + lcur;
        const uint32_t *cur_pk = cx.pk_pool + ((size_t)rep * kParkRing + (size_t)(t & (kParkRing - 1))) * cx.n_pad;  // today's bucket
        // today's list lengths: ONE request per CTA and counter (a load by every warp of the grid is 4 736 requests
        // for the same L2 line, which its slice serves one at a time: 2 us at the start of every pass)
        // (by another warp than thread 0's: its requests fly while the view is being filled)
        if (tid >= 64 && tid < 67) sm.list_n[tid - 64] = __ldcg((tid == 64 ? cnt_new : tid == 65 ? cnt_old : cnt_pk) + t);
        if (tid >= 32 && tid < 36) sm.cnt[tid - 32] = 0;
        __syncthreads();
        const RepParams rp = sm.rp;
        if (t > rp.T) {  // (a replica with a shorter horizon than the launch's: finished)
            __syncthreads();  // (sm.rp is rewritten for the next replica)
            continue;
        }
        const bool finalize = t >= rp.T;
        if (bank_staged != rp.bank) {  // (CTA-uniform: every thread keeps track of the set it saw staged; read after the round's first barrier)
            const double *src = cx.bank + (size_t)rp.bank * kBankStride + kBankNatEn;
            for (int k = tid; k < 4 * kAges; k += kThreads) sm.nat_enlam[k] = src[k];
            bank_staged = rp.bank;
            __syncthreads();  // (once per CTA and table set)
        }
        // yesterday's bucket has been consumed: its chunks go back to the free ring (one warp of the replica's first CTA;
        // today's allocations cannot reach these ring slots: the ring holds twice the chunks that can be in use)
        // (pass T only finalises: the agents infected on the last day get their records, nobody else is visited)
        if (can_yield) {
            const uint32_t total = sm.list_n[0] + sm.list_n[1] + sm.list_n[2];
            if (total < yield_below || total > yield_above) return true;
        }  // (grid-uniform)
        const uint32_t n_new = sm.list_n[0], n_old = finalize ? 0u : sm.list_n[1], n_pk = finalize ? 0u : sm.list_n[2];
        // a SPARSE day: one warp per entry (sparse_rounds); the staged rounds below serve the days whose lists outgrow the warps
        const bool sparse = inline_sparse && cx.sparse_rounds != 0u && cx.mode == SEIR_MODE_NATIVE && !sharded_now && !(TR && v.tracing) &&
                            n_pk + n_new + n_old <= cx.sparse_rounds * (uint32_t)(kWarps * bpr) && n_pk + n_new + n_old <= cx.sparse_max;
        if (sparse) {
            SEIR_STAMP(0);
            SEIR_SFINE(10);
            sparse_rounds(cx, sm, t, n_new, n_old, n_pk, cur_new, cur_old, cur_pk, my_part, bpr, finalize, sharded_now);
            SEIR_STAMP(1); SEIR_STAMP(2); SEIR_STAMP(3); SEIR_STAMP(4); SEIR_STAMP(5); SEIR_STAMP(6);
            __syncthreads();
        } else {
        for (int k = tid; k < TAL_ROWS * kAges; k += kThreads) sm.tally[k] = 0;
        if (tid == 64) { sm.n_old = 0; sm.n_pk = 0; sm.n_new = 0; sm.n_hit = 0; sm.n_con = 0; sm.n_draw = 0; sm.next_group = 0; }
        if (tid >= 96 && tid < 128) sm.pk_n[tid - 96] = 0;
        __syncthreads();
        // three segments, each starting on a multiple of 32 -- a warp of a dense day serves one kind -- the heaviest
        // first (the warps grab the groups in this order, and the pass lasts as long as its last group): parked
        // agents due today (every one of them has an event, most of them draws), new infectees (record build),
        // agents that may transmit
        const uint32_t o_n = (n_pk + 31u) & ~31u, o_t = o_n + ((n_new + 31u) & ~31u);
        const uint32_t count = o_t + n_old;
        // this CTA's share of the list: groups of G = 32/S entries (one entry per group while the list is
        // shorter than the replica's warps: a warp pays for every divergent path of its lanes), dealt round-robin
        // to the replica's CTAs (the new infectees, whose record build is the longest path, lead the list: a
        // contiguous split would give them all to the first CTAs) and grabbed dynamically by the CTA's warps.
        uint32_t lgG = 0;  // G = 32 / spread_of(count, threads of the replica) as a shift
        if (count >= (uint32_t)bpr * kThreads) lgG = 5u;
        else while (lgG < 5u && (count << (5u - lgG)) > (uint32_t)bpr * kThreads) ++lgG;
        const uint32_t G = 1u << lgG;
        const uint32_t n_groups = (count + G - 1u) >> lgG;
        const uint32_t my_groups = n_groups > (uint32_t)my_part ? (n_groups - (uint32_t)my_part + (uint32_t)bpr - 1u) / (uint32_t)bpr : 0u;
        constexpr uint32_t kGroupsPerRound = (uint32_t)(kSlots / 32);
        const uint32_t rounds = (my_groups + kGroupsPerRound - 1u) / kGroupsPerRound;
        SEIR_STAMP(0);

        uint32_t reserved = 0;  // thread 0: offset of the round's staged survivors in tomorrow's list
        uint32_t park_reserved = 0;  // threads 128 .. 142: offset of the round's parks in their day bucket
        for (uint32_t round = 0; round < rounds; ++round) {
            // ================= stage A: the round's groups, grabbed by the warps =================
            const uint32_t g_lo = round * kGroupsPerRound;
            const uint32_t g_hi = g_lo + kGroupsPerRound < my_groups ? g_lo + kGroupsPerRound : my_groups;
            for (uint32_t sl = tid; sl < (g_hi - g_lo) * 32u; sl += kThreads) sm.hit_n[sl] = 0;
            __syncthreads();  // (a full draw queue makes stage A process hits in place)
            for (;;) {
                uint32_t k = 0;
                if ((tid & 31) == 0) k = g_lo + atomicAdd(&sm.next_group, 1u);
                k = __shfl_sync(0xFFFFFFFFu, k, 0);
                if (k >= g_hi) break;
                const uint32_t lane = (uint32_t)tid & 31u;
#ifdef SEIR_FINE_STAMPS
                { const uint32_t slot = (k - g_lo) * 32u + lane; SEIR_FINE(0); }
#endif
                const uint32_t e = ((k * (uint32_t)bpr + (uint32_t)my_part) << lgG) + lane;
                const bool is_q = e < n_pk, is_new = e >= o_n && e - o_n < n_new, is_t = e >= o_t && e - o_t < n_old;
                const bool has = lane < G && (is_new || is_t || is_q);
                const uint32_t slot = (k - g_lo) * 32u + lane;
                int keep = DAY_DROP;
                uint32_t i = 0;
                bool own = true;
#if SEIR_PREFETCH
                // the group one of this CTA's warps takes kWarps grabs from now: ask the L2 for its agents' records and
                // household words now (the loads of a group are dependent DRAM round trips: entry -> record -> household)
                uint32_t i2 = 0xFFFFFFFFu;
                bool new2 = false;
                {
                    const uint32_t k2 = k + (uint32_t)kWarps;
                    const uint32_t e2 = ((k2 * (uint32_t)bpr + (uint32_t)my_part) << lgG) + lane;
                    if (k2 < g_hi && lane < G) {
                        if (e2 < n_pk) i2 = __ldcg(cur_pk + e2);
                        else if (e2 >= o_n && e2 - o_n < n_new) { i2 = __ldcg(cur_new + (e2 - o_n)); new2 = true; }
                        else if (e2 >= o_t && e2 - o_t < n_old) i2 = __ldcg(cur_old + (e2 - o_t));
                    }
                }
#endif
                if (has) {
                    i = is_q ? __ldcg(cur_pk + e) : is_new ? __ldcg(cur_new + (e - o_n)) : __ldcg(cur_old + (e - o_t));
                }
                const bool hole = has && i == kHole;  // (left behind by a warp-cooperative pass)
                if (has && !hole) {
#if SEIR_PREFETCH
                    if (!is_new && (int64_t)i < cx.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(v.rec + i));  // (its own record: in flight while i2 arrives)
#endif
#if SEIR_PREFETCH
                    if (i2 != 0xFFFFFFFFu && (int64_t)i2 < cx.n) {
                        if (!new2) asm volatile("prefetch.global.L2 [%0];" ::"l"(v.rec + i2));
                        else asm volatile("prefetch.global.L2 [%0];" ::"l"(cx.stat + i2));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(v.winner + i2));
                    }
#endif
                    SEIR_FINE(1);
                    // (a sharded run after its switch: the lists still hold what the replicated days left; each GPU keeps its own agents)
                    own = cx.n_shards <= 1 || (i >= cx.shard_lo[cx.my_shard] && i < cx.shard_lo[cx.my_shard + 1]);
                    if ((own || !sharded_now) && SEIR_CHECK((int64_t)i < cx.n && slot < (uint32_t)kSlots && (int64_t)(n_new + n_old + n_pk) <= cx.n, cx, rep))
                        keep = stage_a_entry(cx, sm, slot, i, is_new, t, finalize, own);
                    // (the first sharded day: an agent of another GPU that was infected on the last replicated day is in
                    // everybody's list; its owner steps it, the others complete its record -- their inbox bit says "infected",
                    // and the output kernels read the record of every such agent)
                    else if (is_new && !own && SEIR_CHECK((int64_t)i < cx.n && slot < (uint32_t)kSlots, cx, rep))
                        stage_a_entry(cx, sm, slot, i, true, t, true, false);
                    SEIR_FINE(8);
                }
                append_survivors(sm, keep, i);
                SEIR_FINE(9);
                // (run counters: every agent is counted by its owner -- a replicated day steps the others' agents too)
                const unsigned hm = __ballot_sync(0xFFFFFFFFu, has && !hole && own), nm = __ballot_sync(0xFFFFFFFFu, has && !hole && own && is_new);
                if (lane == 0 && hm) {
                    atomicAdd(&sm.cnt[0], (uint32_t)__popc(hm));
                    if (nm) atomicAdd(&sm.cnt[1], (uint32_t)__popc(nm));
                }
            }
            __syncthreads();
            SEIR_STAMP(1);
            SEIR_STAMP(2);
            reserved = reserve_survivors(sm);  // (nobody stages a survivor after this point of the round)
            park_reserved = reserve_parks(sm, t);
            // ================= stage D: draw items =================
            if (sm.n_draw) {
                const uint32_t nd = sm.n_draw < (uint32_t)kDrawCap ? sm.n_draw : (uint32_t)kDrawCap;
                for_items(nd, [&](uint32_t q) { process_draw_item(cx, sm, sm.drawq[q], t); });
                __syncthreads();
            }
            SEIR_STAMP(3);
            // ================= stage C: candidate contacts =================
            if (sm.n_con) {
                const uint32_t nc = sm.n_con < (uint32_t)kConCap ? sm.n_con : (uint32_t)kConCap;
                for_items(nc, [&](uint32_t q) { process_contact(cx, sm, sm.conq[q], t); });
                __syncthreads();
            }
            SEIR_STAMP(4);
            // ================= stage H: hits =================
            const bool any_hit = sm.n_hit != 0u;
            if (any_hit) {
                const uint32_t nh = sm.n_hit < (uint32_t)kHitCap ? sm.n_hit : (uint32_t)kHitCap;
                for_items(nh, [&](uint32_t q) { process_hit(cx, sm, sm.hit_c[q], sm.hit_meta[q], sm.hit_age[q], t); });
                __syncthreads();
            }
            SEIR_STAMP(5);
            // (last round: the day's new infectees of this CTA are all staged -- thread 64 reserves their room in tomorrow's
            // list now, and the round trip overlaps stage W instead of standing in front of the grid barrier)
            // (a sharded pass reserves per owner, inside the flush)
            if (round + 1u == rounds && tid == 64 && !v.sharded) {
                const uint32_t sn = sm.n_new < (uint32_t)kOutNew ? sm.n_new : (uint32_t)kOutNew;
                if (sn) reserved = atomicAdd_system(v.next_new_cnt, sn);
            }
            // ================= stage W: infector counters (:357-359, :386-390) =================
            {
                const uint32_t ns = (g_hi - g_lo) * 32u;
                if (any_hit) {
                    for (uint32_t sl = tid; sl < ns; sl += kThreads) {
                        const uint32_t hn = sm.hit_n[sl];
                        if (hn) {
                            const uint32_t hits = hn & 0xFFFFu, hits_out = hn >> 16;
                            // second half of the record: n_inf_by | n_inf_asympt<<16, n_inf_outside | t_end<<16, ...
                            uint2 *q = reinterpret_cast<uint2 *>(v.rec + sm.slot_agent[sl]) + 2;
                            uint2 u = __ldcg(q);
                            uint32_t by = inc_sat16((uint16_t)(u.x & 0xFFFFu), hits), as = u.x >> 16;
                            if (sm.slot_info[sl] & 128u) as = inc_sat16((uint16_t)as, hits);
                            u.x = by | (as << 16);
                            u.y = (u.y & 0xFFFF0000u) | inc_sat16((uint16_t)(u.y & 0xFFFFu), hits_out);
                            __stcg(q, u);
                            if (!(sm.slot_info[sl] & 0x8000u)) atomicAdd(&sm.cnt[2], hits);
                        }
                    }
                }
            }
            __syncthreads();
            // more rounds follow: empty the staging buffers
            if (round + 1u < rounds) {
                flush_survivors(cx, sm, t, sm.n_new > (uint32_t)(kOutNew - kHitCap), reserved, park_reserved);
                if (tid == 0) { sm.n_hit = 0; sm.n_con = 0; sm.n_draw = 0; sm.next_group = 0; }
                __syncthreads();
            }
            SEIR_STAMP(6);
            stamped = true;
        }
        stamped = false;
        flush_survivors(cx, sm, t, true, reserved, park_reserved, rounds != 0u);
        // ---- flush the tallies: rows [0,kTalLagRows) belong to day t-1, the rest to day t
        uint32_t *tal = cx.tally + (size_t)rep * cx.T * TAL_ROWS * kAges;
        for (int k = tid; k < TAL_ROWS * kAges; k += kThreads) {
            const uint32_t val = sm.tally[k];
            if (val) {
                const int row = k / kAges;
                const int day = row < kTalLagRows ? t - 1 : t;
                if (day < rp.T) atomicAdd(&tal[(size_t)day * TAL_ROWS * kAges + k], val);
            }
        }
        }  // (staged pass)
        // run counters (list entries visited, infections, hits): 296 CTAs adding to the same three words every pass
        // queue up at one L2 slice in front of the grid barrier's fence; a CTA that serves ONE replica for the
        // whole kernel keeps its share in shared memory and adds it once, at the end
        if (tid == 0) flush_pass_counters(cx, sm, rep, defer_counters);
        __syncthreads();
        SEIR_STAMP(7);
    }
    return false;
}

};  // struct Pass

// ---------------------------------------------------------------------------------------
// Contact tracing of day t (do_contact_tracing, seir_individual.py:60-88), after the day's transmission is
// complete.  The reference resolves the day's traces INSIDE its sequential agent loop, so what a trace
// sees depends on the agent order: (a) the tracer of root r sees the day-t entries of infected_by[c]
// only for c < r (c's iteration has run), (b) a contact that recovers or dies today keeps the tracer's
// Q[t,c] = True only if the tracer's root comes after it (r > c), (c) a household edge is followed only
// by the first tracer that reaches the member (`traced`).  To keep all of that exact, ONE warp per replica
// walks the day's roots in ascending agent order (two-level bitmap) and runs each root's search with an
// explicit stack; its 32 lanes evaluate the edges of the node being expanded (household members and log
// entries) in parallel.  Every draw is keyed on (tracer, day, site, edge slot), so re-expansions by later
// roots -- which the reference performs, its outside edges have no `traced` guard -- repeat the same draws.
#ifdef SEIR_TRACE_STATS  // (debug) per-day counters of the parallel tracer in fine_ns[t][k], replica 0 only
#define SEIR_TSTAT_ADD(k, v) do { if (tv.rep == 0) atomicAdd(cx.fine_ns + (size_t)tv.t * 16 + (k), (unsigned long long)(v)); } while (0)
#define SEIR_TSTAT_MAX(k, v) do { if (tv.rep == 0) atomicMax(cx.fine_ns + (size_t)tv.t * 16 + (k), (unsigned long long)(v)); } while (0)
#else
#define SEIR_TSTAT_ADD(k, v) do { } while (0)
#define SEIR_TSTAT_MAX(k, v) do { } while (0)
#endif
// the day's edge cache (see trace_expand_superset): one word per successful edge
constexpr uint32_t kEdgeHH = 0x80000000u, kEdgeToday = 0x40000000u, kEdgeTarget = 0x3FFFFFFFu, kEdgeNotCached = 0xFFFFFFFFu;
struct TraceView {
    int rep, t;
    uint32_t root;
    uint64_t seed;
    unsigned long long *winner;
    uint32_t *inbox;
    AgentRec *rec;
    uint32_t *trc;
    unsigned long long *tmark;
    uint32_t *tally_day;  // tally rows of day t
    const uint32_t *eoff, *ecnt, *edge;  // the day's edge cache (parallel tracer; nullptr: evaluate the edges from the log)
};

__device__ __forceinline__ unsigned long long tkey(int t, uint32_t root) { return ((unsigned long long)t << 32) | ((unsigned long long)root + 1ull); }

// Q[t,x] = Documented[t,x] = traced[x] = True, time_documented[x] = t  (:73-76, :84-87), by ONE lane
// (w = winner[x] and, unless x was infected today, r = rec[x] were loaded by the caller while it evaluated the edge:
// nobody else writes them during this root's search)
__device__ __noinline__ void trace_mark(const DevCtx &cx, const TraceView &tv, uint32_t x, unsigned long long w, AgentRec r)
{
    const int t = tv.t;
    if (!SEIR_CHECK((int64_t)x < cx.n, cx, tv.rep)) return;
    const int age = wday(w) == t ? (__ldg(cx.stat + x) & 127) : (int)(r.stat() & 127);
    if (wday(w) == t) {  // infected today: no record before tomorrow's pass; the marks travel in winner / inbox
        const unsigned long long old = atomicOr(tv.winner + x, W_TRACED);
        atomicOr(tv.inbox + (x >> 4), inbox_q_bit((int)(x & 15)));
        if (!(old & W_TRACED)) atomicAdd(tv.tally_day + TAL_NEW_DOC * kAges + age, 1u);
        return;
    }
    uint32_t st = r.state();
    const uint32_t fl = r.flags();
    const int comp = (int)(st & 7u);
    uint32_t tw = (fl & F_TRX) ? __ldcg(tv.trc + x) : (uint32_t)NEVER;  // t_q2 | t_doc_last << 16
    if (!(st & ST_DOC)) {
        st |= ST_DOC;
        r.set_t_documented((uint16_t)t);
        atomicAdd(tv.tally_day + TAL_NEW_DOC * kAges + age, 1u);
    }
    if (comp == C_R || comp == C_D) {
        // the agent's own iteration cleared Q[t,x] if it ended today (:278, :326); the tracer's write stays
        // only if the tracer's root is iterated later (root > x)
        const bool sticks = (int)r.t_end() != t || tv.root > x;
        if (sticks && (tw & 0xFFFFu) == (uint32_t)NEVER) {
            tw = (tw & 0xFFFF0000u) | (uint32_t)t;
            st |= ST_Q;
            atomicAdd(tv.tally_day + TAL_NEW_Q2 * kAges + age, 1u);
        }
    } else if (!(st & ST_Q)) {  // an E / Mild / Severe / Critical agent: one more quarantined active agent from today on
        st |= ST_Q;
        if (r.t_q_on() == NEVER) r.set_t_q_on((uint16_t)t);
        atomicAdd(tv.tally_day + TAL_D_Q * kAges + age, 1u);
    }
    tw = (tw & 0xFFFFu) | ((uint32_t)t << 16);
    r.set_state(st);
    r.set_flags(fl | F_TRX);
    store_rec(tv.rec + x, r);
    __stcg(tv.trc + x, tw);
}

// (log entries of one agent staged in shared memory: 504; the reference keeps 100)
#ifndef SEIR_TRACE_LOGCAP
#define SEIR_TRACE_LOGCAP 504
#endif
constexpr int kTraceLogCap = SEIR_TRACE_LOGCAP;
struct TraceSmem {
    uint32_t ent_x[kTraceLogCap];
    uint32_t ent_key[kTraceLogCap];  // day<<16 | ordinal: the order of infected_by[c, :]
    uint32_t n, pad[15];             // length of the warp's search list
};

// One node of a root's search, by one warp: evaluate the edges of c with the reference's guards, claim and
// mark the contacts reached, append them to the search list.  Within ONE root the order of the visits does not
// matter (the visited set is a closure and every mark is idempotent), so the list is used as a queue and
// several warps may expand different nodes of the same root concurrently.  Returns false on list overflow.
__device__ __forceinline__ bool trace_expand_exact(const DevCtx &cx, const TraceView &tv, TraceSmem &ts, uint32_t c,
                                                   uint32_t *list, uint32_t cap, uint32_t *n_ptr)
{
    const int lane = threadIdx.x & 31;
    const int t = tv.t;
    const uint32_t root = tv.root;
    const int td = (int)cx.tday0 + t;  // the day stamp of this run
    const unsigned long long key = tkey(td, root);
    const uint32_t *log_head = cx.log_head + (size_t)tv.rep * cx.n_pad;
    const uint4 *log_ent = cx.log_ent + (size_t)tv.rep * cx.log_cap;
    bool ok_all = true;
    // ---- the edges whose draws succeeded, kept by the frontier step of the parallel tracer: only the guards that
    // depend on the root and on what earlier roots did remain to be checked
    const uint32_t n_cached = tv.ecnt ? __ldcg(tv.ecnt + c) : kEdgeNotCached;
    if (n_cached != kEdgeNotCached) {
        const uint32_t *edges = tv.edge + __ldcg(tv.eoff + c);
        for (uint32_t e0 = 0; e0 < n_cached; e0 += 32u) {
            const uint32_t e = e0 + (uint32_t)lane;
            bool ok = false;
            uint32_t target = 0;
            unsigned long long w = 0ull;
            AgentRec mrec;
            mrec.clear();
            if (e < n_cached) {
                const uint32_t ed = __ldcg(edges + e);
                target = ed & kEdgeTarget;
                if (ed & kEdgeHH) {  // (:68-77) not S[t-1, contact] held when the edge was kept; traced[contact] is today's state
                    w = __ldcg(tv.winner + target);
                    mrec = load_rec(tv.rec + target);
                    bool traced = (mrec.state() & ST_DOC) != 0;
                    if (traced && (int)mrec.t_documented() == t && target > root &&
                        (int)(__ldcg(tv.tmark + target) >> 32) != td)
                        traced = false;
                    ok = !traced;
                } else if (!(ed & kEdgeToday) || c < root) {  // (:79-88) today's entries exist only if c's iteration came first
                    w = __ldcg(tv.winner + target);
                    if (wday(w) != t) mrec = load_rec(tv.rec + target);
                    ok = true;
                }
            }
            if (ok) {
                const unsigned long long old = atomicExch(tv.tmark + target, key);
                if (old == key) ok = false;  // already reached under this root
                else trace_mark(cx, tv, target, w, mrec);
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            if (m) {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(n_ptr, (uint32_t)__popc(m));
                at = __shfl_sync(0xFFFFFFFFu, at, 0);
                if (ok) {
                    const uint32_t mine = at + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    if (mine < cap) __stcg(list + mine, target);
                }
                if (at + (uint32_t)__popc(m) > cap) ok_all = false;
            }
            __syncwarp();
        }
        return ok_all;
    }
    const uint16_t stt = __ldg(cx.stat + c);
    const int pos = (stt >> 9) & 7, size = ((stt >> 12) & 7) + 1;
    const uint32_t base = c - (uint32_t)pos;
    // ---- infected_by[c, :]: walk the list (lane 0), newest entry first
    uint32_t L = 0;
    if (lane == 0) {
        uint32_t e = __ldcg(log_head + c);
        while (e != 0u && L < (uint32_t)kTraceLogCap) {
            const uint4 en = __ldcg(log_ent + (e - 1u));
            ts.ent_x[L] = en.x;
            ts.ent_key[L] = en.z;
            ++L;
            e = en.y;
        }
        if (e != 0u) atomicOr(&cx.counters[tv.rep * 4 + 0], 8ull);  // more entries than the staging holds: seir_run reports it
    }
    L = __shfl_sync(0xFFFFFFFFu, L, 0);
    __syncwarp();
    const uint32_t n_edges = (uint32_t)(size - 1) + L;
    for (uint32_t e0 = 0; e0 < n_edges; e0 += 32u) {
        const uint32_t e = e0 + (uint32_t)lane;
        bool ok = false;
        uint32_t target = 0;
        unsigned long long w = 0ull;
        AgentRec mrec;
        mrec.clear();
        if (e < (uint32_t)(size - 1)) {  // household edge j (:68-77)
            const int j = (int)e;
            target = base + (uint32_t)j + (j >= pos ? 1u : 0u);
            w = __ldcg(tv.winner + target);
            if (w != 0ull && wday(w) != t) {  // not S[t-1, contact]
                // traced[contact]: documented before today, or by its own iteration today if that came first
                // (contact <= root), or reached by a tracer today
                mrec = load_rec(tv.rec + target);
                bool traced = (mrec.state() & ST_DOC) != 0;
                if (traced && (int)mrec.t_documented() == t && target > root &&
                    (int)(__ldcg(tv.tmark + target) >> 32) != td)
                    traced = false;
                if (!traced) ok = sd_lane(draw_block(tv.seed, c, (uint32_t)t, SD_SITE_TR_HH, (uint32_t)j), 0) < cx.p_trace_hh;
            }
        } else if (e < n_edges) {  // outside edge (:79-88); slot = rank of the entry in (day, ordinal) order
            const uint32_t k = e - (uint32_t)(size - 1);
            const uint32_t kk = ts.ent_key[k];
            uint32_t rank = 0;
            for (uint32_t q = 0; q < L; ++q) rank += ts.ent_key[q] < kk ? 1u : 0u;
            target = ts.ent_x[k];
            // today's entries exist only if c's iteration has run before the root's (c < root)
            if (rank < (uint32_t)kLogWidth && ((int)(kk >> 16) < t || c < root)) {
                w = __ldcg(tv.winner + target);  // (for trace_mark: in flight while the draw is computed)
                if (wday(w) != t) mrec = load_rec(tv.rec + target);
                ok = sd_lane(draw_block(tv.seed, c, (uint32_t)t, SD_SITE_TR_OUT, rank), 0) < cx.p_trace_out;
            }
        }
        if (ok) {
            const unsigned long long old = atomicExch(tv.tmark + target, key);
            if (old == key) ok = false;  // already reached under this root
            else trace_mark(cx, tv, target, w, mrec);
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
        if (m) {
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(n_ptr, (uint32_t)__popc(m));
            at = __shfl_sync(0xFFFFFFFFu, at, 0);
            if (ok) {
                const uint32_t mine = at + (uint32_t)__popc(m & ((1u << lane) - 1u));
                if (mine < cap) __stcg(list + mine, target);
            }
            if (at + (uint32_t)__popc(m) > cap) ok_all = false;
        }
        __syncwarp();
    }
    return ok_all;
}

// one root, one warp
__device__ __forceinline__ bool trace_root(const DevCtx &cx, const TraceView &tv, TraceSmem &ts, uint32_t *list, uint32_t cap)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        __stcg(tv.tmark + tv.root, tkey((int)cx.tday0 + tv.t, tv.root));
        __stcg(list, tv.root);
        ts.n = 1;
    }
    __syncwarp();
    bool ok_all = true;
    for (uint32_t lo = 0;; ++lo) {
        const uint32_t n = min(*(volatile uint32_t *)&ts.n, cap);
        if (lo >= n) break;
        ok_all = trace_expand_exact(cx, tv, ts, __ldcg(list + lo), list, cap, &ts.n) && ok_all;
        if (lane == 0) SEIR_TSTAT_ADD(5, 1);
        __syncwarp();
    }
    return ok_all;
}

// one root, the whole CTA: level-synchronous over the search list (a root whose search covers thousands of agents)
__device__ __forceinline__ bool trace_root_cta(const DevCtx &cx, const TraceView &tv, TraceSmem &ts, uint32_t *list, uint32_t cap,
                                               uint32_t *n_shared)
{
    const int wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        __stcg(tv.tmark + tv.root, tkey((int)cx.tday0 + tv.t, tv.root));
        __stcg(list, tv.root);
        *n_shared = 1;
        SEIR_TSTAT_ADD(7, 1);
    }
    __syncthreads();
    bool ok_all = true;
    uint32_t lo = 0;
    for (;;) {
        const uint32_t hi = min(*(volatile uint32_t *)n_shared, cap);
        __syncthreads();  // (everybody has read the level's end before anybody appends)
        if (lo >= hi) break;
        if (threadIdx.x == 0) { SEIR_TSTAT_ADD(8, 1); SEIR_TSTAT_ADD(4, hi - lo); }
        for (uint32_t idx = lo + (uint32_t)wib; idx < hi; idx += (uint32_t)(kThreads / 32))
            ok_all = trace_expand_exact(cx, tv, ts, __ldcg(list + idx), list, cap, n_shared) && ok_all;
        lo = hi;
        __threadfence_block();
        __syncthreads();
    }
    return __syncthreads_and(ok_all ? 1 : 0) != 0;
}

// all roots of day t of the replicas this CTA serves (warp 0 only; no block-wide synchronisation inside)
__device__ __noinline__ void trace_pass(const DevCtx &cx, TraceSmem &ts, int t)
{
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    for (int rep = blockIdx.x; rep < cx.n_rep; rep += gridDim.x) {
        if (__ldcg(cx.root_cnt + (size_t)rep * (cx.T + 2) + t) == 0u) continue;
        TraceView tv;
        tv.rep = rep; tv.t = t; tv.root = 0;
        tv.seed = cx.seeds[rep];
        tv.winner = cx.winner + (size_t)rep * cx.n_pad;
        tv.inbox = cx.inbox + (size_t)rep * (cx.n_pad / 16);
        tv.rec = cx.rec + (size_t)rep * cx.n_pad;
        tv.trc = cx.trc + (size_t)rep * cx.n_pad;
        tv.tmark = cx.tmark + (size_t)rep * cx.n_pad;
        tv.tally_day = cx.tally + ((size_t)rep * cx.T + t) * TAL_ROWS * kAges;
        tv.eoff = tv.ecnt = tv.edge = nullptr;  // (no frontier step here: every edge is evaluated from the log)
        uint32_t *bits = cx.rootbits + (size_t)rep * (cx.n_pad / 32);
        uint32_t *sums = cx.rootsum + (size_t)rep * (cx.n_pad / 1024);
        uint32_t *stack = cx.tstack + (size_t)rep * cx.n_pad;
        const uint32_t n_sum = (uint32_t)(cx.n_pad / 1024);
        bool ok = true;
        for (uint32_t s0 = 0; s0 < n_sum; s0 += 32u) {  // roots in ascending agent order
            const uint32_t si = s0 + (uint32_t)lane;
            const uint32_t sw = si < n_sum ? __ldcg(sums + si) : 0u;
            unsigned any = __ballot_sync(0xFFFFFFFFu, sw != 0u);
            while (any) {
                const int src = __ffs(any) - 1;
                any &= any - 1u;
                uint32_t sword = __shfl_sync(0xFFFFFFFFu, sw, src);
                const uint32_t sidx = s0 + (uint32_t)src;
                while (sword) {
                    const int wb = __ffs(sword) - 1;
                    sword &= sword - 1u;
                    const uint32_t widx = sidx * 32u + (uint32_t)wb;
                    uint32_t word = __ldcg(bits + widx);
                    while (word) {
                        const int b = __ffs(word) - 1;
                        word &= word - 1u;
                        tv.root = widx * 32u + (uint32_t)b;
                        ok = trace_root(cx, tv, ts, stack, (uint32_t)cx.n_pad) && ok;
                    }
                    if (lane == 0) __stcg(bits + widx, 0u);
                }
                if (lane == 0) __stcg(sums + sidx, 0u);
            }
        }
        if (!ok && lane == 0) atomicOr(&cx.counters[rep * 4 + 0], 2ull);  // stack overflow: seir_run reports it
    }
}

// grid-wide barrier on a monotonically increasing counter (cooperative launch guarantees residency): thread 0 of
// every CTA arrives with ONE release reduction (nothing to wait for) and polls the same word with acquire loads.
// What counts is the time from the LAST arrival to the release of everybody, and the CTAs that arrive early have long
// finished by then.  The two-level variant (groups of kBarGroup CTAs on their own lines, the last of a group arrives
// on the global word) needs two dependent atomics on that path and measured slower.  `target` = barrier number *
// gridDim.x; word 0 of the buffer is the counter of the sharded barrier, words 32.. belong to this one.
#ifndef SEIR_FLAT_BARRIER
#define SEIR_FLAT_BARRIER 1   // (measured on C2: day loop 1.391 ms flat against 1.430 ms two-level)
#endif

constexpr unsigned int kBarGroup = 16, kBarStride = 32;
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int epoch = target / gridDim.x;
        unsigned int seen;
#if SEIR_FLAT_BARRIER
        // one word: a release reduction (no return value to wait for) and an acquire poll
        unsigned int *global = counter + kBarStride;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(global), "r"(1u) : "memory");
        do {  // (polling with relaxed loads and one fence at the end measured slower: 1.400 against 1.385 ms)
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(global) : "memory");
        } while (seen < epoch * gridDim.x);
#else
        const unsigned int g = blockIdx.x / kBarGroup, n_groups = (gridDim.x + kBarGroup - 1u) / kBarGroup;
        const unsigned int size = min(kBarGroup, gridDim.x - g * kBarGroup);
        unsigned int *global = counter + kBarStride, *group = counter + (2u + g) * kBarStride;
        unsigned int before;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(before) : "l"(group), "r"(1u) : "memory");
        if (before + 1u == epoch * size)  // the last one of the group
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(global), "r"(1u) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(global) : "memory");
        } while (seen < epoch * n_groups);
#endif
    }
    __syncthreads();
}

// The morning of a sharded day t: every GPU brings ITS copy of xday up to date from the owners' lists of today's new
// infectees (the agents of owner g that anybody infected yesterday; complete since last night's cross-GPU barrier): bulk,
// coalesced reads over NVLink instead of a remote read per candidate contact.  The caller meets at a grid barrier
// afterwards: no contact test of day t may run before the copy is complete.
// Returns false when nobody was infected yesterday (grid-uniform: every CTA reads the same counts): nothing to wait for.
__device__ __forceinline__ bool xday_refresh(const DevCtx &cx, PassSmem &sm, int t)
{
    if (threadIdx.x < (unsigned)cx.n_shards) sm.own_n[threadIdx.x] = __ldcg(cx.peer_list_cnt[threadIdx.x] + (cx.T + 2) + t);
    __syncthreads();
    uint32_t any = 0;
    for (int g = 0; g < cx.n_shards; ++g) any |= sm.own_n[g];
    if (!any) {
        __syncthreads();
        return false;
    }
    const size_t lcur = (size_t)(t & 1) * cx.n_pad;
    for (int g = 0; g < cx.n_shards; ++g) {
        const uint32_t cnt = sm.own_n[g];
        const uint32_t *row = cx.peer_list_new[g] + lcur;
        // (16 B per thread and NVLink round trip: the rows start on a 16 B boundary)
        const uint32_t nv = (cnt + 3u) >> 2;
        for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q < nv; q += gridDim.x * kThreads) {
            const uint4 e = __ldcg(reinterpret_cast<const uint4 *>(row) + q);
            const uint32_t c4[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (q * 4u + (uint32_t)j < cnt && c4[j] != kHole && (int64_t)c4[j] < cx.n) __stcg(cx.xday + c4[j], (uint16_t)t);  // infected on day t-1
        }
    }
    __syncthreads();
    return true;
}

// ---------------------------------------------------------------------------------------
// Parallel tracing for the persistent kernel.  The sequential order of the reference only matters between
// roots whose searches can touch the same agents.  So: (1) all roots of the day grow a union-find forest
// over the SUPERSET graph of every edge any search order could follow (household edge: draw succeeds, the
// member is not S and was not documented before today; outside edge: draw succeeds, today's entries
// included) with a level-synchronous frontier; (2) every connected component elects its smallest root as
// leader; (3) one warp per component runs the exact ordered search (trace_root, above) over the
// component's roots in ascending agent order.  Components share no agent, so they run concurrently and the
// result is the reference's, bit for bit.  Needs grid barriers between the steps: every thread of the grid
// calls trace_parallel().
struct TraceCtl {          // per replica
    uint32_t n_roots, n_reached, alloc, n_big;
    uint32_t edges;        // slots of the edge cache handed out today
    uint32_t appended[2];  // nodes appended to the frontier by the even / odd levels of the day (cumulative)
};

__device__ __forceinline__ uint32_t uf_find(unsigned long long *comp, uint32_t x, unsigned long long stamp)
{
    for (;;) {
        const uint32_t p = (uint32_t)__ldcg(comp + x);
        if (p == x) return x;
        const uint32_t gp = (uint32_t)__ldcg(comp + p);
        if (gp != p) atomicCAS(comp + x, stamp | p, stamp | gp);  // path halving
        x = p;
    }
}
__device__ __forceinline__ void uf_union(unsigned long long *comp, uint32_t a, uint32_t b, unsigned long long stamp)
{
    for (;;) {
        a = uf_find(comp, a, stamp);
        b = uf_find(comp, b, stamp);
        if (a == b) return;
        if (a < b) { const uint32_t s = a; a = b; b = s; }
        if (atomicCAS(comp + a, stamp | a, stamp | b) == (stamp | a)) return;  // the larger id hangs under the smaller
    }
}

struct TracePar {          // per-replica arrays of the parallel tracer
    unsigned long long *comp, *lead;   // [n_pad] day<<32 | union-find parent ; day<<32 | smallest root slot of the component
    uint32_t *reached, *roots, *reps, *csz;  // [n_pad] each: frontier / reached nodes, the day's roots (ascending), their representatives, component sizes
    TraceCtl *ctl;
};

// one node of the frontier: evaluate its superset edges with the warp, claim / unite the targets
__device__ __forceinline__ void trace_expand_superset(const DevCtx &cx, const TraceView &tv, const TracePar &tp, TraceSmem &ts,
                                                      uint32_t c, int parity, unsigned long long *total_appended)
{
    const int lane = threadIdx.x & 31;
    const int t = tv.t;
    const unsigned long long td = (unsigned long long)cx.tday0 + (unsigned long long)t;
    const unsigned long long stamp = td << 32;
    const uint32_t *log_head = cx.log_head + (size_t)tv.rep * cx.n_pad;
    const uint4 *log_ent = cx.log_ent + (size_t)tv.rep * cx.log_cap;
    const uint16_t stt = __ldg(cx.stat + c);
    const int pos = (stt >> 9) & 7, size = ((stt >> 12) & 7) + 1;
    const uint32_t base = c - (uint32_t)pos;
    uint32_t L = 0;
    if (lane == 0) {
        uint32_t e = __ldcg(log_head + c);
        while (e != 0u && L < (uint32_t)kTraceLogCap) {
            const uint4 en = __ldcg(log_ent + (e - 1u));
            ts.ent_x[L] = en.x;
            ts.ent_key[L] = en.z;
            ++L;
            e = en.y;
        }
        if (e != 0u) atomicOr(&cx.counters[tv.rep * 4 + 0], 8ull);  // more entries than the staging holds: seir_run reports it
    }
    L = __shfl_sync(0xFFFFFFFFu, L, 0);
    __syncwarp();
    const uint32_t n_edges = (uint32_t)(size - 1) + L;
    // The successful edges of c are kept for the searches of the roots (trace_expand_exact): every root that reaches c
    // would otherwise walk c's log, rank its entries and draw again.  Room for all n_edges is reserved up front (the
    // successes are written compactly at its start); when the cache is full the searches fall back to the log.
    uint32_t ebase = kEdgeNotCached, ekept = 0;
    if (lane == 0 && n_edges != 0u) {
        const uint32_t at = atomicAdd(&tp.ctl->edges, n_edges);
        if (at + n_edges <= cx.tedge_cap) ebase = at;
    }
    ebase = __shfl_sync(0xFFFFFFFFu, ebase, 0);
    uint32_t *epool = cx.tedge + (size_t)tv.rep * cx.tedge_cap;
    for (uint32_t e0 = 0; e0 < n_edges; e0 += 32u) {
        const uint32_t e = e0 + (uint32_t)lane;
        bool ok = false;
        uint32_t target = 0, eflags = 0;
        if (e < (uint32_t)(size - 1)) {
            const int j = (int)e;
            target = base + (uint32_t)j + (j >= pos ? 1u : 0u);
            eflags = kEdgeHH;
            const unsigned long long w = __ldcg(tv.winner + target);
            if (w != 0ull && wday(w) != t) {
                const AgentRec m = load_rec(tv.rec + target);
                const bool before_today = (m.state() & ST_DOC) != 0 && (int)m.t_documented() != t;
                if (!before_today) ok = sd_lane(draw_block(tv.seed, c, (uint32_t)t, SD_SITE_TR_HH, (uint32_t)j), 0) < cx.p_trace_hh;
            }
        } else if (e < n_edges) {
            const uint32_t k = e - (uint32_t)(size - 1);
            const uint32_t kk = ts.ent_key[k];
            uint32_t rank = 0;
            for (uint32_t q = 0; q < L; ++q) rank += ts.ent_key[q] < kk ? 1u : 0u;
            target = ts.ent_x[k];
            if ((int)(kk >> 16) >= t) eflags = kEdgeToday;
            if (rank < (uint32_t)kLogWidth)
                ok = sd_lane(draw_block(tv.seed, c, (uint32_t)t, SD_SITE_TR_OUT, rank), 0) < cx.p_trace_out;
        }
        {
            const unsigned mk = __ballot_sync(0xFFFFFFFFu, ok);
            if (ok && ebase != kEdgeNotCached) __stcg(epool + ebase + ekept + (uint32_t)__popc(mk & ((1u << lane) - 1u)), target | eflags);
            ekept += (uint32_t)__popc(mk);
        }
        bool fresh = false;
        if (ok) {
            const unsigned long long old = __ldcg(tp.comp + target);
            if ((old >> 32) != td && atomicCAS(tp.comp + target, old, stamp | target) == old) {
                fresh = true;
                __stcg(tp.lead + target, stamp | 0xFFFFFFFFull);
            }
            uf_union(tp.comp, c, target, stamp);
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, fresh);
        if (m) {
            uint32_t at = 0;
            if (lane == 0) {
                at = atomicAdd(&tp.ctl->n_reached, (uint32_t)__popc(m));
                atomicAdd(&tp.ctl->appended[parity], (uint32_t)__popc(m));
                atomicAdd(total_appended, (unsigned long long)__popc(m));
            }
            at = __shfl_sync(0xFFFFFFFFu, at, 0);
            if (fresh) __stcg(tp.reached + at + (uint32_t)__popc(m & ((1u << lane) - 1u)), target);
        }
        __syncwarp();
    }
    if (lane == 0) {
        __stcg(cx.tedge_off + (size_t)tv.rep * cx.n_pad + c, ebase);
        __stcg(cx.tedge_cnt + (size_t)tv.rep * cx.n_pad + c, ebase != kEdgeNotCached || n_edges == 0u ? ekept : kEdgeNotCached);
    }
}

// the day's roots of one replica in ascending agent order (three-level bitmap), by ONE warp; resets the bitmaps
__device__ __forceinline__ uint32_t trace_collect_roots(const DevCtx &cx, int rep, const TracePar &tp, unsigned long long stamp)
{
    const int lane = threadIdx.x & 31;
    uint32_t *bits = cx.rootbits + (size_t)rep * (cx.n_pad / 32);
    uint32_t *sums = cx.rootsum + (size_t)rep * (cx.n_pad / 1024);
    const uint32_t n_sum2 = (uint32_t)((cx.n_pad + 32767) / 32768);
    uint32_t *sums2 = cx.rootsum2 + (size_t)rep * n_sum2;
    uint32_t R = 0;
    for (uint32_t s0 = 0; s0 < n_sum2; s0 += 32u) {
        const uint32_t w2 = s0 + (uint32_t)lane < n_sum2 ? __ldcg(sums2 + s0 + lane) : 0u;
        unsigned any2 = __ballot_sync(0xFFFFFFFFu, w2 != 0u);
        while (any2) {
            const int src2 = __ffs(any2) - 1;
            any2 &= any2 - 1u;
            const uint32_t word2 = __shfl_sync(0xFFFFFFFFu, w2, src2);
            const uint32_t s2idx = s0 + (uint32_t)src2;
            const uint32_t sidx = s2idx * 32u + (uint32_t)lane;
            const bool have1 = (word2 >> lane) & 1u;
            const uint32_t sw = have1 ? __ldcg(sums + sidx) : 0u;
            unsigned any1 = __ballot_sync(0xFFFFFFFFu, sw != 0u);
            while (any1) {
                const int src1 = __ffs(any1) - 1;
                any1 &= any1 - 1u;
                const uint32_t word1 = __shfl_sync(0xFFFFFFFFu, sw, src1);
                const uint32_t widx = (s2idx * 32u + (uint32_t)src1) * 32u + (uint32_t)lane;
                const bool have0 = (word1 >> lane) & 1u;
                uint32_t bw = have0 ? __ldcg(bits + widx) : 0u;
                const uint32_t cnt = (uint32_t)__popc(bw);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
                uint32_t at = R + incl - cnt;
                while (bw) {
                    const int b = __ffs(bw) - 1;
                    bw &= bw - 1u;
                    const uint32_t root = widx * 32u + (uint32_t)b;
                    __stcg(tp.roots + at, root);
                    __stcg(tp.reached + at, root);
                    __stcg(tp.csz + at, 0u);
                    __stcg(tp.comp + root, stamp | root);
                    __stcg(tp.lead + root, stamp | 0xFFFFFFFFull);
                    ++at;
                }
                R += tot;
                if (have0) __stcg(bits + widx, 0u);
            }
            if (have1) __stcg(sums + sidx, 0u);
        }
        if (w2) __stcg(sums2 + s0 + lane, 0u);
    }
    return R;
}

constexpr int kMaxRepsPerCta = 16;
#ifndef SEIR_BIG_COMPONENT
#define SEIR_BIG_COMPONENT 16   // (measured on C2 with tracing from day 37: 96 -> 33.1 ms per run, 48 -> 30.8, 24 -> 29.5, 12 -> 29.3, 6 -> 28.9)
#endif
constexpr int kBigComponent = SEIR_BIG_COMPONENT;  // agents: larger components are searched by a whole CTA
constexpr int kBigCap = 16384;     // such components per replica and day (more: searched by their leader's warp)
__device__ __noinline__ void trace_parallel(const DevCtx &cx, PassSmem &sm, int t, unsigned int &epoch)
{
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    TraceSmem &ts = reinterpret_cast<TraceSmem *>(sm.out_old)[wib];  // 16 x 4 KB over the pass queues (idle between passes)
    static_assert(sizeof(TraceSmem) * (kThreads / 32) <= offsetof(PassSmem, tally) - offsetof(PassSmem, out_old),
                  "the tracing scratch must fit over the pass queues");
    const int bpr = cx.blocks_per_rep;
    const int reps_per_sweep = gridDim.x / bpr;
    const int my_slot = blockIdx.x / bpr, my_part = blockIdx.x % bpr;
    const bool serving = my_slot < reps_per_sweep;
    const uint32_t gw = (uint32_t)my_part * (uint32_t)(kThreads / 32) + (uint32_t)wib, ngw = (uint32_t)bpr * (uint32_t)(kThreads / 32);
    const unsigned long long stamp = ((unsigned long long)cx.tday0 + (unsigned long long)t) << 32;
    unsigned long long *total = cx.ttotal;

    auto make = [&](int rep, TraceView &tv, TracePar &tp) {
        tv.rep = rep; tv.t = t; tv.root = 0;
        tv.seed = cx.seeds[rep];
        tv.winner = cx.winner + (size_t)rep * cx.n_pad;
        tv.inbox = cx.inbox + (size_t)rep * (cx.n_pad / 16);
        tv.rec = cx.rec + (size_t)rep * cx.n_pad;
        tv.trc = cx.trc + (size_t)rep * cx.n_pad;
        tv.tmark = cx.tmark + (size_t)rep * cx.n_pad;
        tv.tally_day = cx.tally + ((size_t)rep * cx.T + t) * TAL_ROWS * kAges;
        tv.eoff = cx.tedge_off + (size_t)rep * cx.n_pad;
        tv.ecnt = cx.tedge_cnt + (size_t)rep * cx.n_pad;
        tv.edge = cx.tedge + (size_t)rep * cx.tedge_cap;
        tp.comp = cx.tcomp + (size_t)rep * cx.n_pad;
        tp.lead = cx.tlead + (size_t)rep * cx.n_pad;
        tp.reached = cx.treached + (size_t)rep * cx.n_pad;
        tp.roots = cx.troots + (size_t)rep * cx.n_pad;
        tp.reps = cx.treps + (size_t)rep * cx.n_pad;
        tp.csz = cx.tcsz + (size_t)rep * cx.n_pad;
        tp.ctl = cx.tctl + rep;
    };
    TraceView tv;
    TracePar tp;

    // ---- roots, in ascending agent order: every warp of the replica's CTAs owns a contiguous slice of the root
    // bitmap; count, exchange the CTA totals (one grid barrier), then write at the prefix offsets and clear the bits.
    // (A CTA serves at most kMaxRepsPerCta replicas here; the host falls back to trace_pass beyond that.)
    uint32_t *warp_cnt = sm.tally;              // [kMaxRepsPerCta][kWarps + 1] scratch, free between passes
    const uint32_t n_words = (uint32_t)(cx.n_pad / 32);
    const uint32_t words_per_warp = ((n_words + ngw - 1u) / ngw + 31u) & ~31u;
    const uint32_t w_lo = min(gw * words_per_warp, n_words), w_hi = min(w_lo + words_per_warp, n_words);
    // (Words that every warp of a replica's CTAs needs -- root counts, CTA totals, frontier bounds, the stop test -- are
    // read by ONE thread per CTA and handed on through shared memory: thousands of warps asking for the same L2 line
    // are served one at a time by its slice, microseconds per read.)
    uint32_t *bc = sm.tally + 320;              // [2 + 2 * kMaxRepsPerCta] broadcast scratch (warp_cnt ends at 272)
    if (serving && tid < kMaxRepsPerCta) {
        const int rep = my_slot + tid * reps_per_sweep;
        bc[2 + tid] = rep < cx.n_rep ? __ldcg(cx.root_cnt + (size_t)rep * (cx.T + 2) + t) : 0u;
    }
    __syncthreads();
    if (serving) {
        int k = 0;
        for (int rep = my_slot; rep < cx.n_rep && k < kMaxRepsPerCta; rep += reps_per_sweep, ++k) {
            uint32_t cnt = 0;
            if (bc[2 + k] != 0u) {
                const uint32_t *bits = cx.rootbits + (size_t)rep * n_words;
                for (uint32_t w = w_lo + (uint32_t)lane; w < w_hi; w += 32u) cnt += (uint32_t)__popc(__ldcg(bits + w));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            }
            if (lane == 0) warp_cnt[k * (kWarps + 1) + wib] = cnt;
        }
        __syncthreads();
        if (tid < k) {  // one thread per replica of this CTA: the CTA's total
            uint32_t tot = 0;
            for (int w = 0; w < kWarps; ++w) tot += warp_cnt[tid * (kWarps + 1) + w];
            warp_cnt[tid * (kWarps + 1) + kWarps] = tot;
            __stcg(cx.tcta + (size_t)(my_slot + tid * reps_per_sweep) * gridDim.x + my_part, tot);
        }
    }
    grid_barrier(cx.barrier, ++epoch * gridDim.x);
    if (serving) {
        if (wib == 0) {  // the CTAs before mine and the replica's total, per replica of this CTA
            int k = 0;
            for (int rep = my_slot; rep < cx.n_rep && k < kMaxRepsPerCta; rep += reps_per_sweep, ++k) {
                uint32_t before = 0, all = 0;
                for (int p0 = 0; p0 < bpr; p0 += 32) {
                    const int pp = p0 + lane;
                    const uint32_t v = pp < bpr ? __ldcg(cx.tcta + (size_t)rep * gridDim.x + pp) : 0u;
                    all += v;
                    if (pp < my_part) before += v;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    before += __shfl_xor_sync(0xFFFFFFFFu, before, o);
                    all += __shfl_xor_sync(0xFFFFFFFFu, all, o);
                }
                if (lane == 0) { bc[2 + 2 * k] = before; bc[3 + 2 * k] = all; }
            }
        }
        __syncthreads();
        int k = 0;
        for (int rep = my_slot; rep < cx.n_rep && k < kMaxRepsPerCta; rep += reps_per_sweep, ++k) {
            make(rep, tv, tp);
            // offset of this warp: the CTAs before mine, then the warps before mine
            uint32_t before = bc[2 + 2 * k];
            const uint32_t all = bc[3 + 2 * k];
            for (int w = 0; w < wib; ++w) before += warp_cnt[k * (kWarps + 1) + w];
            if (my_part == 0 && tid == 0) {
                tp.ctl->n_roots = all;
                tp.ctl->n_reached = all;
                tp.ctl->alloc = 0;
                tp.ctl->n_big = 0;
                tp.ctl->appended[0] = 0;
                tp.ctl->appended[1] = 0;
                tp.ctl->edges = 0;
            }
            if (warp_cnt[k * (kWarps + 1) + wib] == 0u) continue;
            uint32_t *bits = cx.rootbits + (size_t)rep * n_words;
            uint32_t at0 = before;
            for (uint32_t w0 = w_lo; w0 < w_hi; w0 += 32u) {
                const uint32_t w = w0 + (uint32_t)lane;
                uint32_t bw = w < w_hi ? __ldcg(bits + w) : 0u;
                const uint32_t cnt = (uint32_t)__popc(bw);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += v;
                }
                uint32_t at = at0 + incl - cnt;
                at0 += __shfl_sync(0xFFFFFFFFu, incl, 31);
                if (bw) __stcg(bits + w, 0u);
                while (bw) {
                    const int b = __ffs(bw) - 1;
                    bw &= bw - 1u;
                    const uint32_t root = w * 32u + (uint32_t)b;
                    __stcg(tp.roots + at, root);
                    __stcg(tp.reached + at, root);
                    __stcg(tp.csz + at, 0u);
                    __stcg(tp.comp + root, stamp | root);
                    __stcg(tp.lead + root, stamp | 0xFFFFFFFFull);
                    ++at;
                }
            }
        }
    }
    grid_barrier(cx.barrier, ++epoch * gridDim.x);
#define SEIR_TSTAMP(k) do { if (blockIdx.x == 0 && tid == 0) cx.stage_ns[(size_t)t * 8 + (k)] = global_timer_ns(); } while (0)
    SEIR_TSTAMP(1);

    // ---- union-find over the superset graph, level by level
    // Everything a CTA reads to steer the loop must be STABLE when it reads it, because fast CTAs (and fast warps
    // of a CTA that serves several replicas) are already appending for the next level:
    //  * the frontier of level i + 1 of a replica is [hi_i, hi_i + what level i appended); the appends of level i are
    //    counted in ctl->appended[i & 1], which nobody touches again before level i + 2.  (n_reached itself is the
    //    allocator: it may run ahead of the entries written behind it, so it is never read inside the loop);
    //  * the loop ends after a level that appended nothing anywhere, counted the same way in total[i & 1], and every
    //    CTA takes the same decision (they meet at the grid barrier).
    uint32_t lo[kMaxRepsPerCta], hi[kMaxRepsPerCta], seen[kMaxRepsPerCta][2], n_r[kMaxRepsPerCta];
    if (serving) {
        int k = 0;
        for (int rep = my_slot; rep < cx.n_rep && k < kMaxRepsPerCta; rep += reps_per_sweep, ++k) {
            lo[k] = 0;
            hi[k] = n_r[k] = bc[3 + 2 * k];  // the replica's roots (left there by the prefix step above)
            seen[k][0] = seen[k][1] = 0;
        }
    }
    unsigned long long prev_total[2] = {0ull, 0ull};  // (thread 0 only)
    if (tid == 0) { prev_total[0] = __ldcg(total); prev_total[1] = __ldcg(total + 1); }
    __syncthreads();  // (bc is rewritten below)
    for (unsigned int level = 0;; ++level) {
        const int par = (int)(level & 1u);
        unsigned long long *counter = total + par;
        if (serving) {
            int k = 0;
            for (int rep = my_slot; rep < cx.n_rep && k < kMaxRepsPerCta; rep += reps_per_sweep, ++k) {
                if (lo[k] == hi[k]) continue;
                make(rep, tv, tp);
                for (uint32_t idx = lo[k] + gw; idx < hi[k]; idx += ngw)
                    trace_expand_superset(cx, tv, tp, ts, __ldcg(tp.reached + idx), par, counter);
            }
        }
        grid_barrier(cx.barrier, ++epoch * gridDim.x);
#ifdef SEIR_TRACE_STATS
        if (blockIdx.x == 0 && tid == 0) atomicAdd(cx.fine_ns + (size_t)t * 16 + 6, 1ull);
#endif
        if (tid == 0) {
            const unsigned long long now = __ldcg(counter);
            bc[0] = now == prev_total[par] ? 1u : 0u;
            prev_total[par] = now;
        }
        if (serving && tid >= 32 && tid < 32 + kMaxRepsPerCta) {
            const int rep = my_slot + (tid - 32) * reps_per_sweep;
            if (rep < cx.n_rep) bc[2 + (tid - 32)] = __ldcg(&cx.tctl[rep].appended[par]);
        }
        __syncthreads();
        if (bc[0]) break;
        if (serving) {
            int k = 0;
            for (int rep = my_slot; rep < cx.n_rep && k < kMaxRepsPerCta; rep += reps_per_sweep, ++k) {
                const uint32_t a = bc[2 + k];
                lo[k] = hi[k];
                hi[k] += a - seen[k][par];
                seen[k][par] = a;
            }
        }
        // (the next writes to bc come after the next grid barrier, which begins with a __syncthreads)
    }
    SEIR_TSTAMP(2);

    // ---- representatives of the roots, leader (smallest root slot) of every component
    // (from here on hi[k] is the replica's n_reached -- the last level appended nothing -- and n_r[k] its n_roots)
    if (serving) {
        int kk = 0;
        for (int rep = my_slot; rep < cx.n_rep && kk < kMaxRepsPerCta; rep += reps_per_sweep, ++kk) {
            make(rep, tv, tp);
            const uint32_t R = n_r[kk];
            if (my_part == 0 && tid == 0) { SEIR_TSTAT_ADD(0, R); SEIR_TSTAT_ADD(1, hi[kk]); }
            for (uint32_t k = gw * 32u + (uint32_t)lane; k < R; k += ngw * 32u) {
                const uint32_t rp = uf_find(tp.comp, __ldcg(tp.roots + k), stamp);
                __stcg(tp.reps + k, rp);
                atomicMin(tp.lead + rp, stamp | k);
            }
        }
    }
    grid_barrier(cx.barrier, ++epoch * gridDim.x);
    SEIR_TSTAMP(3);
    // ---- component sizes (the capacity of the leader's search stack)
    if (serving) {
        int kk = 0;
        for (int rep = my_slot; rep < cx.n_rep && kk < kMaxRepsPerCta; rep += reps_per_sweep, ++kk) {
            make(rep, tv, tp);
            const uint32_t NR = hi[kk];
            for (uint32_t idx = gw * 32u + (uint32_t)lane; idx < NR; idx += ngw * 32u) {
                const uint32_t L = (uint32_t)__ldcg(tp.lead + uf_find(tp.comp, __ldcg(tp.reached + idx), stamp));
                atomicAdd(tp.csz + L, 1u);
            }
        }
    }
    grid_barrier(cx.barrier, ++epoch * gridDim.x);
    SEIR_TSTAMP(4);
    // ---- every component: its roots in ascending order -- by one warp, or by a whole CTA when the component is large
    if (serving) {
        int kk = 0;
        for (int rep = my_slot; rep < cx.n_rep && kk < kMaxRepsPerCta; rep += reps_per_sweep, ++kk) {
            make(rep, tv, tp);
            const uint32_t R = n_r[kk];
            bool ok = true;
            for (uint32_t k = gw; k < R; k += ngw) {
                const uint32_t rp = __ldcg(tp.reps + k);
                if ((uint32_t)__ldcg(tp.lead + rp) != k) continue;
                const uint32_t cap = __ldcg(tp.csz + k);
                if (lane == 0 && cap > (uint32_t)kBigComponent) { SEIR_TSTAT_ADD(2, 1); SEIR_TSTAT_MAX(3, cap); }
                if (cap > (uint32_t)kBigComponent) {  // queued for the CTA pass below
                    uint32_t pos = 0;
                    if (lane == 0) pos = atomicAdd(&tp.ctl->n_big, 1u);
                    pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
                    if (pos < (uint32_t)kBigCap) {
                        if (lane == 0) __stcg(cx.tbig + (size_t)rep * kBigCap + pos, k);
                        continue;
                    }
                }
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&tp.ctl->alloc, cap);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                uint32_t *list = cx.tstack + (size_t)rep * cx.n_pad + base;
                for (uint32_t j0 = k; j0 < R; j0 += 32u) {
                    const uint32_t j = j0 + (uint32_t)lane;
                    unsigned m = __ballot_sync(0xFFFFFFFFu, j < R && __ldcg(tp.reps + (j < R ? j : k)) == rp);
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1u;
                        tv.root = __ldcg(tp.roots + j0 + (uint32_t)b);
                        const bool ok1 = trace_root(cx, tv, ts, list, cap);
                        ok = ok1 && ok;
                    }
                }
            }
            if (!ok && lane == 0) atomicOr(&cx.counters[rep * 4 + 0], 2ull | 16ull);
        }
    }
    grid_barrier(cx.barrier, ++epoch * gridDim.x);
    SEIR_TSTAMP(5);
    if (serving) {
        uint32_t *wc = sm.tally;  // [kWarps + 2] scratch
        int kk = 0;
        for (int rep = my_slot; rep < cx.n_rep && kk < kMaxRepsPerCta; rep += reps_per_sweep, ++kk) {
            make(rep, tv, tp);
            const uint32_t R = n_r[kk];
            const uint32_t n_big = min(__ldcg(&tp.ctl->n_big), (uint32_t)kBigCap);
            bool ok = true;
            for (uint32_t q = (uint32_t)my_part; q < n_big; q += (uint32_t)bpr) {  // (CTA-uniform)
                const uint32_t k = __ldcg(cx.tbig + (size_t)rep * kBigCap + q);
                const uint32_t rp = __ldcg(tp.reps + k);
                const uint32_t cap = __ldcg(tp.csz + k);
                if (tid == 0) { sm.base_new = atomicAdd(&tp.ctl->alloc, cap); wc[kWarps + 1] = 0; }
                __syncthreads();
                const uint32_t base = sm.base_new;
                uint32_t *list = cx.tstack + (size_t)rep * cx.n_pad + base;
                uint32_t *croots = tp.reached + base;  // (the frontier array is free by now; a component has at most `cap` roots)
                // the component's roots, ascending: ordered compaction of the matching root slots
                for (uint32_t j0 = k; j0 < R; j0 += kThreads) {
                    const uint32_t j = j0 + (uint32_t)tid;
                    const bool match = j < R && __ldcg(tp.reps + j) == rp;
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, match);
                    if (lane == 0) wc[wib] = (uint32_t)__popc(m);
                    __syncthreads();
                    uint32_t before = wc[kWarps + 1];
                    for (int w = 0; w < wib; ++w) before += wc[w];
                    if (match) __stcg(croots + before + (uint32_t)__popc(m & ((1u << lane) - 1u)), __ldcg(tp.roots + j));
                    __syncthreads();
                    if (tid == 0) {
                        uint32_t tot = 0;
                        for (int w = 0; w < kWarps; ++w) tot += wc[w];
                        wc[kWarps + 1] += tot;
                    }
                    __syncthreads();
                }
                const uint32_t n_cr = wc[kWarps + 1];
                if (tid == 0) SEIR_TSTAT_MAX(9, n_cr);
                for (uint32_t j = 0; j < n_cr; ++j) {
                    tv.root = __ldcg(croots + j);
                    ok = trace_root_cta(cx, tv, ts, list, cap, &sm.base_old) && ok;
                }
                __syncthreads();
            }
            if (!ok && tid == 0) atomicOr(&cx.counters[rep * 4 + 0], 2ull | 32ull);
        }
    }
}

// Barrier over all CTAs of all GPUs of a sharded run.  Local arrival as above (each CTA fences at system
// scope first: its remote atomics / stores into the peers' memory must be visible before the peers go on);
// CTA 0 then publishes the epoch in every peer's flag array over NVLink, waits for all peers' epochs in
// its own, and releases the local CTAs.  A peer that never arrives (its launch failed) ends the wait
// after kPeerTimeoutNs and raises the error word instead of hanging the box.
constexpr unsigned long long kPeerTimeoutNs = 4000000000ull;
__device__ __forceinline__ void grid_barrier_multi(const DevCtx &cx, unsigned int *counter, unsigned int target, unsigned int epoch)
{
    unsigned int *flags = cx.peer_flags[cx.my_shard];
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.sys;" ::: "memory");  // (this CTA's stores into the peers' memory are performed before it arrives)
        atomicAdd(counter, 1u);
        unsigned int seen;
        if (blockIdx.x == 0) {
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            } while (seen < target);
            // release pattern = ONE system-scope fence, then relaxed stores: the flag stores to the peers travel side by side
            // (a st.release.sys per peer fences again each time and waits for the previous peer's store to be acknowledged:
            // seven NVLink round trips in a row on an 8-GPU box)
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int g = 0; g < cx.n_shards; ++g)
                if (g != cx.my_shard)
                    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(cx.peer_flags[g] + cx.my_shard), "r"(epoch) : "memory");
            const unsigned long long t0 = global_timer_ns();
            for (int g = 0; g < cx.n_shards; ++g) {
                if (g == cx.my_shard) continue;
                if (*(volatile unsigned int *)(flags + kMaxShards + 1)) break;  // a peer is gone: do not wait again
                do {
                    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + g) : "memory");  // (acquire = the fence below)
                    if (seen < epoch && global_timer_ns() - t0 > kPeerTimeoutNs) {
                        atomicExch(flags + kMaxShards + 1, 1u);  // error word
                        break;
                    }
                } while (seen < epoch);
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + kMaxShards), "r"(epoch) : "memory");
        } else {
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + kMaxShards) : "memory");
            } while (seen < epoch);
        }
        // (the acquire load above orders this CTA behind CTA 0, which is ordered behind the peers: no further fence)
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// packed state byte of an ever-infected agent on day t, from its record
__device__ __forceinline__ uint8_t state_on_day(const AgentRec &r, int t, int t_q2 = 0xFFFF)
{
    if (t < (int)r.t_exposed()) return C_S;
    const bool ended = r.t_end() != NEVER && t >= (int)r.t_end();
    int comp;
    if (ended) comp = r.state() & 7;
    else if ((r.flags() & F_CRIT) && t >= t_critical_of(r)) comp = C_CRIT;
    else if ((r.flags() & F_SEV) && t >= t_severe_of(r)) comp = C_SEV;
    else if ((r.flags() & F_MILD) && t >= t_infected_of(r)) comp = C_MILD;
    else comp = C_E;
    uint8_t s = (uint8_t)comp;
    if ((r.t_q_on() != NEVER && t >= (int)r.t_q_on() && !ended) || t >= t_q2) s |= ST_Q;
    if (r.t_documented() != 0 && t >= (int)r.t_documented()) s |= ST_DOC;
    return s;
}

}  // namespace seir
