// seir_kernels.cuh -- device side of the B200 SEIR day step (sm_100a).
//
// One PASS per simulated day t (1 <= t <= T), one thread per 16 agents for the dense part:
//   (a) read the packed state row t-1 (one byte per agent: bits 0-2 compartment, bit 3 Q,
//       bit 4 Documented) as 16 B vectors plus one 32-bit "inbox" word per 16 agents;
//   (b) carry the row to row t (the reference's nine row copies, seir_individual.py:245-253,
//       are this one byte);
//   (c) agents that are E/Mild/Severe/Critical at t-1, and S agents whose inbox says they were
//       infected on day t-1, are queued in shared memory and processed one per thread:
//       own progression (:256-337), household push (:339-359), community push (:361-390),
//       per-day per-age tallies (run_simulation.py:448-460 and siblings).
// All cross-agent effects of day t travel through two atomics on the INFECTEE:
//   winner[c]  = atomicMax of (day<<48 | infector<<16 | event ordinal)  -- the reference's
//                "last writer in agent order wins" for the infectee's timers;
//   inbox[c]  |= "ever infected" bit, and the sticky "isolation draw was 0 => Q" bit (:353-354).
// The infectee's own thread applies them in the NEXT pass (it reads row t as S + inbox bit +
// winner.day == t), which is why one grid-wide barrier per day is enough and why the kernel can
// run persistently over the whole horizon.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seir_b200.h"
#include "../../include/seir_draws.h"

namespace seir {

constexpr int kAges = SEIR_N_AGES;
constexpr int kThreads = 256;                 // threads per CTA
constexpr int kVec = 16;                      // agents per thread (one 16 B vector of state bytes)
constexpr int kTile = kThreads * kVec;        // agents per tile = 4096
// device tally rows (per day, per age)
enum { TAL_E = 0, TAL_MILD, TAL_SEV, TAL_CRIT, TAL_QACT, TAL_NEW_E, TAL_NEW_R, TAL_NEW_D, TAL_NEW_DOC, TAL_ROWS };
constexpr int kTalLagRows = TAL_NEW_E + 1;    // rows [0, kTalLagRows) describe day t-1, the rest day t

// compartments (bits 0-2 of the state byte)
enum { C_S = 0, C_E = 1, C_MILD = 2, C_SEV = 3, C_CRIT = 4, C_R = 5, C_D = 6, C_PAD = 7 };
constexpr uint8_t ST_Q = 0x08, ST_DOC = 0x10;
constexpr uint16_t NEVER = 0xFFFFu;

// 32 B per-agent record: one sector, touched only while the agent is active.
struct __align__(16) AgentRec {
    uint16_t t_exposed;      // :200,:355  (0xFFFF == -1 never)
    uint16_t tt_activation;  // :218,:356
    uint16_t t_infected;     // :201,:259
    uint16_t tt_severe;      // :213,:263  (0xFFFF == inf)
    uint16_t tt_recovery;    // :214
    uint16_t t_severe;       // :202,:305
    uint16_t tt_critical;    // :215
    uint16_t t_critical;     // :203,:315
    uint16_t tt_death;       // :216
    uint16_t tt_isolate;     // :217
    uint16_t t_documented;   // :204
    uint16_t n_inf_by;       // :207
    uint16_t n_inf_asympt;   // :209
    uint16_t n_inf_outside;  // :208
    uint16_t pad0, pad1;
};
static_assert(sizeof(AgentRec) == 32, "AgentRec must be one 32 B sector");

// static per-agent word: age(7) | diabetes<<7 | hypertension<<8 | household position(3)<<9 | (household size-1)(3)<<12
__host__ __device__ inline uint16_t pack_static(int age, int diab, int hyp, int pos, int size)
{
    return (uint16_t)(age | (diab << 7) | (hyp << 8) | (pos << 9) | ((size - 1) << 12));
}

// inbox word (16 agents): byte b holds agents b, b+4, b+8, b+12 -- "infected" in bits 0-3,
// "Q on infection" in bits 4-7 -- so that (word >> w) & 0x01010101 lines the infected bits of
// state word w (agents 4w..4w+3) up with that word's bytes.
__host__ __device__ inline uint32_t inbox_infected_bit(int k) { return 1u << (8 * (k & 3) + (k >> 2)); }
__host__ __device__ inline uint32_t inbox_q_bit(int k) { return 1u << (8 * (k & 3) + 4 + (k >> 2)); }

struct DevCtx {
    // population, shared by all replicas
    int64_t n;            // agents
    int64_t n_pad;        // row stride, multiple of kTile
    int tiles_per_rep;
    int n_rep;
    int chunks_per_rep;   // work items per replica and day (contiguous tile ranges)
    int T;
    const uint16_t *stat;
    const int32_t *grp_off;      // [102]
    const int32_t *grp_members;  // [n]
    const double *p_hh_vec;      // [n] or nullptr
    double p_hh_scalar;
    // tables
    const double *p_ms, *p_sc, *p_cd;  // [101*4]
    const double *iso_mean, *iso_asym; // [101]
    const double *nat_enlam;           // [2][101][2]
    const double *nat_cdf;             // [101][101]
    const double *enlam;               // [2][101][101] (replay) or nullptr
    // per-replica state
    uint8_t *hist;                 // [rep][T][n_pad]
    uint32_t *inbox;               // [rep][n_pad/16]
    unsigned long long *winner;    // [rep][n_pad]
    AgentRec *rec;                 // [rep][n_pad]
    uint32_t *tally;               // [rep][T][TAL_ROWS][101]
    const uint32_t *home;          // [rep][n_pad/32] bitmap of Home_real
    const uint64_t *seeds;         // [rep]
    unsigned long long *counters;  // [rep][4]
    unsigned int *barrier;         // grid barrier counter
    // scalars
    int t_lockdown, t_stayhome, mode;
    double p_doc, p_contact, asym;
    double m_sev, m_mild_rec, m_crit, m_sev_rec, m_crit_rec, m_death, act_mu, act_sigma;
};

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t add_sat16(uint16_t a, uint16_t b)
{
    uint32_t s = (uint32_t)a + (uint32_t)b;
    return (a == NEVER || b == NEVER || s >= 0xFFFFu) ? NEVER : (uint16_t)s;
}
__device__ __forceinline__ uint16_t inc_sat16(uint16_t a) { return a >= 0xFFFEu ? a : (uint16_t)(a + 1); }
// `t - t0 == tt` of the reference on integer-valued doubles; tt == NEVER encodes inf/nan
__device__ __forceinline__ bool due(int t, uint16_t t0, uint16_t tt) { return tt != NEVER && (int)t0 + (int)tt == t; }

__device__ __forceinline__ AgentRec load_rec(const AgentRec *p)
{
    union { AgentRec r; uint4 v[2]; } u;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    u.v[0] = __ldcg(q);
    u.v[1] = __ldcg(q + 1);
    return u.r;
}
__device__ __forceinline__ void store_rec(AgentRec *p, const AgentRec &r)
{
    union { AgentRec r; uint4 v[2]; } u;
    u.r = r;
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = u.v[0];
    q[1] = u.v[1];
}

struct RepView {
    const uint8_t *prev;  // row t-1
    uint8_t *cur;         // row t (nullptr in the finalize pass)
    uint8_t *prev_w;      // row t-1, writable (fix-up patch)
    uint32_t *inbox;
    unsigned long long *winner;
    AgentRec *rec;
    const uint32_t *home;
    uint64_t seed;
    bool home_on;
    int phase;
};

// S[t-1, c] of the reference, seen from another agent's thread during pass t.
__device__ __forceinline__ bool was_susceptible(const RepView &v, int64_t c, int t)
{
    uint8_t s = __ldcg(v.prev + c);
    if ((s & 7) != C_S) return false;
    uint32_t ib = __ldcg(v.inbox + (c >> 4));
    if (!(ib & inbox_infected_bit((int)(c & 15)))) return true;
    unsigned long long w = __ldcg(v.winner + c);
    return (int)(w >> 48) == t;  // hit earlier today: still S in row t-1
}
__device__ __forceinline__ bool at_home(const RepView &v, int64_t c)
{
    return v.home_on && ((__ldg(v.home + (c >> 5)) >> (c & 31)) & 1u);
}

// One successful transmission i -> c on day t (:347-359 / :376-390, infectee side).
__device__ __forceinline__ void hit(const DevCtx &cx, const RepView &v, int64_t i, int64_t c, int t, uint32_t ord)
{
    unsigned long long key = ((unsigned long long)t << 48) | ((unsigned long long)(uint32_t)i << 16) | ord;
    atomicMax(v.winner + c, key);
    int age_c = __ldg(cx.stat + c) & 127;
    double tiso = sd_exp_days(sd_lane(sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_INF, ord << 8), 0),
                              __ldg(cx.iso_asym + age_c));
    uint32_t bits = inbox_infected_bit((int)(c & 15));
    if (tiso == 0.0) bits |= inbox_q_bit((int)(c & 15));
    __threadfence();  // the winner must be visible before the bit that announces it
    atomicOr(v.inbox + (c >> 4), bits);
}

// ---------------------------------------------------------------------------------------
// One queued agent, pass t.  `st0` is its byte in row t-1 as staged by the dense phase.
__device__ void process_agent(const DevCtx &cx, const RepView &v, int64_t i, int t, uint8_t st0, uint32_t ib_word,
                              uint32_t *s_tally, unsigned long long *s_cnt)
{
    const uint16_t stt = __ldg(cx.stat + i);
    const int age = stt & 127;
    const int k3 = age * 4 + ((stt >> 7) & 1) * 2 + ((stt >> 8) & 1);
    AgentRec r;
    uint8_t st = st0;
    bool dirty = false;

    if ((st & 7) == C_S) {
        // infected by someone's push: apply it if it happened on day t-1
        unsigned long long w = __ldcg(v.winner + i);
        if ((int)(w >> 48) != t - 1) return;  // hit today by a racing sender: still S in row t-1
        const uint32_t inf = (uint32_t)(w >> 16), ord = (uint32_t)(w & 0xFFFFu);
        const double tiso = sd_exp_days(sd_lane(sd_draw(v.seed, inf, (uint32_t)(t - 1), SD_SITE_INF, ord << 8), 0),
                                        __ldg(cx.iso_asym + age));
        const double tact = sd_lognormal_days(v.seed, inf, (uint32_t)(t - 1), SD_SITE_INF, ord << 8, cx.act_mu, cx.act_sigma);
        r = AgentRec{};
        r.t_exposed = (uint16_t)(t - 1);
        r.tt_isolate = sd_enc16(tiso);
        r.tt_activation = sd_enc16(tact);
        st = C_E | ((ib_word & inbox_q_bit((int)(i & 15))) ? ST_Q : 0);
        v.prev_w[i] = st;  // patch row t-1 (:347,:351,:354)
        atomicAdd(&s_tally[TAL_NEW_E * kAges + age], 1u);
        atomicAdd(&s_cnt[1], 1ull);
        dirty = true;
    } else {
        r = load_rec(v.rec + i);
    }

    const int comp = st & 7;
    // state tallies of day t-1
    atomicAdd(&s_tally[(comp - C_E) * kAges + age], 1u);
    if (st & ST_Q) atomicAdd(&s_tally[TAL_QACT * kAges + age], 1u);
    atomicAdd(&s_cnt[0], 1ull);

    if (v.cur == nullptr) {  // finalize pass (t == T): nothing moves
        if (dirty) store_rec(v.rec + i, r);
        return;
    }

    const bool Ep = comp == C_E, Mp = comp == C_MILD, Sevp = comp == C_SEV, Cp = comp == C_CRIT;
    const bool Qp = (st & ST_Q) != 0, Dp = (st & ST_DOC) != 0;
    int ncomp = comp;
    bool nq = Qp, ndoc = Dp;
    bool transmit = true;

    if (Ep && due(t, r.t_exposed, r.tt_activation)) {  // :256-272
        ncomp = C_MILD;
        r.t_infected = (uint16_t)t;
        sd_block b = sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_ACT, 0);
        if (sd_lane(b, 0) < __ldg(cx.p_ms + k3)) {
            r.tt_severe = sd_enc16(sd_exp_days(sd_lane(b, 1), cx.m_sev));
            r.tt_recovery = NEVER;
        } else {
            r.tt_recovery = sd_enc16(sd_exp_days(sd_lane(b, 1), cx.m_mild_rec));
            r.tt_severe = NEVER;
        }
        double tiso = sd_exp_days(sd_lane(sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_ACT, 1), 0),
                                  __ldg(cx.iso_mean + age));
        r.tt_isolate = sd_enc16(tiso);
        if (tiso == 0.0) nq = true;
        dirty = true;
    }
    if (Mp || Sevp || Cp) {  // :274-327
        if (due(t, r.t_infected, r.tt_recovery)) {  // :276-279
            ncomp = C_R;
            nq = false;
            transmit = false;
            atomicAdd(&s_tally[TAL_NEW_R * kAges + age], 1u);
            goto write_back;
        }
        if (Mp && !Dp) {  // :280-289
            if (sd_lane(sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_DOC, 0), 0) < cx.p_doc) {
                ndoc = true;
                r.t_documented = (uint16_t)t;
                dirty = true;
            }
        }
        if (Mp && due(t, r.t_infected, r.tt_severe)) {  // :291-311
            ncomp = C_SEV;
            if (!Dp) {
                ndoc = true;
                r.t_documented = (uint16_t)t;
            }
            nq = true;
            r.t_severe = (uint16_t)t;
            sd_block b = sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_SEV, 0);
            if (sd_lane(b, 0) < __ldg(cx.p_sc + k3)) {
                r.tt_critical = sd_enc16(sd_exp_days(sd_lane(b, 1), cx.m_crit));
                r.tt_recovery = NEVER;
            } else {
                r.tt_recovery = add_sat16(sd_enc16(sd_exp_days(sd_lane(b, 1), cx.m_sev_rec)), r.tt_severe);
                r.tt_critical = NEVER;
            }
            dirty = true;
        } else if (Sevp && due(t, r.t_severe, r.tt_critical)) {  // :312-321
            ncomp = C_CRIT;
            r.t_critical = (uint16_t)t;
            sd_block b = sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_CRIT, 0);
            if (sd_lane(b, 0) < __ldg(cx.p_cd + k3)) {
                r.tt_death = sd_enc16(sd_exp_days(sd_lane(b, 1), cx.m_death));
                r.tt_recovery = NEVER;
            } else {
                r.tt_recovery = add_sat16(add_sat16(sd_enc16(sd_exp_days(sd_lane(b, 1), cx.m_crit_rec)), r.tt_severe),
                                          r.tt_critical);
                r.tt_death = NEVER;
            }
            dirty = true;
        } else if (Cp) {  // :323-327
            if (due(t, r.t_critical, r.tt_death)) {
                ncomp = C_D;
                nq = false;
                atomicAdd(&s_tally[TAL_NEW_D * kAges + age], 1u);
            }
        }
    }
    if (ndoc && !Dp) atomicAdd(&s_tally[TAL_NEW_DOC * kAges + age], 1u);

    // :328-337
    if (Qp) transmit = false;
    if (transmit) {
        if (!Ep && due(t, r.t_infected, r.tt_isolate)) { nq = true; transmit = false; }
        else if (Ep && due(t, r.t_exposed, r.tt_isolate)) { nq = true; transmit = false; }
    }
    if (transmit) {
        uint32_t hits = 0, hits_out = 0;
        // ---- household push :339-359.  Households are contiguous ranges [i-pos, i-pos+size).
        {
            const int pos = (stt >> 9) & 7, size = ((stt >> 12) & 7) + 1;
            double p_h = cx.p_hh_vec ? __ldg(cx.p_hh_vec + i) : cx.p_hh_scalar;
            if (Ep) p_h = SD_MUL(p_h, cx.asym);
            const int64_t base = i - pos;
            for (int j = 0; j < size - 1; ++j) {
                const int64_t c = base + j + (j >= pos ? 1 : 0);
                if (!was_susceptible(v, c, t)) continue;
                if (sd_lane(sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_HH, (uint32_t)j), 0) < p_h) {
                    hit(cx, v, i, c, t, SD_ORD_HH(j));
                    ++hits;
                }
            }
        }
        // ---- community push :361-390
        if (!at_home(v, i)) {
            if (cx.mode == SEIR_MODE_NATIVE) {
                const double en = __ldg(cx.nat_enlam + ((v.phase * kAges + age) * 2 + (Ep ? 1 : 0)));
                const uint32_t kept = sd_poisson(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_NCNT, 0, en, SD_MAX_CONTACTS_NATIVE);
                const double *cdf = cx.nat_cdf + age * kAges;
                for (uint32_t k = 0; k < kept; ++k) {
                    sd_block b = sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_NPICK, k);
                    const double u = sd_lane(b, 0);
                    int lo = 0, hi = kAges - 1;
                    while (lo < hi) {
                        int mid = (lo + hi) >> 1;
                        if (u < __ldg(cdf + mid)) hi = mid; else lo = mid + 1;
                    }
                    const int g0 = __ldg(cx.grp_off + lo), glen = __ldg(cx.grp_off + lo + 1) - g0;
                    if (glen == 0) continue;
                    const int64_t c = __ldg(cx.grp_members + g0 + (int64_t)sd_index(b, (uint64_t)glen));
                    if (was_susceptible(v, c, t) && !at_home(v, c)) {
                        hit(cx, v, i, c, t, SD_ORD_COMM_NATIVE(k));
                        ++hits;
                        ++hits_out;
                    }
                }
            } else {
                double p_c = cx.p_contact;
                if (Ep) p_c = SD_MUL(p_c, cx.asym);
                const double *en = cx.enlam + ((size_t)v.phase * kAges + age) * kAges;
                for (int a = 0; a < kAges; ++a) {
                    const int g0 = __ldg(cx.grp_off + a), glen = __ldg(cx.grp_off + a + 1) - g0;
                    if (glen == 0) continue;  // :368-369
                    const uint32_t nc = sd_poisson(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_POIS, (uint32_t)a << 8,
                                                   __ldg(en + a), SD_MAX_CONTACTS_PER_AGE);
                    for (uint32_t j = 0; j < nc; ++j) {
                        sd_block b = sd_draw(v.seed, (uint32_t)i, (uint32_t)t, SD_SITE_KEEP, ((uint32_t)a << 8) | j);
                        if (!(sd_lane(b, 0) < p_c)) continue;  // :373
                        const int64_t c = __ldg(cx.grp_members + g0 + (int64_t)sd_index(b, (uint64_t)glen));  // :374
                        if (was_susceptible(v, c, t) && !at_home(v, c)) {  // :375
                            hit(cx, v, i, c, t, SD_ORD_COMM_REPLAY(a, j));
                            ++hits;
                            ++hits_out;
                        }
                    }
                }
            }
        }
        if (hits) {
            for (uint32_t h = 0; h < hits; ++h) {
                r.n_inf_by = inc_sat16(r.n_inf_by);
                if (Ep) r.n_inf_asympt = inc_sat16(r.n_inf_asympt);
            }
            for (uint32_t h = 0; h < hits_out; ++h) r.n_inf_outside = inc_sat16(r.n_inf_outside);
            atomicAdd(&s_cnt[2], (unsigned long long)hits);
            dirty = true;
        }
    }

write_back:
    {
        const uint8_t nst = (uint8_t)(ncomp | (nq ? ST_Q : 0) | (ndoc ? ST_DOC : 0));
        if (nst != st0) v.cur[i] = nst;  // the dense phase already carried st0 into row t
        if (dirty) store_rec(v.rec + i, r);
    }
}

// ---------------------------------------------------------------------------------------
struct __align__(16) TileSmem {
    uint8_t prev[kTile];      // staged row t-1 of this tile
    uint32_t ib[kThreads];    // staged inbox words
    uint16_t queue[kTile];    // local indices of agents that need work
    uint32_t tally[TAL_ROWS * kAges];
    unsigned long long cnt[4];
    uint32_t q_count;
};

// One work item: tiles [tile_lo, tile_hi) of replica `rep`, pass t.
__device__ void run_item(const DevCtx &cx, TileSmem &sm, int rep, int tile_lo, int tile_hi, int t)
{
    const int tid = threadIdx.x;
    RepView v;
    uint8_t *hist = cx.hist + (size_t)rep * cx.T * cx.n_pad;
    v.prev = hist + (size_t)(t - 1) * cx.n_pad;
    v.prev_w = hist + (size_t)(t - 1) * cx.n_pad;
    v.cur = (t < cx.T) ? hist + (size_t)t * cx.n_pad : nullptr;
    v.inbox = cx.inbox + (size_t)rep * (cx.n_pad / 16);
    v.winner = cx.winner + (size_t)rep * cx.n_pad;
    v.rec = cx.rec + (size_t)rep * cx.n_pad;
    v.home = cx.home + (size_t)rep * (cx.n_pad / 32);
    v.seed = cx.seeds[rep];
    v.home_on = cx.t_stayhome >= 1 && t >= cx.t_stayhome;   // :243-244
    v.phase = (cx.t_lockdown >= 1 && t >= cx.t_lockdown) ? 1 : 0;  // :234-240

    for (int k = tid; k < TAL_ROWS * kAges; k += kThreads) sm.tally[k] = 0;
    if (tid < 4) sm.cnt[tid] = 0;
    if (tid == 0) sm.q_count = 0;
    __syncthreads();

    for (int tile = tile_lo; tile < tile_hi; ++tile) {
        const int64_t base = (int64_t)tile * kTile;
        const int64_t vi = base / kVec + tid;
        // ---- dense phase: 16 agents per thread
        const uint4 sv = __ldcg(reinterpret_cast<const uint4 *>(v.prev) + vi);
        const uint32_t ib = __ldcg(v.inbox + vi);
        if (v.cur) reinterpret_cast<uint4 *>(v.cur)[vi] = sv;
        reinterpret_cast<uint4 *>(sm.prev)[tid] = sv;
        sm.ib[tid] = ib;
        const uint32_t wds[4] = {sv.x, sv.y, sv.z, sv.w};
        uint32_t work[4];
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t c = wds[w] & 0x07070707u;
            const uint32_t active = ((c + 0x03030303u) & 0x04040404u) >> 2;               // bit0 of byte: comp in 1..4
            const uint32_t is_s = (~(c + 0x07070707u) & 0x08080808u) >> 3;                // bit0 of byte: comp == 0
            const uint32_t infected = (ib >> w) & 0x01010101u;                            // bit0 of byte: inbox bit
            work[w] = active | (is_s & infected);
            any |= work[w];
        }
        if (any) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t m = work[w];
                while (m) {
                    const int b = (__ffs(m) - 1) >> 3;
                    m &= m - 1;
                    const uint32_t slot = atomicAdd(&sm.q_count, 1u);
                    sm.queue[slot] = (uint16_t)(tid * kVec + w * 4 + b);
                }
            }
        }
        __syncthreads();
        // ---- agent phase: one queued agent per thread
        const uint32_t qn = sm.q_count;
        for (uint32_t e = tid; e < qn; e += kThreads) {
            const int li = sm.queue[e];
            process_agent(cx, v, base + li, t, sm.prev[li], sm.ib[li >> 4], sm.tally, sm.cnt);
        }
        __syncthreads();
        if (tid == 0) sm.q_count = 0;
        // (the next iteration's first __syncthreads orders this reset before the next agent phase;
        //  queue pushes of the next dense phase need it: fence below)
        __syncthreads();
    }

    // ---- flush the per-item tallies: rows [0,kTalLagRows) belong to day t-1, the rest to day t
    uint32_t *tal = cx.tally + (size_t)rep * cx.T * TAL_ROWS * kAges;
    for (int k = tid; k < TAL_ROWS * kAges; k += kThreads) {
        const uint32_t val = sm.tally[k];
        if (val) {
            const int row = k / kAges;
            const int day = row < kTalLagRows ? t - 1 : t;
            if (day < cx.T) atomicAdd(&tal[(size_t)day * TAL_ROWS * kAges + k], val);
        }
    }
    if (tid < 3 && sm.cnt[tid]) atomicAdd(&cx.counters[rep * 4 + 1 + tid], sm.cnt[tid]);
    __syncthreads();
}

__device__ __forceinline__ void day_pass(const DevCtx &cx, TileSmem &sm, int t)
{
    const int items = cx.n_rep * cx.chunks_per_rep;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int rep = it / cx.chunks_per_rep, chunk = it % cx.chunks_per_rep;
        const int lo = (int)((int64_t)cx.tiles_per_rep * chunk / cx.chunks_per_rep);
        const int hi = (int)((int64_t)cx.tiles_per_rep * (chunk + 1) / cx.chunks_per_rep);
        run_item(cx, sm, rep, lo, hi, t);
    }
}

// grid-wide barrier on a monotonically increasing counter (cooperative launch guarantees residency)
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

}  // namespace seir
