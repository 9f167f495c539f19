/*
 * seir_draws.h -- the keyed-draw contract of the B200 SEIR day-step.
 *
 * The reference (`seir_individual.py:163`) consumes ONE global Mersenne-Twister
 * stream in data-dependent agent order, which cannot be reproduced by a parallel
 * machine.  The CUDA path therefore re-keys every draw site of the day loop
 * (SURVEY.md section 8a "Draw-site catalogue", D1..D15, T1, T2) on
 *        Philox4x32-10( key = seed , counter = (agent, day, site, sub) )
 * so that a draw depends only on WHO draws, WHEN and WHERE in the statement
 * order -- never on how many draws other agents made before.
 *
 * This header is compiled by BOTH nvcc (device code of libseir_b200.so) and gcc
 * (the CPU oracle under oracle/, provider "keyed").  Bit-exact agreement between
 * the two needs every floating-point step to be an IEEE-754 correctly rounded
 * operation issued in the same order on both sides, so:
 *   - no libm / CUDA-libdevice log, exp (their results differ by an ulp);
 *     sd_log / sd_exp below are built from +, *, /, fma and bit moves only;
 *   - every multiply/add goes through SD_MUL/SD_ADD/SD_FMA so that neither
 *     compiler may contract or re-associate (device: __dmul_rn/__dadd_rn/__fma_rn
 *     are never fused; host: build with -ffp-contract=off).
 *
 * Draw sites (counter word 2).  `lane` 0 = words 0,1 of the Philox block,
 * lane 1 = words 2,3; a uniform double is built from two words exactly as the
 * reference's generator builds one (numba/cpython/randomimpl.py:134-147):
 *        u = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 .
 */
#ifndef SEIR_DRAWS_H
#define SEIR_DRAWS_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SD_FN static __host__ __device__ __forceinline__
#else
#define SD_FN static inline
#endif

#if defined(__CUDA_ARCH__)
#define SD_MUL(a, b) __dmul_rn((a), (b))
#define SD_ADD(a, b) __dadd_rn((a), (b))
#define SD_SUB(a, b) __dsub_rn((a), (b))
#define SD_FMA(a, b, c) __fma_rn((a), (b), (c))
#define SD_DIV(a, b) __ddiv_rn((a), (b))
#define SD_SQRT(a) __dsqrt_rn((a))
#define SD_RINT(a) rint((a))
#define SD_MULHI64(a, b) __umul64hi((a), (b))
#define SD_MULHI32(a, b) __umulhi((a), (b))
#else
#include <math.h>
#define SD_MUL(a, b) ((a) * (b))
#define SD_ADD(a, b) ((a) + (b))
#define SD_SUB(a, b) ((a) - (b))
#define SD_FMA(a, b, c) fma((a), (b), (c))
#define SD_DIV(a, b) ((a) / (b))
#define SD_SQRT(a) sqrt((a))
#define SD_RINT(a) rint((a))
#define SD_MULHI64(a, b) ((uint64_t)(((unsigned __int128)(a) * (unsigned __int128)(b)) >> 64))
#define SD_MULHI32(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#endif

/* ---- draw sites ------------------------------------------------------- */
enum {
    SD_SITE_ACT = 1,    /* E->Mild   (seir_individual.py:262-270) sub0: lane0=D1 lane1=D2; sub1: lane0=D3 */
    SD_SITE_DAY = 2,    /* the agent's day block, sub0: lane0 = D4 Mild documentation (:282),
                           lane1 = first Knuth uniform of the native kept-contact count              */
    SD_SITE_SEV = 3,    /* Mild->Severe (:306-310)                lane0=D5 lane1=D6                     */
    SD_SITE_CRIT = 4,   /* Severe->Critical (:316-320)            lane0=D7 lane1=D8                     */
    SD_SITE_HH = 5,     /* household push (:346)  slot j: sub = j>>1, lane = j&1   (D9)                 */
    SD_SITE_INF = 6,    /* infectee timers (:352,356,381,385) agent=INFECTOR, sub=(ord<<8)|call:
                           call0 lane0 = time_to_isolate exponential, calls 1.. = polar normal tries   */
    SD_SITE_POIS = 7,   /* replay community Poisson (:370) sub=(contact_age<<8)|call, Knuth products    */
    SD_SITE_KEEP = 8,   /* replay contact (:373,374) sub=(contact_age<<8)|j lane0=D13, words2,3=D14     */
    SD_SITE_NCNT = 9,   /* native community: kept-contact count, Knuth uniforms 2,3,.. (sub=call, lanes 0,1);
                           uniform 1 is lane1 of the day block                                         */
    SD_SITE_NPICK = 10, /* native community: sub=k, lane0 -> contact age (alias table), words2,3 -> member index */
    SD_SITE_TR_HH = 11, /* tracing, household edge (:72)  agent=tracer sub=slot j lane0                 */
    SD_SITE_TR_OUT = 12,/* tracing, outside edge (:83)    agent=tracer sub=log slot j lane0             */
    SD_SITE_INIT = 13,  /* prologue (:227) activation time of an initially infected agent, polar calls  */
    SD_SITE_HOME = 14,  /* keyed prologue: Home_real of one age group (:178-182), agent = age, round keys of sd_perm */
    SD_SITE_CASES = 15  /* keyed prologue: initial_infected (:187), agent = 0, round keys of sd_perm             */
};

/* event ordinal of an infection inside its infector's day, monotone in the
 * reference's statement order (household slots first, then community) */
#define SD_ORD_HH(j) ((uint32_t)(j))
#define SD_ORD_COMM_REPLAY(age, j) (16u + ((uint32_t)(age) << 8) + (uint32_t)(j))
#define SD_ORD_COMM_NATIVE(k) (16u + (uint32_t)(k))
#define SD_MAX_CONTACTS_PER_AGE 255u   /* replay: Poisson count per (sender, age) is clamped here */
#define SD_MAX_CONTACTS_NATIVE 65000u  /* native: kept contacts per sender-day are clamped here    */

/* ---- Philox4x32-10 (Salmon et al., SC'11) ------------------------------ */
typedef struct { uint32_t w[4]; } sd_block;

SD_FN sd_block sd_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = SD_MULHI32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = SD_MULHI32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    sd_block b;
    b.w[0] = c0; b.w[1] = c1; b.w[2] = c2; b.w[3] = c3;
    return b;
}

/* keyed block for (agent, day, site, sub) under a 64-bit seed */
SD_FN sd_block sd_draw(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub)
{
    return sd_philox((uint32_t)seed, (uint32_t)(seed >> 32), agent, day, site, sub);
}

/* 53-bit uniform in [0,1) from two words (randomimpl.py:134-147) */
SD_FN double sd_u53(uint32_t a, uint32_t b)
{
    /* (a>>5)*2^26 + (b>>6) < 2^53 is exact as an integer; the scale is a power of two */
    uint64_t m = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return (double)m * (1.0 / 9007199254740992.0);
}
SD_FN double sd_lane(sd_block b, int lane) { return lane ? sd_u53(b.w[2], b.w[3]) : sd_u53(b.w[0], b.w[1]); }
/* uniform index in [0,n) from words 2,3: floor(r64 * n / 2^64) */
SD_FN uint64_t sd_index(sd_block b, uint64_t n)
{
    uint64_t r = ((uint64_t)b.w[2] << 32) | (uint64_t)b.w[3];
    return SD_MULHI64(r, n);
}

/* ---- keyed permutation of [0, n): the prologue's `choice(..., replace=False)` (:181, :187) -----------------------
 * The reference draws a k-subset as the first k entries of a random permutation.  Here the permutation is a keyed
 * bijection: a 6-round Feistel network over the smallest even number of bits that holds n, with cycle walking (apply
 * again while the image is >= n), its round keys one pair of Philox blocks per (seed, site, group).  The k-subset is
 * { sd_perm_at(P, r) : r < k }: exactly k distinct members, any of them computable on its own -- O(k) work, no scan
 * over the n candidates, no atomics on a counter, the same on the host and on the device.                         */
typedef struct { uint32_t k[6]; uint32_t half, mask, n; } sd_perm;
SD_FN uint32_t sd_fmix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
SD_FN sd_perm sd_perm_make(uint64_t seed, uint32_t site, uint32_t group, uint32_t n)
{
    sd_perm P;
    const sd_block a = sd_draw(seed, group, 0u, site, 0u), b = sd_draw(seed, group, 0u, site, 1u);
    P.k[0] = a.w[0]; P.k[1] = a.w[1]; P.k[2] = a.w[2]; P.k[3] = a.w[3]; P.k[4] = b.w[0]; P.k[5] = b.w[1];
    uint32_t bits = 0;
    while (bits < 32u && ((uint64_t)1 << bits) < (uint64_t)n) ++bits;   /* 2^bits >= n */
    P.half = bits < 2u ? 1u : (bits + 1u) / 2u;
    P.mask = (1u << P.half) - 1u;
    P.n = n;
    return P;
}
SD_FN uint32_t sd_perm_at(const sd_perm *P, uint32_t rank)            /* rank < n */
{
    uint32_t x = rank;
    do {
        uint32_t L = x >> P->half, R = x & P->mask;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 6; ++r) {
            const uint32_t f = sd_fmix32(R ^ P->k[r]) & P->mask;
            const uint32_t nl = R;
            R = L ^ f;
            L = nl;
        }
        x = (L << P->half) | R;
    } while (x >= P->n);
    return x;
}

/* ---- bit moves -------------------------------------------------------- */
SD_FN uint64_t sd_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    union { double d; uint64_t u; } v; v.d = x; return v.u;
#endif
}
SD_FN double sd_from_bits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { double d; uint64_t u; } v; v.u = u; return v.d;
#endif
}

/* ---- deterministic natural log, x finite > 0 (normal or subnormal) ------
 * x = 2^k * m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s=(m-1)/(m+1),
 * |s| <= 0.1716: odd series to s^23 (truncation < 1e-18 relative).         */
SD_FN double sd_log(double x)
{
    if (!(x > 0.0)) {
        if (x == 0.0) return sd_from_bits(0xFFF0000000000000ull);  /* -inf */
        return sd_from_bits(0x7FF8000000000000ull);                /* nan  */
    }
    uint64_t ux = sd_bits(x);
    if (ux >= 0x7FF0000000000000ull) return x;                       /* +inf */
    int k = 0;
    if (ux < 0x0010000000000000ull) {                                /* subnormal: scale by 2^54 */
        x = SD_MUL(x, 18014398509481984.0);
        ux = sd_bits(x);
        k = -54;
    }
    k += (int)(ux >> 52) - 1023;
    uint64_t mant = ux & 0x000FFFFFFFFFFFFFull;
    /* m in [1,2); fold the upper part into [sqrt(1/2),1) */
    if (mant >= 0x6A09E667F3BCDull) { k += 1; ux = mant | 0x3FE0000000000000ull; }
    else ux = mant | 0x3FF0000000000000ull;
    double m = sd_from_bits(ux);
    double f = SD_SUB(m, 1.0);
    double s = SD_DIV(f, SD_ADD(m, 1.0));
    double z = SD_MUL(s, s);
    double p = 1.0 / 23.0;
    p = SD_FMA(p, z, 1.0 / 21.0);
    p = SD_FMA(p, z, 1.0 / 19.0);
    p = SD_FMA(p, z, 1.0 / 17.0);
    p = SD_FMA(p, z, 1.0 / 15.0);
    p = SD_FMA(p, z, 1.0 / 13.0);
    p = SD_FMA(p, z, 1.0 / 11.0);
    p = SD_FMA(p, z, 1.0 / 9.0);
    p = SD_FMA(p, z, 1.0 / 7.0);
    p = SD_FMA(p, z, 1.0 / 5.0);
    p = SD_FMA(p, z, 1.0 / 3.0);
    /* log m = 2s + 2s*z*p */
    double two_s = SD_ADD(s, s);
    double lm = SD_FMA(SD_MUL(two_s, z), p, two_s);
    double dk = (double)k;
    /* k*ln2 split hi/lo (hi has 32 trailing zero bits: k*hi exact for |k|<2^20) */
    double r = SD_FMA(dk, 1.90821492927058770002e-10, lm);
    return SD_FMA(dk, 6.93147180369123816490e-01, r);
}

/* ---- deterministic exp ------------------------------------------------- */
SD_FN double sd_exp(double x)
{
    if (x != x) return x;
    if (x > 709.78) return sd_from_bits(0x7FF0000000000000ull);
    if (x < -745.2) return 0.0;
    double dk = SD_RINT(SD_MUL(x, 1.44269504088896338700e+00));
    double r = SD_FMA(dk, -6.93147180369123816490e-01, x);
    r = SD_FMA(dk, -1.90821492927058770002e-10, r);
    /* |r| <= 0.3466 : Taylor to r^13 */
    double p = 1.0 / 6227020800.0;
    p = SD_FMA(p, r, 1.0 / 479001600.0);
    p = SD_FMA(p, r, 1.0 / 39916800.0);
    p = SD_FMA(p, r, 1.0 / 3628800.0);
    p = SD_FMA(p, r, 1.0 / 362880.0);
    p = SD_FMA(p, r, 1.0 / 40320.0);
    p = SD_FMA(p, r, 1.0 / 5040.0);
    p = SD_FMA(p, r, 1.0 / 720.0);
    p = SD_FMA(p, r, 1.0 / 120.0);
    p = SD_FMA(p, r, 1.0 / 24.0);
    p = SD_FMA(p, r, 1.0 / 6.0);
    p = SD_FMA(p, r, 0.5);
    p = SD_FMA(p, r, 1.0);
    p = SD_FMA(p, r, 1.0);
    int k = (int)dk;
    /* scale by 2^k in two exact-or-single-rounding steps */
    int k1 = k / 2, k2 = k - k1;
    double s1 = sd_from_bits((uint64_t)(k1 + 1023) << 52);
    double s2 = sd_from_bits((uint64_t)(k2 + 1023) << 52);
    return SD_MUL(SD_MUL(p, s1), s2);
}

/* ---- variates ---------------------------------------------------------- */
/* functions.py:16-17  round(Exp(mean)) = rint(-log(1-U)*mean)  (randomimpl.py:966-972;
 * np.round lowers to llvm.rint, numba/np/arraymath.py:3172).  May be +inf or nan. */
SD_FN double sd_exp_days(double u, double mean)
{
    return SD_RINT(SD_MUL(-sd_log(SD_SUB(1.0, u)), mean));
}

/* one try of the polar method (randomimpl.py:358-373): returns 1 and *z on accept.
 * The reference keeps f*x1 cached and returns f*x2 first; the keyed contract uses f*x2. */
SD_FN int sd_polar_try(sd_block b, double *z)
{
    double x1 = SD_SUB(SD_MUL(2.0, sd_lane(b, 0)), 1.0);
    double x2 = SD_SUB(SD_MUL(2.0, sd_lane(b, 1)), 1.0);
    double r2 = SD_ADD(SD_MUL(x1, x1), SD_MUL(x2, x2));
    if (r2 < 1.0 && r2 != 0.0) {
        double f = SD_SQRT(SD_DIV(SD_MUL(-2.0, sd_log(r2)), r2));
        *z = SD_MUL(f, x2);
        return 1;
    }
    return 0;
}

/* functions.py:19-25  round(LogNormal(mu,sigma)); tries are Philox calls
 * (agent, day, site, sub_base + 1 + try) */
SD_FN double sd_lognormal_days(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub_base,
                               double mu, double sigma)
{
    double z = 0.0;
    for (uint32_t call = 1; call < 256u; ++call) {
        if (sd_polar_try(sd_draw(seed, agent, day, site, sub_base + call), &z)) break;
    }
    double x = sd_exp(SD_FMA(sigma, z, mu));
    if (x <= 0.0) return 1.0;
    return SD_RINT(x);
}

/* Knuth product Poisson (randomimpl.py:1672-1691) against a precomputed enlam=exp(-lam);
 * uniforms come from calls (agent, day, site, sub_base + call), lanes 0 then 1. */
SD_FN uint32_t sd_poisson(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub_base,
                          double enlam, uint32_t cap)
{
    uint32_t X = 0;
    double prod = 1.0;
    for (uint32_t call = 0;; ++call) {
        sd_block b = sd_draw(seed, agent, day, site, sub_base + call);
        prod = SD_MUL(prod, sd_lane(b, 0));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
        prod = SD_MUL(prod, sd_lane(b, 1));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
    }
}

/* native kept-contact count: Knuth product whose FIRST uniform is lane 1 of the agent's day block
 * `day` (so that most agent-days need a single Philox block), the rest come from SD_SITE_NCNT */
SD_FN uint32_t sd_poisson_day(uint64_t seed, uint32_t agent, uint32_t day, sd_block day_block, double enlam, uint32_t cap)
{
    double prod = sd_lane(day_block, 1);
    if (prod <= enlam) return 0;
    uint32_t X = 1;
    for (uint32_t call = 0;; ++call) {
        sd_block b = sd_draw(seed, agent, day, SD_SITE_NCNT, call);
        prod = SD_MUL(prod, sd_lane(b, 0));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
        prod = SD_MUL(prod, sd_lane(b, 1));
        if (prod <= enlam || X >= cap) return X;
        X += 1;
    }
}

/* contact age of a kept contact: Walker/Vose alias table of the sender's contact-matrix row
 * (n_ages columns: prob[b] in [0,1], alias[b]); one table look-up instead of a CDF search */
SD_FN int sd_alias_pick(double u, const double *prob_row, const int32_t *alias_row, int n_ages)
{
    double x = SD_MUL(u, (double)n_ages);
    int b = (int)x;
    if (b > n_ages - 1) b = n_ages - 1;
    double frac = SD_SUB(x, (double)b);
    return (frac < prob_row[b]) ? b : (int)alias_row[b];
}

/* saturating day encoding used by the device records: nan/inf/>=65535 -> 0xFFFF ("never") */
SD_FN uint16_t sd_enc16(double days)
{
    if (!(days < 65535.0)) return (uint16_t)0xFFFFu;
    if (days < 0.0) return 0;
    return (uint16_t)days;
}

#endif /* SEIR_DRAWS_H */
