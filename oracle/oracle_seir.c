/*
 * oracle_seir.c -- CPU restatement of the reference's daily agent-stepping path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libseir_b200.so and the Python host code above it) never does.
 *
 * Follows /root/reference/seir_individual.py:90-392 (`run_model`), :60-88
 * (`do_contact_tracing`), :46-51 (`get_isolation_factor`) and
 * /root/reference/functions.py:16-25 statement for statement, including the
 * `continue`s at :279,:334,:337, the overwrite of time_to_isolate at :270 before
 * the E-branch test at :335, the draw-is-zero freeze, the same-day multiple-hit
 * double counting and the last-writer timers.
 *
 * Parity is pinned by running the UNMODIFIED reference in the build container
 * (oracle/ref_shim.py) -- the reference holds no tests or golden vectors of its
 * own -- and by fixtures frozen from those runs under tests/golden/
 * (oracle/make_golden.py).  With provider ORACLE_MT the draws come from an
 * MT19937 stream consumed in the reference's order through restatements of
 * numba 0.65.0's variate algorithms (numba/cpython/randomimpl.py, cited below;
 * numba is a third-party dependency the reference does not pin), and all 20
 * outputs are bit-identical to the reference.
 *
 * Providers ORACLE_KEYED_REPLAY / ORACLE_KEYED_NATIVE keep the same control
 * flow but take every draw from the keyed Philox contract in
 * include/seir_draws.h; those are what the CUDA kernels are compared with.
 * NATIVE replaces the per-sender loop over 101 Poisson draws (:367-374) by the
 * distribution-identical "Poisson number of kept contacts, then contact age
 * from the row's alias table" (Poisson superposition + thinning).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/seir_draws.h"

#define ORACLE_MT 0
#define ORACLE_KEYED_REPLAY 1
#define ORACLE_KEYED_NATIVE 2

/* ------------------------------------------------------------------------
 * MT19937 as numba drives it (numba/_random.c:40-75) + variates
 * ---------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397
typedef struct {
    uint32_t mt[MT_N];
    int index;
    int has_gauss;
    double gauss;
} mt_state;

static void mt_init(mt_state *s, uint32_t seed) /* numba_rnd_init, _random.c:59-75 */
{
    for (int pos = 0; pos < MT_N; pos++) {
        s->mt[pos] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
    }
    s->index = MT_N;
    s->has_gauss = 0;
    s->gauss = 0.0;
}

static void mt_shuffle(mt_state *s) /* numba_rnd_shuffle, _random.c:37-56 */
{
    const uint32_t UP = 0x80000000u, LOW = 0x7fffffffu, A = 0x9908b0dfu;
    int i;
    uint32_t y;
    for (i = 0; i < MT_N - MT_M; i++) {
        y = (s->mt[i] & UP) | (s->mt[i + 1] & LOW);
        s->mt[i] = s->mt[i + MT_M] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    for (; i < MT_N - 1; i++) {
        y = (s->mt[i] & UP) | (s->mt[i + 1] & LOW);
        s->mt[i] = s->mt[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    y = (s->mt[MT_N - 1] & UP) | (s->mt[0] & LOW);
    s->mt[MT_N - 1] = s->mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
}

static inline uint32_t mt_next32(mt_state *s) /* randomimpl.py:109-131 */
{
    if (s->index >= MT_N) {
        mt_shuffle(s);
        s->index = 0;
    }
    uint32_t y = s->mt[s->index++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static inline double mt_double(mt_state *s) /* randomimpl.py:134-147 */
{
    double a = (double)(mt_next32(s) >> 5);
    double b = (double)(mt_next32(s) >> 6);
    return (b + a * 67108864.0) / 9007199254740992.0;
}

/* np.random.randint(0, n), 64-bit operands (randomimpl.py:454-521, 150-195) */
static inline int64_t mt_randint(mt_state *s, int64_t n)
{
    if (n == 1) return 0;
    int nbits = 64 - __builtin_clzll((uint64_t)(n - 1));
    for (;;) {
        int64_t r;
        if (nbits <= 32) {
            uint32_t mask = 0xffffffffu >> (32 - nbits);
            r = (int64_t)(mt_next32(s) & mask);
        } else {
            uint32_t mask = 0xffffffffu >> (64 - nbits);
            uint64_t high = mt_next32(s) & mask;
            uint64_t low = mt_next32(s);
            r = (int64_t)(low + (high << 32));
        }
        if (r < n) return r;
    }
}

/* np.random.shuffle on a 1-d buffer (randomimpl.py:1927-1946) */
static void mt_shuffle_i64(mt_state *s, int64_t *x, int64_t len)
{
    int64_t i = len - 1;
    while (i > 0) {
        int64_t j = mt_randint(s, i + 1);
        int64_t tmp = x[i];
        x[i] = x[j];
        x[j] = tmp;
        i -= 1;
    }
}

static double mt_gauss(mt_state *s) /* randomimpl.py:358-405 */
{
    if (s->has_gauss) {
        s->has_gauss = 0;
        return s->gauss;
    }
    double x1, x2, r2;
    for (;;) {
        x1 = 2.0 * mt_double(s) - 1.0;
        x2 = 2.0 * mt_double(s) - 1.0;
        r2 = x1 * x1 + x2 * x2;
        if (r2 < 1.0 && r2 != 0.0) break;
    }
    double f = sqrt(-2.0 * log(r2) / r2);
    s->gauss = f * x1;
    s->has_gauss = 1;
    return f * x2;
}

static int64_t mt_poisson(mt_state *s, double lam) /* randomimpl.py:1672-1691 (lam < 10) */
{
    if (lam == 0.0) return 0;
    double enlam = exp(-lam);
    int64_t X = 0;
    double prod = 1.0;
    for (;;) {
        double U = mt_double(s);
        prod *= U;
        if (prod <= enlam) return X;
        X += 1;
    }
}

/* ------------------------------------------------------------------------
 * Interface
 * ---------------------------------------------------------------------- */
typedef struct {
    int64_t n, T, n_ages, W;
    int64_t t_lockdown, t_tracing_start, t_stayhome_start;
    int64_t tracing_enabled; /* value of `tracing_enabled` at :241.  The reference, as compiled by
                                numba 0.65, behaves as 1 whatever params['contact_tracing'] says
                                (SURVEY.md section 0 item 1). */
    int64_t provider;
    int64_t n_iso_rows;      /* rows of mean_time_to_isolate_factor */
    int64_t n_initial;       /* keyed providers: length of initial_infected[] */
    uint64_t seed;
    double lockdown_factor, p_trace_household, p_trace_outside, p_documented_in_mild;
    double p_infect_given_contact, asymptomatic_transmissibility;
    double mean_time_to_isolate, mean_time_to_isolate_asympt, mean_time_to_severe;
    double mean_time_mild_recovery, mean_time_to_critical, mean_time_severe_recovery;
    double mean_time_critical_recovery, mean_time_to_death;
    double time_to_activation_mean, time_to_activation_std, initial_infected_fraction;
} oracle_params;

typedef struct {
    const int64_t *households;     /* [n, W], -1 terminated rows */
    const int64_t *age;            /* [n] */
    const int64_t *diabetes;       /* [n] */
    const int64_t *hypertension;   /* [n] */
    const int64_t *group_offsets;  /* [n_ages+1]  CSR form of the reference's age_groups tuple */
    const int64_t *group_members;  /* [n] */
    const double *contact_matrix;  /* [n_ages, n_ages] */
    const double *p_mild_severe;   /* [n_ages,2,2] */
    const double *p_severe_critical;
    const double *p_critical_death;
    const int64_t *iso_factor;     /* [n_iso_rows,3] */
    const double *p_infect_household; /* [n] */
    const double *fraction_stay_home; /* [n_ages] (MT provider) */
    /* keyed providers: prologue results supplied by the host code under test */
    const uint8_t *home_real;      /* [n] */
    const int64_t *initial_infected; /* [n_initial] */
    /* keyed replay: exp(-C) before / after lockdown, [n_ages,n_ages] each */
    const double *enlam_pre, *enlam_post;
    /* keyed native: exp(-p*w*rowsum) [2 phases][n_ages][2: 0=symptomatic,1=E], row alias tables */
    const double *nat_enlam;
    const double *nat_alias_prob;   /* [n_ages,n_ages] */
    const int32_t *nat_alias_idx;   /* [n_ages,n_ages] */
} oracle_inputs;

typedef struct {
    uint8_t *S, *E, *Mild, *Documented, *Severe, *Critical, *R, *D, *Q; /* [T,n] */
    double *num_infected_by, *time_documented, *time_to_activation, *time_to_death, *time_to_recovery;
    double *time_critical, *time_exposed, *num_infected_asympt, *time_infected, *time_to_severe; /* [n] */
    /* extras (not returned by the reference; exposed for deeper checks) */
    double *time_to_isolate, *time_severe, *time_to_critical; /* [n] */
    int32_t *num_infected_by_outside; /* [n] */
    uint8_t *home_real_out;           /* [n] */
    int64_t *initial_infected_out;    /* [int(frac*n)] */
    uint8_t *traced;                  /* [n] */
    int64_t *draw_count;              /* [1] MT words consumed / keyed blocks drawn */
} oracle_outputs;

#define INFECTED_BY_W 100

typedef struct {
    const oracle_params *p;
    const oracle_inputs *in;
    oracle_outputs *out;
    mt_state mt;
    int32_t *infected_by; /* [n,100] */
    int64_t blocks;
} ctx_t;

/* ---- provider-dispatched draws ---------------------------------------- */
static inline sd_block K(ctx_t *c, int64_t agent, int64_t day, uint32_t site, uint32_t sub)
{
    c->blocks++;
    return sd_draw(c->p->seed, (uint32_t)agent, (uint32_t)day, site, sub);
}

static inline double U(ctx_t *c, int64_t agent, int64_t day, uint32_t site, uint32_t sub, int lane)
{
    if (c->p->provider == ORACLE_MT) return mt_double(&c->mt);
    return sd_lane(K(c, agent, day, site, sub), lane);
}

/* functions.py:16-17 */
static inline double threshold_exponential(ctx_t *c, double mean, int64_t agent, int64_t day, uint32_t site,
                                           uint32_t sub, int lane)
{
    if (c->p->provider == ORACLE_MT) return rint(-log(1.0 - mt_double(&c->mt)) * mean);
    return sd_exp_days(sd_lane(K(c, agent, day, site, sub), lane), mean);
}

/* functions.py:19-25 */
static inline double threshold_log_normal(ctx_t *c, double mean, double sigma, int64_t agent, int64_t day,
                                          uint32_t site, uint32_t sub_base)
{
    if (c->p->provider == ORACLE_MT) {
        double x = exp(mean + sigma * mt_gauss(&c->mt));
        if (x <= 0) return 1;
        return rint(x);
    }
    c->blocks++; /* at least one polar call */
    return sd_lognormal_days(c->p->seed, (uint32_t)agent, (uint32_t)day, site, sub_base, mean, sigma);
}

/* seir_individual.py:46-51 */
static inline int64_t get_isolation_factor(const ctx_t *c, int64_t age)
{
    const int64_t *f = c->in->iso_factor;
    for (int64_t i = 0; i < c->p->n_iso_rows; i++)
        if (age >= f[i * 3 + 0] && age <= f[i * 3 + 1]) return f[i * 3 + 2];
    return 1;
}

/* seir_individual.py:60-88 */
static void do_contact_tracing(ctx_t *c, int64_t i, int64_t t)
{
    const oracle_params *p = c->p;
    oracle_outputs *o = c->out;
    const int64_t n = p->n, W = p->W;
    for (int64_t j = 0; j < W; j++) {
        int64_t contact = c->in->households[i * W + j];
        if (contact == -1) break;
        if (!o->S[(t - 1) * n + contact] && !o->traced[contact] &&
            U(c, i, t, SD_SITE_TR_HH, (uint32_t)j, 0) < p->p_trace_household) {
            o->Q[t * n + contact] = 1;
            o->Documented[t * n + contact] = 1;
            o->traced[contact] = 1;
            o->time_documented[contact] = (double)t;
            do_contact_tracing(c, contact, t);
        }
    }
    for (int64_t j = 0; j < INFECTED_BY_W; j++) {
        int64_t contact = c->infected_by[i * INFECTED_BY_W + j];
        if (contact == -1) break;
        if (U(c, i, t, SD_SITE_TR_OUT, (uint32_t)j, 0) < p->p_trace_outside) {
            o->Q[t * n + contact] = 1;
            o->Documented[t * n + contact] = 1;
            o->time_documented[contact] = (double)t;
            o->traced[contact] = 1;
            do_contact_tracing(c, contact, t);
        }
    }
}

/* bookkeeping of one successful infection, :347-356 / :376-385.  `ord` is the event
 * ordinal inside the infector's day (keyed providers key the infectee's timers on it). */
static inline void infect(ctx_t *c, int64_t i, int64_t contact, int64_t t, uint32_t ord)
{
    const oracle_params *p = c->p;
    oracle_outputs *o = c->out;
    const int64_t n = p->n;
    o->E[t * n + contact] = 1;
    o->num_infected_by[contact] = 0;
    o->num_infected_by_outside[contact] = 0;
    o->num_infected_asympt[contact] = 0;
    o->S[t * n + contact] = 0;
    o->time_to_isolate[contact] = threshold_exponential(
        c, p->mean_time_to_isolate_asympt * (double)get_isolation_factor(c, c->in->age[contact]), i, t,
        SD_SITE_INF, ord << 8, 0);
    if (o->time_to_isolate[contact] == 0) o->Q[t * n + contact] = 1;
    o->time_exposed[contact] = (double)t;
    o->time_to_activation[contact] =
        threshold_log_normal(c, p->time_to_activation_mean, p->time_to_activation_std, i, t, SD_SITE_INF, ord << 8);
}

int oracle_run_model(const oracle_params *p, const oracle_inputs *in, oracle_outputs *o)
{
    const int64_t n = p->n, T = p->T, n_ages = p->n_ages, W = p->W;
    ctx_t cx;
    ctx_t *c = &cx;
    memset(c, 0, sizeof(*c));
    c->p = p;
    c->in = in;
    c->out = o;
    const int mt = (p->provider == ORACLE_MT);
    const int native = (p->provider == ORACLE_KEYED_NATIVE);

    /* :147-161 */
    int contact_tracing = 0; /* `contact_tracing` is False after :161 whichever way the param was set */
    const int tracing_enabled = (int)p->tracing_enabled;

    /* :163 */
    if (mt) mt_init(&c->mt, (uint32_t)p->seed);

    /* :165-174 (zero-initialised by the caller is NOT assumed) */
    size_t Tn = (size_t)T * (size_t)n;
    memset(o->S, 0, Tn); memset(o->E, 0, Tn); memset(o->Mild, 0, Tn); memset(o->Documented, 0, Tn);
    memset(o->Severe, 0, Tn); memset(o->Critical, 0, Tn); memset(o->R, 0, Tn); memset(o->D, 0, Tn);
    memset(o->Q, 0, Tn);
    memset(o->traced, 0, (size_t)n);

    /* :176-182 shelter-in-place assignment */
    uint8_t *Home_real = o->home_real_out;
    memset(Home_real, 0, (size_t)n);
    int64_t n_init = (int64_t)(p->initial_infected_fraction * (double)n);
    int64_t *initial_infected = o->initial_infected_out;
    if (mt) {
        int64_t *matches = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
        for (int64_t a = 0; a < n_ages; a++) {
            int64_t m = 0;
            for (int64_t i = 0; i < n; i++)
                if (in->age[i] == a) matches[m++] = i;
            if (m > 0) {
                /* np.random.choice(matches, k, replace=False) == permutation(matches)[:k]
                   (randomimpl.py:2084-2100): the whole copy is shuffled even when k == 0 */
                int64_t k = (int64_t)(in->fraction_stay_home[a] * (double)m);
                mt_shuffle_i64(&c->mt, matches, m);
                for (int64_t j = 0; j < k; j++) Home_real[matches[j]] = 1;
            }
        }
        /* :187 */
        for (int64_t i = 0; i < n; i++) matches[i] = i;
        mt_shuffle_i64(&c->mt, matches, n);
        for (int64_t j = 0; j < n_init; j++) initial_infected[j] = matches[j];
        free(matches);
    } else {
        memcpy(Home_real, in->home_real, (size_t)n);
        n_init = p->n_initial;
        for (int64_t j = 0; j < n_init; j++) initial_infected[j] = in->initial_infected[j];
    }
    /* :184-186 */
    const uint8_t *Home = NULL; /* NULL == dummy_Home (all False) */

    /* :188 */
    memset(o->S, 1, (size_t)n);

    /* :197-198 */
    c->infected_by = (int32_t *)malloc(sizeof(int32_t) * (size_t)n * INFECTED_BY_W);
    if (!c->infected_by) return 2;
    memset(c->infected_by, 0xff, sizeof(int32_t) * (size_t)n * INFECTED_BY_W);

    /* :200-218 */
    double *time_exposed = o->time_exposed, *time_infected = o->time_infected, *time_severe = o->time_severe;
    double *time_critical = o->time_critical, *time_documented = o->time_documented;
    double *num_infected_by = o->num_infected_by, *num_infected_asympt = o->num_infected_asympt;
    int32_t *num_infected_by_outside = o->num_infected_by_outside;
    double *time_to_severe = o->time_to_severe, *time_to_recovery = o->time_to_recovery;
    double *time_to_critical = o->time_to_critical, *time_to_death = o->time_to_death;
    double *time_to_isolate = o->time_to_isolate, *time_to_activation = o->time_to_activation;
    for (int64_t i = 0; i < n; i++) {
        time_exposed[i] = -1; time_infected[i] = 0; time_severe[i] = 0; time_critical[i] = 0;
        time_documented[i] = 0; num_infected_by[i] = -1; num_infected_by_outside[i] = -1;
        num_infected_asympt[i] = -1; time_to_severe[i] = 0; time_to_recovery[i] = 0;
        time_to_critical[i] = 0; time_to_death[i] = 0; time_to_isolate[i] = 0; time_to_activation[i] = 0;
    }
    uint8_t *S = o->S, *E = o->E, *Mild = o->Mild, *Documented = o->Documented, *Severe = o->Severe;
    uint8_t *Critical = o->Critical, *R = o->R, *D = o->D, *Q = o->Q, *traced = o->traced;

    /* :220-227 */
    for (int64_t k = 0; k < n_init; k++) {
        int64_t a = initial_infected[k];
        E[a] = 1;
        S[a] = 0;
        time_exposed[a] = 0;
        num_infected_by[a] = 0;
        num_infected_by_outside[a] = 0;
        num_infected_asympt[a] = 0;
        time_to_activation[a] =
            threshold_log_normal(c, p->time_to_activation_mean, p->time_to_activation_std, a, 0, SD_SITE_INIT, 0);
    }

    /* the matrix in force; :240 rebinds contact_matrix = contact_matrix / lockdown_factor */
    double *cm = (double *)malloc(sizeof(double) * (size_t)(n_ages * n_ages));
    memcpy(cm, in->contact_matrix, sizeof(double) * (size_t)(n_ages * n_ages));
    int phase = 0;

    const int64_t *age = in->age, *diabetes = in->diabetes, *hypertension = in->hypertension;
    const int64_t *households = in->households;

    for (int64_t t = 1; t < T; t++) { /* :231 */
        if (t == p->t_lockdown) { /* :234-240 */
            for (int64_t k = 0; k < n_ages * n_ages; k++) cm[k] = cm[k] / p->lockdown_factor;
            phase = 1;
        }
        if (t == p->t_tracing_start && tracing_enabled) contact_tracing = 1; /* :241-242 */
        if (t == p->t_stayhome_start) Home = Home_real;                      /* :243-244 */
        /* :245-253 */
        memcpy(S + t * n, S + (t - 1) * n, (size_t)n);
        memcpy(E + t * n, E + (t - 1) * n, (size_t)n);
        memcpy(Mild + t * n, Mild + (t - 1) * n, (size_t)n);
        memcpy(Documented + t * n, Documented + (t - 1) * n, (size_t)n);
        memcpy(Severe + t * n, Severe + (t - 1) * n, (size_t)n);
        memcpy(Critical + t * n, Critical + (t - 1) * n, (size_t)n);
        memcpy(R + t * n, R + (t - 1) * n, (size_t)n);
        memcpy(D + t * n, D + (t - 1) * n, (size_t)n);
        memcpy(Q + t * n, Q + (t - 1) * n, (size_t)n);
        const uint8_t *Sp = S + (t - 1) * n, *Ep = E + (t - 1) * n, *Mp = Mild + (t - 1) * n;
        const uint8_t *Dp = Documented + (t - 1) * n, *Sevp = Severe + (t - 1) * n;
        const uint8_t *Cp = Critical + (t - 1) * n, *Qp = Q + (t - 1) * n;
        uint8_t *En = E + t * n, *Mn = Mild + t * n, *Dn = Documented + t * n;   /* (S[t] is written by infect()) */
        uint8_t *Sevn = Severe + t * n, *Cn = Critical + t * n, *Rn = R + t * n, *Deadn = D + t * n, *Qn = Q + t * n;
        const double td = (double)t;

        for (int64_t i = 0; i < n; i++) { /* :254 */
            /* exposed -> (mildly) infected  :256-272 */
            if (Ep[i]) {
                if (td - time_exposed[i] == time_to_activation[i]) {
                    Mn[i] = 1;
                    time_infected[i] = td;
                    En[i] = 0;
                    int64_t k3 = age[i] * 4 + diabetes[i] * 2 + hypertension[i];
                    if (U(c, i, t, SD_SITE_ACT, 0, 0) < in->p_mild_severe[k3]) {
                        time_to_severe[i] = threshold_exponential(c, p->mean_time_to_severe, i, t, SD_SITE_ACT, 0, 1);
                        time_to_recovery[i] = INFINITY;
                    } else {
                        time_to_recovery[i] =
                            threshold_exponential(c, p->mean_time_mild_recovery, i, t, SD_SITE_ACT, 0, 1);
                        time_to_severe[i] = INFINITY;
                    }
                    time_to_isolate[i] = threshold_exponential(
                        c, p->mean_time_to_isolate * (double)get_isolation_factor(c, age[i]), i, t, SD_SITE_ACT, 1, 0);
                    if (time_to_isolate[i] == 0) Qn[i] = 1;
                }
            }
            /* symptomatic individuals :274-327 */
            if (Mp[i] || Sevp[i] || Cp[i]) {
                if (td - time_infected[i] == time_to_recovery[i]) { /* :276-279 */
                    Rn[i] = 1;
                    Mn[i] = Sevn[i] = Cn[i] = Qn[i] = 0;
                    continue;
                }
                if (Mp[i] && !Dp[i]) { /* :280-289 */
                    if (U(c, i, t, SD_SITE_DAY, 0, 0) < p->p_documented_in_mild) {
                        Dn[i] = 1;
                        time_documented[i] = td;
                        traced[i] = 1;
                        if (contact_tracing) {
                            Qn[i] = 1;
                            do_contact_tracing(c, i, t);
                        }
                    }
                }
                if (Mp[i] && td - time_infected[i] == time_to_severe[i]) { /* :291-311 */
                    Mn[i] = 0;
                    Sevn[i] = 1;
                    if (!Dp[i]) {
                        Dn[i] = 1;
                        time_documented[i] = td;
                        traced[i] = 1;
                        if (contact_tracing) {
                            Qn[i] = 1;
                            do_contact_tracing(c, i, t);
                        }
                    }
                    Qn[i] = 1;
                    time_severe[i] = td;
                    int64_t k3 = age[i] * 4 + diabetes[i] * 2 + hypertension[i];
                    if (U(c, i, t, SD_SITE_SEV, 0, 0) < in->p_severe_critical[k3]) {
                        time_to_critical[i] = threshold_exponential(c, p->mean_time_to_critical, i, t, SD_SITE_SEV, 0, 1);
                        time_to_recovery[i] = INFINITY;
                    } else {
                        time_to_recovery[i] =
                            threshold_exponential(c, p->mean_time_severe_recovery, i, t, SD_SITE_SEV, 0, 1) +
                            time_to_severe[i];
                        time_to_critical[i] = INFINITY;
                    }
                } else if (Sevp[i] && td - time_severe[i] == time_to_critical[i]) { /* :312-321 */
                    Sevn[i] = 0;
                    Cn[i] = 1;
                    time_critical[i] = td;
                    int64_t k3 = age[i] * 4 + diabetes[i] * 2 + hypertension[i];
                    if (U(c, i, t, SD_SITE_CRIT, 0, 0) < in->p_critical_death[k3]) {
                        time_to_death[i] = threshold_exponential(c, p->mean_time_to_death, i, t, SD_SITE_CRIT, 0, 1);
                        time_to_recovery[i] = INFINITY;
                    } else {
                        time_to_recovery[i] =
                            threshold_exponential(c, p->mean_time_critical_recovery, i, t, SD_SITE_CRIT, 0, 1) +
                            time_to_severe[i] + time_to_critical[i];
                        time_to_death[i] = INFINITY;
                    }
                } else if (Cp[i]) { /* :323-327 */
                    if (td - time_critical[i] == time_to_death[i]) {
                        Cn[i] = 0;
                        Qn[i] = 0;
                        Deadn[i] = 1;
                    }
                }
            }
            if (Ep[i] || Mp[i] || Sevp[i] || Cp[i]) { /* :328 */
                if (!Qp[i]) {                          /* :330 */
                    if (!Ep[i] && td - time_infected[i] == time_to_isolate[i]) { /* :332-334 */
                        Qn[i] = 1;
                        continue;
                    }
                    if (Ep[i] && td - time_exposed[i] == time_to_isolate[i]) { /* :335-337 */
                        Qn[i] = 1;
                        continue;
                    }
                    /* infect within family :339-359 */
                    for (int64_t j = 0; j < W; j++) {
                        if (households[i * W + j] == -1) break;
                        int64_t contact = households[i * W + j];
                        double infectiousness = in->p_infect_household[i];
                        if (Ep[i]) infectiousness *= p->asymptomatic_transmissibility;
                        if (Sp[contact] && U(c, i, t, SD_SITE_HH, (uint32_t)j >> 1, (int)(j & 1)) < infectiousness) {
                            infect(c, i, contact, t, SD_ORD_HH(j));
                            num_infected_by[i] += 1;
                            if (Ep[i]) num_infected_asympt[i] += 1;
                        }
                    }
                    /* infect across families :361-390 */
                    if (!(Home && Home[i])) {
                        double infectiousness = p->p_infect_given_contact;
                        if (Ep[i]) infectiousness *= p->asymptomatic_transmissibility;
                        if (!native) {
                            for (int64_t contact_age = 0; contact_age < n_ages; contact_age++) {
                                int64_t glen = in->group_offsets[contact_age + 1] - in->group_offsets[contact_age];
                                if (glen == 0) continue; /* :368-369 */
                                int64_t num_contacts;
                                if (mt)
                                    num_contacts = mt_poisson(&c->mt, cm[age[i] * n_ages + contact_age]);
                                else {
                                    const double *en = phase ? in->enlam_post : in->enlam_pre;
                                    c->blocks++;
                                    num_contacts = sd_poisson(p->seed, (uint32_t)i, (uint32_t)t, SD_SITE_POIS,
                                                              (uint32_t)contact_age << 8,
                                                              en[age[i] * n_ages + contact_age], SD_MAX_CONTACTS_PER_AGE);
                                }
                                for (int64_t j = 0; j < num_contacts; j++) {
                                    int64_t contact;
                                    if (mt) {
                                        if (!(mt_double(&c->mt) < infectiousness)) continue; /* :373 */
                                        contact = in->group_members[in->group_offsets[contact_age] +
                                                                    mt_randint(&c->mt, glen)]; /* :374 */
                                    } else {
                                        sd_block b = K(c, i, t, SD_SITE_KEEP, ((uint32_t)contact_age << 8) | (uint32_t)j);
                                        if (!(sd_lane(b, 0) < infectiousness)) continue;
                                        contact = in->group_members[in->group_offsets[contact_age] +
                                                                    (int64_t)sd_index(b, (uint64_t)glen)];
                                    }
                                    if (Sp[contact] && !(Home && Home[contact])) { /* :375 */
                                        infect(c, i, contact, t, SD_ORD_COMM_REPLAY(contact_age, j));
                                        num_infected_by[i] += 1;
                                        if (num_infected_by_outside[i] < INFECTED_BY_W) /* reference: unchecked OOB */
                                            c->infected_by[i * INFECTED_BY_W + num_infected_by_outside[i]] = (int32_t)contact;
                                        num_infected_by_outside[i] += 1;
                                        if (Ep[i]) num_infected_asympt[i] += 1;
                                    }
                                }
                            }
                        } else {
                            /* distribution-identical aggregation of :367-374:
                               kept contacts ~ Poisson(p*w*rowsum(C[age_i,:] over non-empty groups)),
                               each kept contact's age ~ C[age_i,:]/rowsum (alias table), member uniform in that group */
                            const double en = in->nat_enlam[((int64_t)phase * n_ages + age[i]) * 2 + (Ep[i] ? 1 : 0)];
                            uint32_t kept = sd_poisson_day(p->seed, (uint32_t)i, (uint32_t)t,
                                                           K(c, i, t, SD_SITE_DAY, 0), en, SD_MAX_CONTACTS_NATIVE);
                            for (uint32_t k = 0; k < kept; k++) {
                                sd_block b = K(c, i, t, SD_SITE_NPICK, k);
                                int64_t contact_age = sd_alias_pick(sd_lane(b, 0), in->nat_alias_prob + age[i] * n_ages,
                                                                    in->nat_alias_idx + age[i] * n_ages, (int)n_ages);
                                int64_t glen = in->group_offsets[contact_age + 1] - in->group_offsets[contact_age];
                                if (glen == 0) continue; /* zero-weight columns are never picked; belt and braces */
                                int64_t contact = in->group_members[in->group_offsets[contact_age] +
                                                                    (int64_t)sd_index(b, (uint64_t)glen)];
                                if (Sp[contact] && !(Home && Home[contact])) {
                                    infect(c, i, contact, t, SD_ORD_COMM_NATIVE(k));
                                    num_infected_by[i] += 1;
                                    if (num_infected_by_outside[i] < INFECTED_BY_W)
                                        c->infected_by[i * INFECTED_BY_W + num_infected_by_outside[i]] = (int32_t)contact;
                                    num_infected_by_outside[i] += 1;
                                    if (Ep[i]) num_infected_asympt[i] += 1;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    if (o->draw_count) o->draw_count[0] = mt ? 0 : c->blocks;
    free(cm);
    free(c->infected_by);
    return 0;
}

/* ---- small helpers exported for unit tests of the draw contract -------- */
void oracle_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out)
{
    sd_block b = sd_philox(k0, k1, c0, c1, c2, c3);
    out[0] = b.w[0]; out[1] = b.w[1]; out[2] = b.w[2]; out[3] = b.w[3];
}
void oracle_sd_log(const double *x, double *y, int64_t m) { for (int64_t i = 0; i < m; i++) y[i] = sd_log(x[i]); }
void oracle_sd_exp(const double *x, double *y, int64_t m) { for (int64_t i = 0; i < m; i++) y[i] = sd_exp(x[i]); }
void oracle_sd_exp_days(const double *u, double mean, double *y, int64_t m)
{
    for (int64_t i = 0; i < m; i++) y[i] = sd_exp_days(u[i], mean);
}
void oracle_sd_lognormal_days(uint64_t seed, int64_t m, double mu, double sigma, double *y)
{
    for (int64_t i = 0; i < m; i++) y[i] = sd_lognormal_days(seed, (uint32_t)i, 0, SD_SITE_INIT, 0, mu, sigma);
}
void oracle_sd_poisson(uint64_t seed, int64_t m, double enlam, int64_t *y)
{
    for (int64_t i = 0; i < m; i++) y[i] = sd_poisson(seed, (uint32_t)i, 0, SD_SITE_NCNT, 0, enlam, SD_MAX_CONTACTS_NATIVE);
}
/* scalar keyed draws, one per draw site of the contract: what oracle/make_instrumented_golden.py binds (ctypes, called
 * from numba's nopython code) in place of every np.random.* / functions.threshold_* call of the reference's run_model */
double oracle_kd_uniform(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub, int lane)
{
    return sd_lane(sd_draw(seed, agent, day, site, sub), lane);
}
double oracle_kd_exp_days(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub, int lane, double mean)
{
    return sd_exp_days(sd_lane(sd_draw(seed, agent, day, site, sub), lane), mean);
}
double oracle_kd_lognormal_days(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub_base,
                                double mu, double sigma)
{
    return sd_lognormal_days(seed, agent, day, site, sub_base, mu, sigma);
}
int64_t oracle_kd_poisson(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub_base, double enlam)
{
    return (int64_t)sd_poisson(seed, agent, day, site, sub_base, enlam, SD_MAX_CONTACTS_PER_AGE);
}
int64_t oracle_kd_index(uint64_t seed, uint32_t agent, uint32_t day, uint32_t site, uint32_t sub, int64_t glen)
{
    return (int64_t)sd_index(sd_draw(seed, agent, day, site, sub), (uint64_t)glen);
}
/* MT stream helpers (pin the generator restatement against numpy's identical MT19937) */
void oracle_mt_doubles(uint32_t seed, double *y, int64_t m)
{
    mt_state s; mt_init(&s, seed);
    for (int64_t i = 0; i < m; i++) y[i] = mt_double(&s);
}

/* keyed prologue (include/seir_draws.h, sd_perm): the first k entries of the keyed permutation of [0, n) -- the picks
 * the device makes for Home_real (site SD_SITE_HOME, group = age) and initial_infected (site SD_SITE_CASES, group 0) */
void oracle_perm_picks(uint64_t seed, uint32_t site, uint32_t group, uint32_t n, int64_t k, int64_t *out)
{
    sd_perm P = sd_perm_make(seed, site, group, n);
    for (int64_t r = 0; r < k; r++) out[r] = (int64_t)sd_perm_at(&P, (uint32_t)r);
}

