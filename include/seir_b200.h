/*
 * seir_b200.h -- C ABI of libseir_b200.so, the B200 (sm_100a) implementation of the daily
 * agent-stepping path of the COVID19-Demography individual-level SEIR simulator.
 *
 * Drop-in boundary: the reference's `run_model(seed, households, age, age_groups, diabetes,
 * hypertension, contact_matrix, p_mild_severe, p_severe_critical, p_critical_death,
 * mean_time_to_isolate_factor, lockdown_factor_age, p_infect_household, fraction_stay_home, params)`
 * at /root/reference/seir_individual.py:90-392, returning the 20-tuple of :392.  The reference has
 * no FFI of its own (it is a numba function); the binding a maintainer adds is the ctypes stub in
 * INTEGRATION.md / covid19-demography_b200/engine.py.
 *
 * Everything here is plain C: pointers, sizes, POD structs.  Host arrays are BORROWED for the
 * duration of the call; outputs go to CALLER-allocated buffers; device memory belongs to the
 * handle.  Every function returns 0 on success, non-zero on error (seir_last_error() gives the
 * thread-local message).  There is no CPU fallback: without a CUDA device seir_create fails.
 */
#ifndef SEIR_B200_H
#define SEIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEIR_N_AGES 101            /* seir_individual.py:153, run_simulation.py:17 */
#define SEIR_N_TALLY 9             /* S,E,Mild,Documented,Severe,Critical,R,D,Q -- order of :392 */
#define SEIR_MODE_NATIVE 0         /* community: Poisson(kept contacts) + alias-table contact age   */
#define SEIR_MODE_REPLAY 1         /* community: 101 Poisson draws per sender-day, as :367-374  */
#define SEIR_LAUNCH_PERSISTENT 0   /* one cooperative kernel over all days, one grid barrier per day */
#define SEIR_LAUNCH_PER_DAY 1      /* one kernel launch per day (debug / comparison)            */
#define SEIR_HISTORY_EAGER 0       /* every run writes the reference's T x n history (one packed byte per agent-day)  */
#define SEIR_HISTORY_ON_DEMAND 1   /* no T x n buffer: seir_get_history computes the rows asked for from the per-agent
                                      records of the LAST run (ensembles and sweeps that consume tallies only)      */

/* scalar parameters: the `params` dict of seir_individual.py:135-158 */
typedef struct {
    int64_t n;                 /* params['n']                                   :152 */
    int32_t T;                 /* params['T']                                   :147 */
    int32_t t_lockdown;        /* params['t_lockdown']                          :149 */
    int32_t t_stayhome_start;  /* params['t_stayhome_start']                    :158 */
    int32_t t_tracing_start;   /* params['t_tracing_start']                     :157 */
    int32_t tracing_enabled;   /* value of `tracing_enabled` at :241 (the reference as compiled: 1; 0 = the paper's
                                  no-tracing semantics).  When tracing_enabled && 1 <= t_tracing_start < T,
                                  do_contact_tracing (:60-88) runs on the device from day t_tracing_start */
    int32_t mode;              /* SEIR_MODE_*                                         */
    int32_t launch;            /* SEIR_LAUNCH_*                                       */
    int32_t history_mode;      /* SEIR_HISTORY_* (read by seir_create only)          */
    double p_documented_in_mild;          /* :143 */
    double p_infect_given_contact;        /* :146 */
    double asymptomatic_transmissibility; /* :145 */
    double mean_time_to_severe;           /* :140 */
    double mean_time_mild_recovery;       /* :141 */
    double mean_time_to_critical;         /* :142 */
    double mean_time_severe_recovery;     /* :139 */
    double mean_time_critical_recovery;   /* :138 */
    double mean_time_to_death;            /* :137 */
    double time_to_activation_mean;       /* :135 */
    double time_to_activation_std;        /* :136 */
    double p_trace_household;             /* :156 */
    double p_trace_outside;               /* :155 */
} seir_params;

/* static population: arguments households/age/age_groups/diabetes/hypertension of :91.
 * Households must be the contiguous all-pairs layout both reference samplers emit
 * (sample_households.py:344-349): agent i's row lists the other members of
 * [i - pos, i - pos + size) in ascending order, -1 padded. */
typedef struct {
    int64_t n;
    int32_t W;                          /* households.shape[1]                       */
    int32_t reserved0;
    const int64_t *households;          /* [n, W]                                    */
    const int64_t *age;                 /* [n] in [0, 100]                           */
    const int64_t *diabetes;            /* [n] in {0,1}                              */
    const int64_t *hypertension;        /* [n] in {0,1}                              */
    const int64_t *group_offsets;       /* [102]  CSR of the age_groups tuple (:36)  */
    const int64_t *group_members;       /* [n]                                       */
    const double *p_infect_household;   /* [n] (:108); a constant vector is detected and kept as a scalar */
} seir_population;

/* draw tables, computed by the host code (covid19-demography_b200/tables.py) */
typedef struct {
    const double *p_mild_severe;        /* [101,2,2]  :103 */
    const double *p_severe_critical;    /* [101,2,2]  :104 */
    const double *p_critical_death;     /* [101,2,2]  :105 */
    const double *iso_mean;             /* [101] mean_time_to_isolate * get_isolation_factor(age)         :270 */
    const double *iso_asympt_mean;      /* [101] mean_time_to_isolate_asympt * get_isolation_factor(age)  :352 */
    const double *nat_enlam;            /* [2,101,2] exp(-p*w*rowsum) : phase (0 before lockdown, 1 after), age, w (0 symptomatic, 1 E) */
    const double *nat_alias_prob;       /* [101,101] alias table (Vose) of each contact-matrix row over non-empty age groups */
    const int32_t *nat_alias_idx;       /* [101,101] alias column                                                 */
    const double *enlam_pre;            /* [101,101] exp(-C)                  (replay mode; may be NULL in native mode) */
    const double *enlam_post;           /* [101,101] exp(-C/lockdown_factor)  (replay mode)                             */
} seir_tables;

/* per-run inputs (one per replica): the prologue of :176-187.
 * keyed_prologue == 0: the caller made the picks (tables.prologue replays the reference's Mersenne-Twister picks):
 *   home_real / initial_infected are copied to the device.
 * keyed_prologue != 0: the DEVICE makes them from the keyed contract (include/seir_draws.h, sd_perm): per age the
 *   first int(fraction_stay_home[a] * N_a) entries of a keyed permutation of the age group, and the first
 *   int(initial_infected_fraction * n) entries of a keyed permutation of the agents -- the same counts and the same
 *   uniform-subset distribution as :178-187, nothing but 101 doubles crosses PCIe (tables.prologue_keyed is the host
 *   restatement of the same picks). */
typedef struct {
    uint64_t seed;                      /* Philox key                                 */
    const uint8_t *home_real;           /* [n] 0/1 shelter-in-place assignment, or NULL (nobody) */
    const int64_t *initial_infected;    /* [n_initial] agent ids exposed at t=0 (:187)  */
    int64_t n_initial;
    const double *fraction_stay_home;   /* [101] (:178-182), keyed prologue only; NULL = nobody */
    double initial_infected_fraction;   /* (:186), keyed prologue only                */
    int32_t keyed_prologue;
    int32_t reserved0;
} seir_replica;

/* Optional per-replica overrides: one launch batches runs of DIFFERENT grid points of a sweep
 * (parameter_sweep.py:54-56,120-126: p_infect_given_contact x mortality_multiplier x start date).  What the grid
 * varies reaches the device as a set of draw tables (severity tables and the kept-contact Poisson table depend on
 * the first two) and as day numbers (the start date moves T and the switch days); every other scalar is the
 * handle's.  T must not exceed the handle's T. */
typedef struct {
    int32_t T, t_lockdown, t_stayhome_start, t_tracing_start;
    int32_t table_set;                  /* index into the sets registered with seir_set_table_bank */
    int32_t reserved0;
} seir_overrides;

/* per-agent output vectors, as doubles with the reference's sentinels (:200-218) */
enum {
    SEIR_VEC_NUM_INFECTED_BY = 0,
    SEIR_VEC_TIME_DOCUMENTED = 1,
    SEIR_VEC_TIME_TO_ACTIVATION = 2,
    SEIR_VEC_TIME_TO_DEATH = 3,
    SEIR_VEC_TIME_TO_RECOVERY = 4,
    SEIR_VEC_TIME_CRITICAL = 5,
    SEIR_VEC_TIME_EXPOSED = 6,
    SEIR_VEC_NUM_INFECTED_ASYMPT = 7,
    SEIR_VEC_TIME_INFECTED = 8,
    SEIR_VEC_TIME_TO_SEVERE = 9,
    /* not returned by the reference, exposed for parity checks */
    SEIR_VEC_TIME_TO_ISOLATE = 10,
    SEIR_VEC_TIME_SEVERE = 11,
    SEIR_VEC_TIME_TO_CRITICAL = 12,
    SEIR_VEC_NUM_INFECTED_OUTSIDE = 13,
    SEIR_VEC_COUNT = 14
};

/* history planes for seir_get_history, in the order of the reference's return tuple (:392) */
enum { SEIR_H_S = 0, SEIR_H_E, SEIR_H_MILD, SEIR_H_DOCUMENTED, SEIR_H_SEVERE, SEIR_H_CRITICAL,
       SEIR_H_R, SEIR_H_D, SEIR_H_Q, SEIR_H_PACKED /* raw byte: bits0-2 compartment, bit3 Q, bit4 Documented */ };
#define SEIR_H_PREVIOUS 16         /* OR into `which`: read the history set aside by seir_history_preserve */
#define SEIR_VEC_PREVIOUS 256      /* OR into the vector id: read the per-agent vector of the run set aside by seir_history_preserve */

typedef struct seir_handle seir_handle;

/* Upload the static population and tables; allocate state for n_replicas concurrent runs. */
int seir_create(const seir_params *params, const seir_population *pop, const seir_tables *tables,
                int n_replicas, int device, seir_handle **out);

/* One population across the GPUs of a box (shards of whole households; every rank creates
 * its handle on its own device with the FULL population and n_replicas = 1).
 * seir_shard_export: SEIR_SHARD_HANDLES x 64 bytes of CUDA IPC handles of this rank's winner / inbox / list / flag / infection-day arrays.
 * seir_shard_attach: all ranks' handles (world x 5 x 64 bytes, rank-major) and the shard bounds
 *   (bounds[world + 1], bounds[g] a household start): rank `rank` owns agents [bounds[rank], bounds[rank+1]).
 * Afterwards seir_run steps only the owned agents; a community contact whose target lives on another GPU is
 * resolved in that GPU's memory over NVLink (system-scope atomics), and the GPUs meet at a device-side
 * barrier once per day.  All ranks must call seir_run together, with the same seed and initial cases.
 * Outputs (tallies, vectors, history) cover the owned agents; tallies of the ranks add up.
 * A sharded run starts REPLICATED: while a day's lists hold at most `entries` agents (seir_shard_set_threshold, default
 * 393216) every GPU steps the whole population on its own -- the runs are deterministic, so the GPUs stay identical
 * without exchanging a byte, and a sparse day costs what it costs on one GPU -- and on the first longer day all GPUs
 * switch for good to sharded passes.  0 = sharded from day 1.  All ranks must use the same value. */
#define SEIR_SHARD_HANDLES 6
int seir_shard_export(seir_handle *h, uint8_t *handles_out);
int seir_shard_attach(seir_handle *h, int rank, int world, const uint8_t *all_handles, const int64_t *bounds);
int seir_shard_set_threshold(seir_handle *h, int64_t entries);

/* Replace the scalar parameters / tables (same population) before the next seir_run. */
int seir_set_params(seir_handle *h, const seir_params *params, const seir_tables *tables);

/* Run all replicas for T days.  reps[n_replicas].  Blocking. */
int seir_run(seir_handle *h, const seir_replica *reps);
/* The same with per-replica overrides (overrides[n_replicas], or NULL = seir_run). */
int seir_run_ex(seir_handle *h, const seir_replica *reps, const seir_overrides *overrides);
/* Register n_sets table sets for seir_overrides.table_set (set 0 replaces the handle's own tables);
 * p_infect_given_contact[n_sets] is the scalar of :146 each set was built for (replay mode reads it). */
int seir_set_table_bank(seir_handle *h, int n_sets, const seir_tables *sets, const double *p_infect_given_contact);
/* The picks of the last run's prologue as the device holds them: home_out[n] 0/1 (all 0 when stay-home never
 * starts within T), init_out[cap] agent ids, *n_init their number. */
int seir_get_prologue(seir_handle *h, int replica, uint8_t *home_out, int64_t *init_out, int64_t cap, int64_t *n_init);

/* Device time of the last seir_run, CUDA events on the launching stream:
 * *day_loop_ms = the day loop (persistent kernel or T launches),
 * *total_ms = state init + day loop + materialisation of the T x n packed history. */
int seir_last_run_ms(seir_handle *h, float *day_loop_ms, float *total_ms);
/* The three timed segments of the last run: state init, day loop, history materialisation. */
int seir_last_run_ms3(seir_handle *h, float *init_ms, float *day_loop_ms, float *materialize_ms);

/* Per-day compartment counts by age, out[T][9][101] int64, planes in the order of :392
 * (what the drivers compute with X.sum(axis=1) and the per-age-group masks). */
int seir_get_tallies(seir_handle *h, int replica, int64_t *out);

/* Exposure-day cohorts, out[T][4][101] int64: for agents exposed on day d of age a, [0] their number,
 * [1] the sum of their num_infected_by, [2] how many are dead at the end (D[-1]), [3] recovered (R[-1]).
 * Everything run_simulation.py:448-453 / parameter_sweep.py:572-600 derive per exposure day (R0, CFR,
 * age fractions, median age, CFR by age group) follows from these counts. */
int seir_get_cohorts(seir_handle *h, int replica, int64_t *out);

/* One per-agent vector, out[n] doubles (`which` = SEIR_VEC_*, optionally | SEIR_VEC_PREVIOUS). */
int seir_get_agent_vector(seir_handle *h, int replica, int which, double *out);

/* History plane `which` for days [t0, t1), out[(t1-t0)*n] bytes (0/1, or the packed byte). */
int seir_get_history(seir_handle *h, int replica, int which, int t0, int t1, uint8_t *out);
/* The reference returns fresh arrays from every run_model call; here the T x n history lives on the device and the next
 * seir_run overwrites it.  seir_history_preserve sets the packed history of the LAST run aside (one device-to-device
 * copy, 0.2 ms at n = 10 M, T = 60) so that it stays readable (which | SEIR_H_PREVIOUS) across the next run; the records
 * behind the per-agent vectors are set aside with it (32 B per agent; seir_get_agent_vector with which | SEIR_VEC_PREVIOUS). */
int seir_history_preserve(seir_handle *h);

/* Counters of the last run: out[0]=kernels launched, out[1]=sum over days of active agents
 * A(t-1) (E/Mild/Severe/Critical), out[2]=infections I (excluding the initial ones),
 * out[3]=transmission hits (>= infections: same-day multiple hits count). */
int seir_get_counters(seir_handle *h, int replica, int64_t *out4);
/* List entries the day loop of the last run visited (new infectees + agents that may transmit + parked agents
 * on their due days): the work actually done, against the A(t-1) sum of seir_get_counters. */
int seir_get_visits(seir_handle *h, int replica, int64_t *out);

/* %globaltimer (ns) stamped by CTA 0 at the start of every day pass of the last run: out[T+2],
 * out[t] = start of pass t (1 <= t <= T), out[T+1] = end of the persistent day loop (0 in per-day mode). */
int seir_get_day_times(seir_handle *h, int64_t *out);
/* Stage stamps of CTA 0 in the first round of every pass: out[(T+2)*8], per pass t the 7 stamps
 * pass set-up done / stage A / C / H / W done / lists flushed / tallies flushed (0 where not reached). */
int seir_get_stage_times(seir_handle *h, int64_t *out);

/* debug: raw copy of one of the parallel tracer's per-replica arrays of the last tracing day
 * (0 ctl {n_roots, n_reached, alloc, n_big}, 1 roots, 2 representatives, 3 component sizes, 4 reached, 5 leaders, 6 parents) */
int seir_debug_trace_array(seir_handle *h, int which, int replica, void *out, uint64_t bytes);

/* HOST ONLY (no device is touched): the prologue picks the reference makes for `seed` -- seir_individual.py:163 seeds
 * numba's MT19937, :178-182 draw np.random.choice(matches, int(fraction_stay_home[a]*len), replace=False) per non-empty
 * age group, :187 np.random.choice(n, int(initial_infected_fraction*n), replace=False) -- so that the drop-in run_model
 * infects and shelters the very agents the reference would.  group_offsets[n_ages+1] / group_members[n]: the age_groups
 * tuple (:36) in CSR form, members ascending (np.where order).  fraction_stay_home may be NULL (all zero).
 * home_real[n] (0/1) and initial_infected[*n_initial] are written; initial_cap is the room in initial_infected.
 * Thread-safe; each calling thread keeps 4 B per agent of scratch between calls (its page faults cost more than the draws). */
int seir_prologue_mt(uint32_t seed, int64_t n, int n_ages, const int64_t *group_offsets, const int64_t *group_members,
                     const double *fraction_stay_home, double initial_infected_fraction, uint8_t *home_real,
                     int64_t *initial_infected, int64_t initial_cap, int64_t *n_initial);

/* HOST ONLY: the bytes `pd.DataFrame(a).to_csv(sep="\t", index=False, header=False, na_rep="NA")` writes for a float64
 * matrix a[rows][cols] -- what parameter_sweep.py:716-729 and simulate_agepolicies.py:646-716 write per job and datatype:
 * every value as Python's repr (shortest round-trip digits; "3.0", "0.1", "1e-05", "inf"), NaN as NA, tab-separated, one
 * line per row.  out must hold 27 bytes per value + 1 per row + 1; *written is the length. */
int seir_format_tsv(const double *a, int64_t rows, int64_t cols, char *out, int64_t cap, int64_t *written);

/* Page-lock / unlock a caller-owned host buffer so the per-run copies in seir_run are DMA transfers. */
int seir_host_register(void *ptr, uint64_t bytes);
int seir_host_unregister(void *ptr);

void seir_destroy(seir_handle *h);
const char *seir_last_error(void);
int seir_device_count(void);
/* library self-description (arch the kernels were compiled for, build flags) */
const char *seir_build_info(void);
/* sizeof of the ABI's structs as compiled: seir_params, seir_population, seir_tables, seir_replica, seir_overrides
 * (a binding checks its own struct definitions against these) */
void seir_abi_sizes(int64_t out5[5]);

/* draw-contract probes (device evaluation of include/seir_draws.h, for parity tests) */
int seir_probe_draws(int device, uint64_t seed, int64_t m, double mean, double mu, double sigma, double enlam,
                     double *exp_days_out, double *lognormal_out, int64_t *poisson_out, double *log_out,
                     const double *log_in);

#ifdef __cplusplus
}
#endif
#endif /* SEIR_B200_H */
