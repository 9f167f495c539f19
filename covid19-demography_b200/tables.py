"""Host-side construction of the draw tables the kernels consume (include/seir_b200.h, seir_tables).

Everything here is plain NumPy run once per parameter set; the kernels (and the CPU oracle in the
parity tests) only ever COMPARE against these doubles, so no transcendental is evaluated on the
device for them.
"""
import numpy as np

N_AGES = 101

PARAM_KEYS = (
    "n", "n_ages", "T", "initial_infected_fraction", "t_lockdown", "lockdown_factor", "t_stayhome_start",
    "t_tracing_start", "contact_tracing", "p_trace_household", "p_trace_outside", "p_documented_in_mild",
    "p_infect_given_contact", "asymptomatic_transmissibility", "mean_time_to_isolate",
    "mean_time_to_isolate_asympt", "mean_time_to_severe", "mean_time_mild_recovery", "mean_time_to_critical",
    "mean_time_severe_recovery", "mean_time_critical_recovery", "mean_time_to_death",
    "time_to_activation_mean", "time_to_activation_std")
"""The 24 keys `run_model` reads (seir_individual.py:135-158)."""


def params_to_dict(params):
    """Accept a numba.typed.Dict or any Mapping[str, float]."""
    out = {}
    for k in PARAM_KEYS:
        try:
            out[k] = float(params[k])
        except KeyError:
            raise KeyError("params is missing %r (seir_individual.py:135-158)" % k)
    return out


def isolation_factor_lut(mean_time_to_isolate_factor, n_ages=N_AGES):
    """`get_isolation_factor` (seir_individual.py:46-51) for every age: first matching row wins, else 1."""
    rows = np.asarray(mean_time_to_isolate_factor)
    lut = np.ones(n_ages, dtype=np.float64)
    for age in range(n_ages):
        for lo, hi, f in rows:
            if age >= lo and age <= hi:
                lut[age] = float(f)
                break
    return lut


def alias_table(weights):
    """Walker/Vose alias table of a non-negative weight vector: (prob float64[K], alias int32[K]).
    A column b is drawn as: x = u*K, j = floor(x); j if x - j < prob[j] else alias[j]
    (include/seir_draws.h, sd_alias_pick).  Zero-weight columns get prob 0 and are never returned."""
    w = np.asarray(weights, dtype=np.float64)
    K = len(w)
    scaled = w * (K / w.sum())
    prob = np.zeros(K, dtype=np.float64)
    alias = np.arange(K, dtype=np.int32)
    small = [i for i in range(K) if scaled[i] < 1.0]
    large = [i for i in range(K) if scaled[i] >= 1.0]
    scaled = scaled.copy()
    while small and large:
        s_, l_ = small.pop(), large.pop()
        prob[s_] = scaled[s_]
        alias[s_] = l_
        scaled[l_] = (scaled[l_] + scaled[s_]) - 1.0
        (small if scaled[l_] < 1.0 else large).append(l_)
    for i in large:
        prob[i] = 1.0
    for i in small:  # numerical leftovers: weight ~1
        prob[i] = 1.0 if w[i] > 0 else 0.0
        if w[i] == 0:  # a zero-weight column must never be returned: alias it to the heaviest column
            alias[i] = int(np.argmax(w))
    return prob, alias


def build_tables(contact_matrix, group_sizes, params, mean_time_to_isolate_factor, p_mild_severe,
                 p_severe_critical, p_critical_death, with_replay=True, alias_from=None):
    """Returns a dict of C-contiguous float64 arrays keyed like the fields of `seir_tables`.
    alias_from: tables of an earlier call with the SAME contact matrix and group sizes -- the alias tables (the only
    part that costs milliseconds: a Python loop per row) do not depend on anything else, a sweep builds them once."""
    p = params
    C_pre = np.ascontiguousarray(contact_matrix, dtype=np.float64)
    if C_pre.shape != (N_AGES, N_AGES):
        raise ValueError("contact_matrix must be %dx%d" % (N_AGES, N_AGES))
    if (C_pre < 0).any():
        raise ValueError("poisson(): lambda < 0")  # numba/cpython/randomimpl.py:1680-1681
    C_post = C_pre / p["lockdown_factor"]  # seir_individual.py:240
    nonempty = np.asarray(group_sizes) > 0  # :368-369 skips empty age groups
    lut = isolation_factor_lut(mean_time_to_isolate_factor)
    nat_enlam = np.empty((2, N_AGES, 2), dtype=np.float64)
    for ph, C in enumerate((C_pre, C_post)):
        rowsum = np.where(nonempty[None, :], C, 0.0).sum(axis=1)
        for w, weight in enumerate((1.0, p["asymptomatic_transmissibility"])):
            nat_enlam[ph, :, w] = np.exp(-(p["p_infect_given_contact"] * weight) * rowsum)
    if alias_from is not None:
        alias_prob, alias_idx = alias_from["nat_alias_prob"], alias_from["nat_alias_idx"]
    else:
        Cz = np.where(nonempty[None, :], C_pre, 0.0)
        alias_prob = np.ones((N_AGES, N_AGES), dtype=np.float64)
        alias_idx = np.tile(np.arange(N_AGES, dtype=np.int32), (N_AGES, 1))
        for a in range(N_AGES):
            if Cz[a].sum() > 0:
                alias_prob[a], alias_idx[a] = alias_table(Cz[a])
    out = {
        "p_mild_severe": np.ascontiguousarray(p_mild_severe, dtype=np.float64).reshape(N_AGES, 2, 2),
        "p_severe_critical": np.ascontiguousarray(p_severe_critical, dtype=np.float64).reshape(N_AGES, 2, 2),
        "p_critical_death": np.ascontiguousarray(p_critical_death, dtype=np.float64).reshape(N_AGES, 2, 2),
        "iso_mean": np.ascontiguousarray(p["mean_time_to_isolate"] * lut),           # :270
        "iso_asympt_mean": np.ascontiguousarray(p["mean_time_to_isolate_asympt"] * lut),  # :352
        "nat_enlam": nat_enlam,
        "nat_alias_prob": alias_prob,
        "nat_alias_idx": alias_idx,
        "enlam_pre": np.exp(-C_pre) if with_replay else None,
        "enlam_post": np.exp(-C_post) if with_replay else None,
    }
    return out


SITE_HOME, SITE_CASES = 14, 15  # include/seir_draws.h


def _philox4x32(k0, k1, c0, c1, c2, c3):
    """Philox4x32-10 on NumPy arrays of counters (include/seir_draws.h, sd_philox)."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    lo32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & lo32, n2, p0 & lo32
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _fmix32(x):
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x85EBCA6B)) & m
    x ^= x >> np.uint64(13); x = (x * np.uint64(0xC2B2AE35)) & m
    x ^= x >> np.uint64(16)
    return x


def keyed_picks(seed, site, group, n, k):
    """The first k entries of the keyed permutation of [0, n) (include/seir_draws.h: sd_perm_make / sd_perm_at):
    a 6-round Feistel network over the smallest even number of bits holding n, cycle-walked into [0, n)."""
    if k <= 0:
        return np.zeros(0, dtype=np.int64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    a = _philox4x32(seed & 0xFFFFFFFF, seed >> 32, [group], [0], [site], [0])
    b = _philox4x32(seed & 0xFFFFFFFF, seed >> 32, [group], [0], [site], [1])
    keys = [np.uint64(int(a[j][0])) for j in range(4)] + [np.uint64(int(b[0][0])), np.uint64(int(b[1][0]))]
    bits = 0
    while bits < 32 and (1 << bits) < n:
        bits += 1
    half = np.uint64(1 if bits < 2 else (bits + 1) // 2)
    mask = np.uint64((1 << int(half)) - 1)
    x = np.arange(k, dtype=np.uint64)
    todo = np.ones(k, dtype=bool)
    while todo.any():
        xs = x[todo]
        L, R = xs >> half, xs & mask
        for r in range(6):
            f = _fmix32(R ^ keys[r]) & mask
            L, R = R, L ^ f
        xs = (L << half) | R
        x[todo] = xs
        todo[todo] = xs >= np.uint64(n)
    return x.astype(np.int64)


def prologue_keyed(seed, age_groups, n, fraction_stay_home, initial_infected_fraction):
    """The prologue's picks (seir_individual.py:176-187) under the KEYED contract -- what the device draws itself
    when a replica asks for it (Engine.run(..., keyed=(fraction_stay_home, initial_infected_fraction))): per age the
    first int(fraction * size) entries of a keyed permutation of the age group shelter at home, the first
    int(fraction * n) entries of a keyed permutation of the agents are the initial cases.  Same counts and the same
    uniform-subset distribution as the reference's `np.random.choice(..., replace=False)`; not its picks per seed
    (tables.prologue replays those).  Returns (home_real bool[n], initial_infected int64[k])."""
    home = np.zeros(n, dtype=np.bool_)
    fsh = np.asarray(fraction_stay_home, dtype=np.float64)
    for a, matches in enumerate(age_groups):
        m = len(matches)
        k = int(fsh[a] * m) if m else 0
        if k > 0:
            home[np.asarray(matches)[keyed_picks(seed, SITE_HOME, a, m, k)]] = True
    k0 = int(initial_infected_fraction * n)
    return home, keyed_picks(seed, SITE_CASES, 0, n, k0)


def prologue_fast(seed, age_groups, n, fraction_stay_home, initial_infected_fraction):
    """The prologue's picks (seir_individual.py:176-187) from the same distributions -- per age a uniform
    subset of int(fraction * size) agents shelters at home, int(fraction * n) uniform agents are the initial
    cases -- but from NumPy's PCG64 generator in O(picks) instead of replaying the reference's Mersenne-Twister
    stream (two full permutations, about 0.3 s per seed at n = 10 M).  For ensembles and sweeps, where the
    host prologue would otherwise cost thirty times the device run; NOT the reference's picks for that seed."""
    rng = np.random.default_rng([int(seed) & 0xFFFFFFFF, 0x5E1B200])
    home = np.zeros(n, dtype=np.bool_)
    fsh = np.asarray(fraction_stay_home, dtype=np.float64)
    for a, matches in enumerate(age_groups):
        k = int(fsh[a] * len(matches))
        if k > 0:
            home[np.asarray(matches)[rng.choice(len(matches), size=k, replace=False)]] = True
    k0 = int(initial_infected_fraction * n)
    init = rng.choice(n, size=k0, replace=False).astype(np.int64) if k0 > 0 else np.zeros(0, dtype=np.int64)
    return home, init


def prologue(seed, age_groups, n, fraction_stay_home, initial_infected_fraction):
    """Host restatement of the run_model prologue (seir_individual.py:163,176-187).

    The reference seeds numba's MT19937 with `seed` and draws, in this order, one
    `np.random.choice(matches, k, replace=False)` per non-empty age and one
    `np.random.choice(n, k, replace=False)`.  numba implements both as a Fisher-Yates shuffle with
    masked-rejection `randint` (numba/cpython/randomimpl.py:1927-1946, 2084-2100), which is exactly
    NumPy's legacy `RandomState.permutation`; so the SAME Home_real and initial_infected as the
    reference come out of `np.random.RandomState(seed)`.
    Returns (home_real bool[n], initial_infected int64[k]).
    """
    rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)
    home = np.zeros(n, dtype=np.bool_)
    fsh = np.asarray(fraction_stay_home, dtype=np.float64)
    for a, matches in enumerate(age_groups):
        m = len(matches)
        if m > 0:
            perm = rs.permutation(np.asarray(matches))
            k = int(fsh[a] * m)
            if k > 0:
                home[perm[:k]] = True
    k0 = int(initial_infected_fraction * n)
    init = rs.permutation(n)[:k0].astype(np.int64)
    return home, init
