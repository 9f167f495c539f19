"""Host-side populations for the day-step kernels.

`synthetic_population` is the fast vectorised generator used for scale tests and bench.py
(SURVEY.md section 8d "Synthetic inputs"): contiguous households with the per-household size pmf
measured on the reference's Italian sampler, ages from a smooth Italy/China-like pyramid,
comorbidities rising with age.  It reproduces the LAYOUT the reference's samplers emit
(sample_households.py:344-349: every household is a contiguous index range, each member's row lists
the others in ascending order, -1 padded), which is the only thing the kernels depend on.

`sample_households` / `sample_joint_comorbidities` are the entry points `run_complete_simulation`
uses (seir_individual.py:24-37): vectorised restatements of the reference's census-driven samplers.
"""
import os

import numpy as np

N_AGES = 101

# per-HOUSEHOLD size pmf (sizes 1..6); SURVEY.md section 8d: Italy ~[0.31,0.28,0.18,0.16,0.05,0.02]
HOUSEHOLD_SIZE_PMF = {
    "Italy": np.array([0.31, 0.28, 0.18, 0.16, 0.05, 0.02]),
    # China: singles, couples, parents + 1-2 children, three generations (up to 6)
    "China": np.array([0.20, 0.25, 0.27, 0.15, 0.08, 0.05]),
}
MAX_HOUSEHOLD_WIDTH = {"Italy": 6, "China": 5}  # sample_households.py:166, :38


def _age_pmf(country):
    a = np.arange(N_AGES, dtype=np.float64)
    if country == "China":
        base = 1.0 + 0.35 * np.exp(-((a - 31) / 6.0) ** 2) + 0.45 * np.exp(-((a - 50) / 8.0) ** 2)
        tail = 1.0 / (1.0 + np.exp((a - 72) / 6.5))
    else:
        base = 0.85 + 0.75 * np.exp(-((a - 50) / 13.0) ** 2)
        tail = 1.0 / (1.0 + np.exp((a - 84) / 5.5))
    p = base * tail
    return p / p.sum()


def _threads():
    try:
        return max(1, min(16, len(os.sched_getaffinity(0))))
    except AttributeError:
        return 4


def _parallel(fn, items):
    """fn over items on a thread pool (NumPy releases the GIL inside its loops)."""
    items = list(items)
    if len(items) <= 1 or _threads() == 1:
        return [fn(x) for x in items]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(_threads()) as ex:
        return list(ex.map(fn, items))


def households_from_sizes(sizes, n, width):
    """All-pairs adjacency rows for contiguous households (layout of sample_households.py:344-349)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    starts = np.zeros(len(sizes), dtype=np.int64)
    np.cumsum(sizes[:-1], out=starts[1:])
    hh_of = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)[:n]
    households = np.empty((n, width), dtype=np.int64)
    j = np.arange(width, dtype=np.int64)[None, :]

    def fill(rng_):
        lo, hi = rng_
        h = hh_of[lo:hi]
        start = starts[h]
        pos = np.arange(lo, hi, dtype=np.int64) - start
        member = start[:, None] + j + (j >= pos[:, None])
        households[lo:hi] = np.where(j < (sizes[h][:, None] - 1), member, -1)

    step = max(1 << 18, -(-n // (4 * _threads())))
    _parallel(fill, [(lo, min(n, lo + step)) for lo in range(0, n, step)])
    return households


def synthetic_population(n, seed=0, country="Italy"):
    """Returns (households int64[n,W], age int64[n], diabetes int64[n], hypertension int64[n],
    age_groups tuple of 101 int64 arrays)."""
    rng = np.random.default_rng(seed)
    pmf = HOUSEHOLD_SIZE_PMF[country] / HOUSEHOLD_SIZE_PMF[country].sum()
    width = MAX_HOUSEHOLD_WIDTH[country]
    mean = float((pmf * np.arange(1, 7)).sum())
    sizes = rng.choice(np.arange(1, 7), size=int(n / mean * 1.05) + 16, p=pmf).astype(np.int64)
    csum = np.cumsum(sizes)
    while csum[-1] < n:
        sizes = np.concatenate([sizes, rng.choice(np.arange(1, 7), size=1024, p=pmf)])
        csum = np.cumsum(sizes)
    k = int(np.searchsorted(csum, n))  # first household whose end reaches n
    sizes = sizes[:k + 1].copy()
    sizes[k] -= csum[k] - n  # truncate the last household
    households = households_from_sizes(sizes, n, width)
    age = rng.choice(N_AGES, size=n, p=_age_pmf(country)).astype(np.int64)
    a = age.astype(np.float64)
    p_diab = 0.005 + 0.20 / (1.0 + np.exp(-(a - 62) / 9.0))
    diabetes = (rng.random(n) < p_diab).astype(np.int64)
    p_hyp = 0.02 + 0.55 / (1.0 + np.exp(-(a - 58) / 10.0))
    p_hyp = np.where(diabetes == 1, np.minimum(1.0, p_hyp * 1.6), p_hyp)
    hypertension = (rng.random(n) < p_hyp).astype(np.int64)
    order = np.argsort(age, kind="stable")
    bounds = np.searchsorted(age[order], np.arange(N_AGES + 1))
    age_groups = tuple(order[bounds[i]:bounds[i + 1]] for i in range(N_AGES))
    return households, age, diabetes, hypertension, age_groups


def age_groups_csr(age):
    """(offsets int64[102], members int64[n]) -- the CSR form of `tuple(np.where(age == a)[0] ...)` (:36)."""
    age = np.asarray(age)
    order = np.argsort(age.astype(np.uint8), kind="stable").astype(np.int64)   # (radix sort on bytes)
    offsets = np.zeros(N_AGES + 1, dtype=np.int64)
    np.cumsum(np.bincount(age, minlength=N_AGES)[:N_AGES], out=offsets[1:])
    return offsets, order


# --------------------------------------------------------------------------------------------------------------
# Census-driven samplers, vectorised.  Same model, same inputs, same output layout
# as the reference's sequential Python loops (sample_households.py:37-163 China, :165-352 Italy) -- 102 s at
# n = 10 M there, about a second here -- but the draws come from a NumPy Generator, so a population agrees with the
# reference's in distribution, not per seed (tests/test_population.py checks the marginals against the
# reference's own sampler).
_CENSUS = None


def _census():
    """Rows of World_Age_2019.csv / AgeSpecificFertility.csv: from the CWD when the files are there (the reference
    opens them relative to the CWD, sample_households.py:12,24), else the copy frozen by oracle/freeze_census.py."""
    global _CENSUS
    if _CENSUS is None:
        import json
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "census.json")) as f:
            _CENSUS = json.load(f)
    return _CENSUS


def _csv_row(path, country, count, encoding=None):
    import csv
    with open(path, encoding=encoding) as f:
        for row in csv.reader(f, delimiter=","):
            if row and row[0] == country:
                return np.array([float(x) for x in row[1:1 + count]])
    return None


def get_age_distribution(country):
    """sample_households.py:10-19."""
    if os.path.exists("World_Age_2019.csv"):
        row = _csv_row("World_Age_2019.csv", country, 101)
        if row is not None:
            return row
    try:
        return np.array(_census()["age_distribution"][country])
    except KeyError:
        raise ValueError("no age distribution for %r (World_Age_2019.csv is not in the CWD)" % country)


def get_mother_birth_age_distribution(country):
    """sample_households.py:22-33 (five-year bins 15-19 ... 45-49)."""
    if os.path.exists("AgeSpecificFertility.csv"):
        row = _csv_row("AgeSpecificFertility.csv", country, 7, "latin-1")
        if row is not None:
            return row
    try:
        return np.array(_census()["mother_birth_age_distribution"][country])
    except KeyError:
        raise ValueError("no fertility row for %r (AgeSpecificFertility.csv is not in the CWD)" % country)


def _rng_from_global():
    """A Generator seeded from NumPy's global stream: run_complete_simulation seeds that stream with the run's seed
    (seir_individual.py:19), so the population is a function of the seed as in the reference."""
    return np.random.default_rng(int(np.random.randint(0, 2**31 - 1)))


def _choice(rng, p, size):
    """size draws from the categorical distribution p (inverse CDF; p need not be normalised)."""
    c = np.cumsum(np.asarray(p, dtype=np.float64))
    return np.minimum(np.searchsorted(c, rng.random(size) * c[-1], side="right"), len(c) - 1).astype(np.int64)


def _layout(ages_by_pos, sizes, n, width, tail_age):
    """Contiguous households from a [K, 6] matrix of member ages and the household sizes; the households that do not
    fit before the reference's end-game (`if n - num_generated < tail: i = 0`, :66, :209) are replaced by
    single-person households drawn by `tail_age(count)`."""
    before = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    tail = width + 1 if width == 6 else 5          # Italy: max_household_size + 1 (:209); China: 5 (:66)
    k = int(np.searchsorted(n - before < tail, True))   # first household drawn with fewer than `tail` agents left
    if k >= len(sizes) and before[-1] + sizes[-1] < n - tail + 1:
        raise RuntimeError("not enough households were drawn")
    sizes = sizes[:k]
    n_regular = int(sizes.sum())
    n_single = n - n_regular
    pos = np.arange(6)[None, :]
    valid = pos < sizes[:, None]
    age = np.concatenate([ages_by_pos[:k][valid], tail_age(n_single)]).astype(np.int64)
    all_sizes = np.concatenate([sizes, np.ones(n_single, dtype=np.int64)])
    return households_from_sizes(all_sizes, n, width), age


def _blocks(rng, K):
    """(child generator, households) per block: the household attributes are drawn block-wise on a thread pool."""
    nb = max(1, min(4 * _threads(), K // (1 << 17)))
    cuts = np.linspace(0, K, nb + 1).astype(np.int64)
    return list(zip(rng.spawn(nb), np.diff(cuts)))


def _italy_block(args):
    """K households of sample_households.py:165-352: (member ages int16[K, 6], sizes int64[K])."""
    rng, K = args
    n_ages = N_AGES
    age_distribution = get_age_distribution("Italy")
    age_distribution = age_distribution / age_distribution.sum()
    # the eleven household types of :183-186 and how many people each holds
    household_probs = np.array([0.309179, 0.196000, 0.0694283, 0.0273065, 0.00450268, 0.152655, 0.132429, 0.0200969,
                                0.049821, 0.033, 0.017])
    type_size = np.array([1, 2, 2, 3, 4, 3, 4, 5, 2, 5, 6], dtype=np.int64)
    n_children = np.array([0, 0, 1, 2, 3, 1, 2, 3, 0, 2, 2], dtype=np.int64)
    mother = get_mother_birth_age_distribution("Italy")
    typ = _choice(rng, household_probs, K)
    sizes = type_size[typ]
    adult = lambda m: rng.integers(30, 101, size=m)                        # noqa: E731  np.random.randint(30,101)
    birth_age = lambda m: (_choice(rng, mother, m) + 3) * 5 + rng.integers(0, 5, size=m)  # noqa: E731  (:232)
    A = np.zeros((K, 6), dtype=np.int16)
    nc = n_children[typ]
    rows = np.arange(K)
    has_kids = nc > 0
    kid_rows = rows[has_kids]
    child = _choice(rng, age_distribution[:30], 3 * len(kid_rows)).reshape(-1, 3)   # renormalized_child (:201)
    child = np.where(nc[kid_rows][:, None] > np.arange(3)[None, :], child, 0)
    A[kid_rows, :3] = child
    m_cur = np.minimum(n_ages - 1, birth_age(len(kid_rows)) + child.max(axis=1))    # mother (:234, :246, ...)
    A[kid_rows, nc[kid_rows]] = m_cur
    couple = np.isin(typ[kid_rows], (5, 6, 7, 9, 10))
    A[kid_rows[couple], nc[kid_rows][couple] + 1] = np.minimum(n_ages - 1, m_cur[couple] + 3)
    t = rows[typ == 0]
    A[t, 0] = adult(len(t))
    t = rows[typ == 1]                                                      # couple, one three years older (:220-226)
    a1 = adult(len(t))
    A[t, 0] = a1
    A[t, 1] = np.minimum(n_ages - 1, a1 + 3)
    t = rows[typ == 8]                                                      # family without a nucleus (:297-300)
    A[t, 0] = adult(len(t))
    A[t, 1] = adult(len(t))
    t9 = typ[kid_rows] == 9                                                 # + one adult >= 60 (:303-316)
    A[kid_rows[t9], 4] = _choice(rng, age_distribution[60:], int(t9.sum())) + 60
    t10 = typ[kid_rows] == 10                                               # + grandparents (:321-337)
    gm = np.minimum(n_ages - 1, birth_age(int(t10.sum())) + m_cur[t10])
    A[kid_rows[t10], 4] = gm
    A[kid_rows[t10], 5] = np.minimum(n_ages - 1, gm + 3)
    return A, sizes


def _sample(block_fn, n, rng, mean_size, width, tail_age):
    K = int(n / mean_size * 1.02) + 64
    As, Ss, total = [], [], 0
    while total < n:
        out = _parallel(block_fn, _blocks(rng, K))
        As += [o[0] for o in out]
        Ss += [o[1] for o in out]
        total += int(sum(o[1].sum() for o in out))
        K = int(K * 0.1) + 64
    return _layout(np.concatenate(As), np.concatenate(Ss), n, width, tail_age)


def sample_households_italy(n, rng=None):
    """sample_households.py:165-352, vectorised: (households int64[n, 6], age int64[n])."""
    rng = rng or _rng_from_global()
    return _sample(_italy_block, n, rng, 2.37, 6, lambda m: rng.integers(30, 101, size=m))


def _china_age_distribution():
    d = np.array(_census()["china_age_distribution_2020"])
    return d / d.sum()


def _china_single(rng, m):
    """Age of a single-person household (:72-81)."""
    ad = _china_age_distribution()
    p_young = ad[22:26].sum() / (ad[22:26].sum() + ad[60:].sum())
    young = rng.random(m) < p_young
    return np.where(young, _choice(rng, ad[22:26], m) + 22, _choice(rng, ad[50:], m) + 50)


def _china_block(args):
    """K households of sample_households.py:37-163: (member ages int16[K, 6], sizes int64[K])."""
    rng, K = args
    ad = _china_age_distribution()
    household_probs = np.array([0.1703, .2117, 0.3557, 0.1126])             # :53
    max_child_age, max_birth_age = 26, 32
    p_one_child, p_single_parent = 1.7 / 3, 0.3
    typ = _choice(rng, household_probs, K)
    fam = typ >= 2
    two_parents = fam & (rng.random(K) < 1 - p_single_parent)               # :108, :131
    child = _choice(rng, ad[:max_child_age], K)
    second = fam & (rng.random(K) < 1 - p_one_child) & (child > 0)          # :112, :142
    sizes = np.where(typ == 0, 1, np.where(typ == 1, 2, 2 + two_parents + second + 2 * (typ == 3))).astype(np.int64)
    A = np.zeros((K, 6), dtype=np.int16)
    rows = np.arange(K)
    t = rows[typ == 0]
    A[t, 0] = _china_single(rng, len(t))
    t = rows[typ == 1]                                                      # couple (:89-94)
    A[t, 0] = _choice(rng, ad[max_child_age:], len(t)) + max_child_age
    A[t, 1] = _choice(rng, ad[max_child_age:], len(t)) + max_child_age
    m_cur = _choice(rng, ad[max_child_age:max_birth_age], K) + max_child_age + child   # :103-105
    A[fam, 0] = child[fam]
    A[fam, 1] = m_cur[fam]
    nxt = np.full(K, 2, dtype=np.int64)
    A[rows[two_parents], 2] = m_cur[two_parents]                            # the second parent has the mother's age (:109)
    nxt += two_parents
    t = typ == 3                                                            # grandparents (:134-140)
    gm = _choice(rng, ad[max_child_age:max_birth_age], K) + max_child_age + m_cur
    A[rows[t], nxt[t]] = gm[t]
    A[rows[t], nxt[t] + 1] = gm[t]
    nxt += 2 * t
    # the second child (:112-118, :142-148): choice(offset, p = ages [child - offset, child)) + child, as written
    age2 = np.zeros(K, dtype=np.int64)
    for c in range(1, max_child_age):
        sel = second & (child == c)
        m = int(sel.sum())
        if m:
            off = min(5, c)
            age2[sel] = _choice(rng, ad[c - off:c], m) + c
    A[rows[second], nxt[second]] = age2[second]
    return A, sizes


def sample_households_china(n, rng=None):
    """sample_households.py:37-163, vectorised: (households int64[n, 5], age int64[n])."""
    rng = rng or _rng_from_global()
    return _sample(_china_block, n, rng, 2.9, 5, lambda m: _china_single(rng, m))


def p_comorbidity(country, comorbidity):
    """Per-age prevalence tables of sample_comorbidities.py:92-175 for the two countries the drivers run."""
    a = np.arange(N_AGES)
    if country == "Italy":
        if comorbidity == "diabetes":   # :117-157, five-year bins (Global Burden of Disease)
            bins = np.array([0.0001, 0.0009, 0.0024, 0.0091, 0.0264, 0.0356, 0.0392, 0.0428, 0.0489, 0.0638, 0.0893,
                             0.1277, 0.1783, 0.2106, 0.2407, 0.2851, 0.3348, 0.3517, 0.3354])
            return bins[np.minimum(a // 5, 18)]
        if comorbidity == "hypertension":  # :159-173
            return np.where(a < 35, 0.14 * (a / 35.0), np.where(a < 39, 0.14, np.where(a < 44, 0.1,
                            np.where(a < 49, 0.16, np.where(a < 54, 0.3, 0.34)))))
    if country == "China":
        if comorbidity == "hypertension":  # p_hypertension_china (:58-72)
            edges = np.array([18, 25, 35, 45, 55, 65, 75])
            vals = np.array([0.0, 4.0, 6.1, 15.0, 29.6, 44.6, 55.7, 60.2]) / 100
            return vals[np.searchsorted(edges, a, side="right")]
        if comorbidity == "diabetes":      # p_diabetes_china (:3-56): sex-specific prevalences mixed by the sex ratio
            male = np.array([5.2755905511811, 8.346456692913385, 13.543307086614174, 17.95275590551181,
                             20.708661417322837, 21.653543307086615]) / 100
            female = np.array([4.015748031496061, 5.1181102362204705, 9.05511811023622, 17.401574803149607,
                               24.488188976377955, 25.196850393700785]) / 100
            ratio5 = np.array([113.91, 118.03, 118.62, 118.14, 112.89, 105.39, 101.05, 102.84, 103.75, 103.64, 102.15,
                               101.65, 100.5, 96.94, 94.42, 89.15, 76.97, 71.16, 48.74, 40.07])
            sex_ratio = ratio5[np.minimum(a // 5, 19)]
            sex_ratio = sex_ratio / (sex_ratio + 100)
            ad = np.array(_census()["china_age_distribution_2020"])
            intervals = [(18, 30), (30, 40), (40, 50), (50, 60), (60, 70), (70, 100)]
            out = np.zeros(N_AGES)
            for k, (lo, hi) in enumerate(intervals):
                w = ad[lo:hi] / ad[lo:hi].sum()
                pm = float(np.dot(w, sex_ratio[lo:hi]))
                out[lo:hi] = male[k] * pm + female[k] * (1 - pm)
            out[100:] = out[99]
            return out
    raise NotImplementedError("no %s prevalence table for %r: the drivers run Italy and China "
                              "(sample_comorbidities.py:92-175 holds further countries)" % (comorbidity, country))


def sample_joint_comorbidities(age, country="China", rng=None):
    """sample_comorbidities.py:74-90: diabetes ~ Bernoulli(p_diabetes[age]); hypertension given diabetes with
    probability 0.5, else with the probability that keeps the marginal prevalence.  (diabetes, hypertension) int64[n]."""
    rng = rng or _rng_from_global()
    age = np.asarray(age, dtype=np.int64)
    p_d, p_h = p_comorbidity(country, "diabetes"), p_comorbidity(country, "hypertension")
    p_h_given_d = 0.5
    p_h_given_not = (p_h - p_h_given_d * p_d) / (1 - p_d)
    diabetes = (rng.random(age.shape[0]) < p_d[age]).astype(np.int64)
    hypertension = (rng.random(age.shape[0]) < np.where(diabetes == 1, p_h_given_d, p_h_given_not[age])).astype(np.int64)
    return diabetes, hypertension


def _reference_module(name):
    """The reference's own host module (`SEIR_B200_POPULATION=reference`): needs the reference repository on
    sys.path and its census files in the CWD.  Its `np.int` / `np.bool8` aliases (removed in NumPy 1.24 / 2.0) are
    put back first.  Raises ImportError when the module is not there -- there is no silent substitute."""
    import importlib
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    mod = importlib.import_module(name)
    if os.path.dirname(os.path.abspath(mod.__file__)) == os.path.dirname(os.path.abspath(__file__)):
        raise ImportError("SEIR_B200_POPULATION=reference: %s resolves to this package, not to the reference" % name)
    return mod


def _source():
    src = os.environ.get("SEIR_B200_POPULATION", "census")
    if src not in ("census", "reference", "synthetic"):
        raise ValueError("SEIR_B200_POPULATION must be census, reference or synthetic")
    return src


def sample_households(country, n):
    """The dispatch of run_complete_simulation (seir_individual.py:24-35).  The reference has samplers for China and
    Italy only -- its Germany / UK / Spain / UN branches call functions that do not exist -- so anything else raises.
    SEIR_B200_POPULATION selects the implementation: `census` (default) the vectorised restatement above;
    `reference` the reference's own sequential sampler (same population as the reference for the same NumPy seed;
    raises if its module or census files are missing); `synthetic` the census-free generator (scale tests only)."""
    if country not in ("China", "Italy"):
        raise NotImplementedError("the reference has no household sampler for %r (seir_individual.py:26-35 calls "
                                  "functions that sample_households.py does not define)" % (country,))
    src = _source()
    if src == "synthetic":
        households, age, _, _, _ = synthetic_population(n, seed=int(np.random.randint(0, 2**31 - 1)), country=country)
        return households, age
    if src == "reference":
        ref = _reference_module("sample_households")
        households, age = (ref.sample_households_china if country == "China" else ref.sample_households_italy)(n)
        return np.asarray(households, dtype=np.int64), np.asarray(age, dtype=np.int64)
    return sample_households_china(n) if country == "China" else sample_households_italy(n)


def sample_comorbidities(age, country):
    """`sample_joint_comorbidities(age, country)` as run_complete_simulation calls it (:37), same dispatch."""
    src = _source()
    if src == "reference":
        d, h = _reference_module("sample_comorbidities").sample_joint_comorbidities(np.asarray(age), country)
        return np.asarray(d, dtype=np.int64), np.asarray(h, dtype=np.int64)
    return sample_joint_comorbidities(age, country)


def census_population(n, seed=0, country="Italy"):
    """(households, age, diabetes, hypertension, age_groups) from the census-driven samplers, seeded explicitly --
    what bench.py and the scale tests build."""
    rng = np.random.default_rng([int(seed), 0xC0FFEE])
    households, age = (sample_households_china if country == "China" else sample_households_italy)(n, rng)
    diabetes, hypertension = sample_joint_comorbidities(age, country, rng)
    offsets, members = age_groups_csr(age)
    age_groups = tuple(members[offsets[i]:offsets[i + 1]] for i in range(N_AGES))
    return households, age, diabetes, hypertension, age_groups
