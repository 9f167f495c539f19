"""Host-side populations for the day-step kernels.

`synthetic_population` is the fast vectorised generator used for scale tests and bench.py
(SURVEY.md section 8d "Synthetic inputs"): contiguous households with the per-household size pmf
measured on the reference's Italian sampler, ages from a smooth Italy/China-like pyramid,
comorbidities rising with age.  It reproduces the LAYOUT the reference's samplers emit
(sample_households.py:344-349: every household is a contiguous index range, each member's row lists
the others in ascending order, -1 padded), which is the only thing the kernels depend on.

`sample_households` / `sample_joint_comorbidities` are the entry points `run_complete_simulation`
uses (seir_individual.py:24-37).
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


def households_from_sizes(sizes, n, width):
    """All-pairs adjacency rows for contiguous households (layout of sample_households.py:344-349)."""
    starts = np.zeros(len(sizes), dtype=np.int64)
    np.cumsum(sizes[:-1], out=starts[1:])
    hh_of = np.repeat(np.arange(len(sizes), dtype=np.int64), sizes)[:n]
    start = starts[hh_of]
    size = sizes[hh_of]
    pos = np.arange(n, dtype=np.int64) - start
    j = np.arange(width, dtype=np.int64)[None, :]
    member = start[:, None] + j + (j >= pos[:, None])
    households = np.where(j < (size[:, None] - 1), member, -1).astype(np.int64)
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
    order = np.argsort(age, kind="stable").astype(np.int64)
    offsets = np.searchsorted(age[order], np.arange(N_AGES + 1)).astype(np.int64)
    return offsets, order


def _reference_module(name):
    """The reference's own host module `name` (sample_households / sample_comorbidities) when the reference
    repository is on sys.path -- the drop-in scenario of INTEGRATION.md -- else None.  Those modules use the
    NumPy aliases `np.int` / `np.bool8`, removed in NumPy 1.24 / 2.0; they are put back before the import."""
    import importlib
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    try:
        mod = importlib.import_module(name)
    except Exception:
        return None
    return mod if hasattr(mod, "sample_households_italy") or hasattr(mod, "sample_joint_comorbidities") else None


def sample_households(country, n):
    """Entry used by `run_complete_simulation` (seir_individual.py:24-35).

    Household sampling stays the reference's own Python host code (BASELINE.json north_star): when the
    reference's `sample_households` module is importable (the drop-in scenario: the reference repository is the
    CWD / on sys.path, its census files next to it) it is called as the reference calls it (:25, :27).  Otherwise
    -- tests, benchmarks, the GPU box -- the synthetic population above is drawn, seeded from NumPy's global
    stream the way the reference's samplers are (:19)."""
    ref = _reference_module("sample_households")
    if ref is not None and os.environ.get("SEIR_B200_POPULATION", "reference") == "reference":
        try:
            fn = ref.sample_households_china if country == "China" else ref.sample_households_italy
            households, age = fn(n)
            return np.asarray(households, dtype=np.int64), np.asarray(age, dtype=np.int64)
        except (OSError, IOError):  # the census CSV files are not in the CWD
            pass
    seed = int(np.random.randint(0, 2**31 - 1))
    households, age, _, _, _ = synthetic_population(n, seed=seed, country="China" if country == "China" else "Italy")
    sample_households._last_seed = seed
    return households, age


def sample_joint_comorbidities(age, country="Italy"):
    """(diabetes, hypertension) int64[n] -- sample_comorbidities.py:74-82 interface; the reference's own module when
    it is importable (see sample_households), else the synthetic prevalences below."""
    ref = _reference_module("sample_comorbidities")
    if ref is not None and os.environ.get("SEIR_B200_POPULATION", "reference") == "reference":
        d, h = ref.sample_joint_comorbidities(np.asarray(age), country)
        return np.asarray(d, dtype=np.int64), np.asarray(h, dtype=np.int64)
    return _synthetic_comorbidities(age)


def _synthetic_comorbidities(age):
    rng = np.random.default_rng(int(np.random.randint(0, 2**31 - 1)))
    a = np.asarray(age, dtype=np.float64)
    n = len(a)
    p_diab = 0.005 + 0.20 / (1.0 + np.exp(-(a - 62) / 9.0))
    diabetes = (rng.random(n) < p_diab).astype(np.int64)
    p_hyp = 0.02 + 0.55 / (1.0 + np.exp(-(a - 58) / 10.0))
    p_hyp = np.where(diabetes == 1, np.minimum(1.0, p_hyp * 1.6), p_hyp)
    hypertension = (rng.random(n) < p_hyp).astype(np.int64)
    return diabetes, hypertension
