"""The vectorised census samplers (covid19-demography_b200/population.py) against the reference's own samplers.

tests/golden/ref_population_stats.npz holds marginals of /root/reference/sample_households.py and
sample_comorbidities.py (4 seeds x 200 000 agents per country, frozen by oracle/make_population_golden.py).
The restatement draws from a different generator, so the comparison is statistical: chi-square of the pooled
histograms and z-scores of the conditional means, with thresholds a correct sampler passes with overwhelming
probability and the typical restatement mistakes (a wrong age window, a swapped household type, a missing
`min(100, .)`) do not."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from covid19_demography_b200 import population  # noqa: E402
import make_population_golden as gold  # noqa: E402  (only its stats() helper: pure NumPy)

G = np.load(os.path.join(ROOT, "tests", "golden", "ref_population_stats.npz"))
N = int(G["n"]) * len(G["seeds"])


def mine(country):
    acc = None
    for seed in range(len(G["seeds"])):
        hh, age, d, h, _ = population.census_population(int(G["n"]), seed=seed, country=country)
        st = gold.stats(hh, age, d, h)
        acc = st if acc is None else {k: acc[k] + st[k] for k in st}
    return acc


def chi2_two_sample(a, b, min_count=40):
    """Chi-square statistic per degree of freedom of two count vectors of equal totals (bins pooled below min_count)."""
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    keep = (a + b) >= min_count
    a2 = np.append(a[keep], a[~keep].sum())
    b2 = np.append(b[keep], b[~keep].sum())
    ok = (a2 + b2) > 0
    stat = (((a2 - b2) ** 2)[ok] / (a2 + b2)[ok]).sum()
    return stat / max(1, ok.sum() - 1)


@pytest.mark.parametrize("country", ["Italy", "China"])
def test_household_and_age_marginals_match_the_reference_sampler(country):
    m = mine(country)
    ref = {k[len(country) + 1:]: G[k] for k in G.files if k.startswith(country + "_")}
    assert m["age_hist"].sum() == ref["age_hist"].sum() == N
    # households are whole units: per-agent counts of one household are correlated, so the histograms of
    # household-level quantities are over-dispersed by at most the household size (6)
    assert chi2_two_sample(m["size_hist"] / np.maximum(1, np.arange(7)), ref["size_hist"] / np.maximum(1, np.arange(7))) < 3.0
    assert chi2_two_sample(m["age_hist"], ref["age_hist"]) < 3.0
    assert chi2_two_sample(m["size_decade"], ref["size_decade"]) < 4.0
    # the same (size, position) slots are populated, with the same mean age (roles sit at fixed positions)
    assert ((m["pos_cnt"] > 0) == (ref["pos_cnt"] > 0)).all()
    both = ref["pos_cnt"] > 200
    mean_m, mean_r = m["age_sum"][both] / m["pos_cnt"][both], ref["age_sum"][both] / ref["pos_cnt"][both]
    se = 30.0 * np.sqrt(1.0 / m["pos_cnt"][both] + 1.0 / ref["pos_cnt"][both])   # ages spread over < 30 years (sd)
    assert (np.abs(mean_m - mean_r) < 5 * se + 0.05).all(), (mean_m, mean_r)


@pytest.mark.parametrize("country", ["Italy", "China"])
def test_comorbidity_prevalences_match_the_reference_sampler(country):
    m = mine(country)
    ref = {k[len(country) + 1:]: G[k] for k in G.files if k.startswith(country + "_")}
    for key in ("diab_by_age", "hyp_by_age", "both_by_age"):
        na, nb = m["age_hist"].astype(float), ref["age_hist"].astype(float)
        pa, pb = m[key] / np.maximum(na, 1), ref[key] / np.maximum(nb, 1)
        p = (m[key] + ref[key]) / np.maximum(na + nb, 1)
        se = np.sqrt(np.maximum(p * (1 - p), 1e-9) * (1 / np.maximum(na, 1) + 1 / np.maximum(nb, 1)))
        z = np.abs(pa - pb) / se
        big = (na > 500) & (nb > 500)
        assert z[big].max() < 5.0, (key, int(np.argmax(np.where(big, z, 0))), z[big].max())
        assert (p[big] == 0).sum() == ((pa[big] == 0) & (pb[big] == 0)).sum()


def test_layout_is_the_reference_layout():
    """Contiguous households, all-pairs rows in ascending order, -1 padded (sample_households.py:344-349); the
    end-game fills the last slots with single-person households (:209, :66)."""
    for country, W in (("Italy", 6), ("China", 5)):
        hh, age, d, h, groups = population.census_population(5003, seed=3, country=country)
        assert hh.shape == (5003, W) and hh.dtype == np.int64 and age.dtype == np.int64
        i = 0
        while i < 5003:
            size = int((hh[i] >= 0).sum()) + 1
            members = list(range(i, i + size))
            for k in members:
                others = [x for x in members if x != k]
                assert list(hh[k][:len(others)]) == others and (hh[k][len(others):] == -1).all()
            i += size
        assert i == 5003
        assert age.min() >= 0 and age.max() <= 100 and set(np.unique(d)) <= {0, 1} and set(np.unique(h)) <= {0, 1}
        assert sum(len(g) for g in groups) == 5003 and all((age[g] == a).all() for a, g in enumerate(groups))


def test_unsupported_country_raises_like_the_reference():
    with pytest.raises(NotImplementedError):
        population.sample_households("Germany", 100)


def test_global_seed_makes_the_population_reproducible():
    np.random.seed(5)
    a = population.sample_households("Italy", 2000)
    np.random.seed(5)
    b = population.sample_households("Italy", 2000)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
