"""Freeze summary statistics of the reference's OWN population samplers (run in the build container only).

    python oracle/make_population_golden.py       # writes tests/golden/ref_population_stats.npz

For Italy and China, `sample_households_{italy,china}(n)` and `sample_joint_comorbidities` of the unmodified
reference (/root/reference/sample_households.py, sample_comorbidities.py) are run for a few seeds and their
marginals stored: age histogram, household-size histogram (per agent), mean age by (household size, position),
the joint (household size, age decade) table, comorbidity counts by age.  tests/test_population.py compares the
vectorised restatement (covid19-demography_b200/population.py) against them.
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "ref_population_stats.npz")
N, SEEDS = 200_000, (0, 1, 2, 3)


def stats(households, age, diabetes, hypertension):
    households, age = np.asarray(households), np.asarray(age)
    n = len(age)
    size = (households >= 0).sum(axis=1) + 1
    first = np.where(size > 1, np.minimum(np.arange(n), np.where(households[:, 0] >= 0, households[:, 0], n)), np.arange(n))
    pos = np.arange(n) - first
    age_sum = np.zeros((7, 6)); cnt = np.zeros((7, 6))
    np.add.at(age_sum, (size, pos), age); np.add.at(cnt, (size, pos), 1)
    joint = np.zeros((7, 11)); np.add.at(joint, (size, age // 10), 1)
    return dict(age_hist=np.bincount(age, minlength=101), size_hist=np.bincount(size, minlength=7), age_sum=age_sum,
                pos_cnt=cnt, size_decade=joint, diab_by_age=np.bincount(age, weights=diabetes, minlength=101),
                hyp_by_age=np.bincount(age, weights=hypertension, minlength=101),
                both_by_age=np.bincount(age, weights=diabetes * hypertension, minlength=101))


if __name__ == "__main__":
    ref_shim.load()
    import sample_households as sh, sample_comorbidities as sc
    out = {"n": N, "seeds": np.array(SEEDS)}
    for country, fn in (("Italy", sh.sample_households_italy), ("China", sh.sample_households_china)):
        acc = None
        for seed in SEEDS:
            np.random.seed(seed)
            hh, age = fn(N)
            d, h = sc.sample_joint_comorbidities(age, country)
            st = stats(hh, age, d, h)
            acc = st if acc is None else {k: acc[k] + st[k] for k in st}
            print(country, seed, "done", flush=True)
        for k, v in acc.items():
            out["%s_%s" % (country, k)] = v
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes")
