"""-m gpu : parity at BASELINE.json's full size (configs[1] = C2: 10 000 000 agents, T = 60, Italy tables,
5 initial cases, lockdown day 46).  The keyed oracle still finishes such a run in tens of seconds, so the
first check is bit-exactness itself (ten per-agent vectors, per-day per-age tallies, the packed history
on sampled days); the rest are the size-independent properties of SURVEY.md section 4 tier T3 (compartments
partition the population, S / R / D / Documented monotone), a checksum of checksums (column sums of the
materialised T x n history == the tallies the day loop counted itself), idempotence (the same seed twice,
persistent and per-day launches) and replica independence at that size.

The population is the synthetic Lombardy-like one bench.py measures on.  When the host lacks the memory for
the oracle's nine T x n planes the test runs on 4 000 000 agents and says so."""
import os
import sys

import numpy as np
import pytest

import oracle
import util
from covid19_demography_b200 import population, tables as tables_mod

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = 60


def _mem_available():
    try:
        return int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        return 16 << 30


@pytest.fixture(scope="module")
def c2():
    import bench
    n = 10_000_000 if _mem_available() > (40 << 30) else 4_000_000
    print("full-size parity on n = %d agents (MemAvailable %.0f GB)" % (n, _mem_available() / 2.0**30))
    p, tabs = bench.workload_params(n, T)
    hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
    return {"n": n, "p": p, "tabs": tabs, "hh": hh, "age": age, "diab": diab, "hyp": hyp, "groups": groups,
            "p_hh": np.full(n, tabs["p_infect_household"]), "stay": np.zeros(101)}


def _tables(w, p):
    t = w["tabs"]
    return tables_mod.build_tables(t["contact_matrix"], np.array([len(g) for g in w["groups"]]), tables_mod.params_to_dict(p),
                                   t["mean_time_to_isolate_factor"], t["p_mild_severe"], t["p_severe_critical"],
                                   t["p_critical_death"])


def _engine(mod, w, p, **kw):
    return mod.Engine(tables_mod.params_to_dict(p), w["hh"], w["age"], w["groups"], w["diab"], w["hyp"], w["p_hh"],
                      _tables(w, p), mode="native", **kw)


def _oracle(w, p, seed, home, init):
    t = w["tabs"]
    return oracle.run_model(seed, w["hh"], w["age"], w["groups"], w["diab"], w["hyp"], t["contact_matrix"],
                            t["p_mild_severe"], t["p_severe_critical"], t["p_critical_death"],
                            t["mean_time_to_isolate_factor"], t["lockdown_factor_age"], w["p_hh"], w["stay"], p,
                            provider="native", tracing_enabled=1, home_real=home, initial_infected=init,
                            tables=_tables(w, p))


def _properties(tal, n):
    assert (tal[:, [0, 1, 2, 4, 5, 6, 7]].sum(axis=(1, 2)) == n).all(), "compartments are not a partition"
    per_age = tal[:, [0, 1, 2, 4, 5, 6, 7]].sum(axis=1)
    assert (per_age == per_age[0]).all(), "agents changed age group"
    S, Doc, R, D = (tal[:, k].sum(axis=1) for k in (0, 3, 6, 7))
    assert (np.diff(S) <= 0).all() and (np.diff(R) >= 0).all() and (np.diff(D) >= 0).all() and (np.diff(Doc) >= 0).all()
    assert (tal[:, 8] >= tal[:, 4] + tal[:, 5]).all(), "Severe / Critical implies Q"


def _checksum_of_checksums(eng, tal, age, days):
    for t in days:
        packed = eng.history("packed", t, t + 1)[0]
        planes = util.unpack_history(packed)
        for k in range(9):
            assert np.array_equal(np.bincount(age[planes[k]], minlength=101), tal[t, k]), \
                "day %d plane %s: the materialised history and the day loop's tallies disagree" % (t, util.HIST[k])


@pytest.mark.parametrize("tracing_from", [None, 37])
def test_c2_full_size_equals_keyed_oracle(gpu_engine_module, c2, tracing_from):
    """tracing_from = 37 is run_simulation.py:115-116 as the reference executes it as compiled."""
    w = c2
    n = w["n"]
    p = dict(w["p"])
    if tracing_from is not None:
        p["t_tracing_start"] = float(tracing_from)
    seed = 20200401
    home, init = tables_mod.prologue(seed, w["groups"], n, w["stay"], p["initial_infected_fraction"])
    assert len(init) == 5
    eng = _engine(gpu_engine_module, w, p)
    eng.run([seed], [home], [init])
    tal = eng.tallies(0)
    _properties(tal, n)
    _checksum_of_checksums(eng, tal, w["age"], (0, 1, 36, 37, 45, 46, T - 1))

    tup, ex = _oracle(w, p, seed, home, init)
    for k in util.VECS:
        got = eng.agent_vector(k)
        assert np.array_equal(got, tup[util.VEC_INDEX[k]], equal_nan=True), "vector %s differs" % k
    for t in (1, 20, 37, 38, 46, T - 1):
        want = util.pack_history([np.asarray(tup[k][t]) for k in range(9)])
        assert np.array_equal(eng.history("packed", t, t + 1)[0], want), "packed history of day %d differs" % t
    for k in range(9):
        assert np.array_equal(np.asarray(tup[k]).sum(axis=1), tal[:, k].sum(axis=1)), "daily totals of %s differ" % util.HIST[k]
    assert eng.counters()["infections"] == int((tup[15] > 0).sum()) > 100_000
    del tup, ex

    # idempotence: the same inputs again, and through the per-day launches
    eng.run([seed], [home], [init])
    assert np.array_equal(eng.tallies(0), tal)
    eng.close()
    eng = _engine(gpu_engine_module, w, p, launch="per_day")
    eng.run([seed], [home], [init])
    assert np.array_equal(eng.tallies(0), tal)
    eng.close()


def test_c2_full_size_replicas_are_independent(gpu_engine_module, c2):
    """Three seeds side by side in one launch == the three single runs (tallies and time_exposed), at full size."""
    w = c2
    n = w["n"]
    p = dict(w["p"])
    seeds = [11, 12, 13]
    pro = [tables_mod.prologue_fast(s, w["groups"], n, w["stay"], p["initial_infected_fraction"]) for s in seeds]
    eng = _engine(gpu_engine_module, w, p, n_replicas=3)
    eng.run(seeds, [q[0] for q in pro], [q[1] for q in pro])
    tal = [eng.tallies(r) for r in range(3)]
    te = [eng.agent_vector("time_exposed", replica=r) for r in range(3)]
    eng.close()
    assert not np.array_equal(tal[0], tal[1])
    eng = _engine(gpu_engine_module, w, p)
    for r, s in enumerate(seeds):
        eng.run([s], [pro[r][0]], [pro[r][1]])
        _properties(eng.tallies(0), n)
        assert np.array_equal(eng.tallies(0), tal[r]) and np.array_equal(eng.agent_vector("time_exposed"), te[r])
    eng.close()
