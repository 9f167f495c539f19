"""-m gpu : the round-2 additions to the C ABI -- device-side keyed prologue, per-replica overrides with a table
bank (one launch over several grid points of a sweep), on-demand history, histories that survive later runs -- each
against the oracle / against the plain single-run path, bit-exact."""
import numpy as np
import pytest

import util
from covid19_demography_b200 import tables as tables_mod
from test_gpu_parity import _compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["italy_policy", "italy_seeded", "china_default"])
def test_keyed_prologue_on_device_equals_its_host_restatement(gpu_engine_module, name):
    """Home_real and initial_infected (seir_individual.py:176-187) drawn BY THE DEVICE (sd_perm, include/seir_draws.h)
    are the picks tables.prologue_keyed computes on the host, and the run that follows is the oracle's run on them."""
    import oracle
    c = util.Case(name)
    params = c.no_tracing_params(t_stayhome_start=5.0, initial_infected_fraction=0.01)
    fsh = np.linspace(0.0, 0.9, 101)
    seeds = [3, 4, 2**40 + 17]
    eng = c.gpu_engine(gpu_engine_module, params, "native", n_replicas=len(seeds))
    eng.run(seeds, keyed=(fsh, params["initial_infected_fraction"]))
    for r, seed in enumerate(seeds):
        home, init = tables_mod.prologue_keyed(seed, c.age_groups, c.n, fsh, params["initial_infected_fraction"])
        for a, g in enumerate(c.age_groups):   # the reference's counts (:181, :186)
            assert home[g].sum() == int(fsh[a] * len(g))
        assert len(init) == int(params["initial_infected_fraction"] * c.n) == len(set(init.tolist()))
        home_d, init_d = eng.prologue_picks(r)
        assert np.array_equal(home_d, home) and np.array_equal(init_d, init)
        tup, ex = oracle.run_model(seed, *c.model_args(params), provider="native", tracing_enabled=1, home_real=home,
                                   initial_infected=init, tables=c.tables(params))
        _compare(eng, tup, ex, c, int(params["T"]), replica=r)
    # mixed: replica 0 with host picks, the others keyed
    home0, init0 = tables_mod.prologue(9, c.age_groups, c.n, fsh, 0.02)
    eng.run([9, 10, 11], [home0, None, None], [init0, None, None], keyed=[None, (fsh, 0.01), (None, 0.005)])
    h, i = eng.prologue_picks(0)
    assert np.array_equal(h, home0) and sorted(i.tolist()) == sorted(init0.tolist())
    h, i = eng.prologue_picks(2)
    assert not h.any() and len(i) == int(0.005 * c.n)
    tup, ex = oracle.run_model(9, *c.model_args(params), provider="native", tracing_enabled=1, home_real=home0,
                               initial_infected=init0, tables=c.tables(params))
    _compare(eng, tup, ex, c, int(params["T"]), replica=0)
    eng.close()


def _grid(c, k):
    """Grid point k of a small sweep in the manner of parameter_sweep.py:54-56: transmissibility, mortality
    multiplier (severity tables), start date (T and the switch days)."""
    p = c.no_tracing_params()
    p["p_infect_given_contact"] = [0.02, 0.03, 0.045, 0.06][k % 4]
    mult = [1.0, 2.5][(k // 4) % 2]
    shift = [0, 7][(k // 2) % 2]
    p["T"] = float(int(p["T"]) - shift)
    p["t_lockdown"] = float(max(1, int(p["t_lockdown"]) - shift))
    p["t_stayhome_start"] = float(max(1, 12 - shift))
    sev = [np.clip(x * mult ** (1.0 / 3.0), 0.0, 1.0) for x in (c.p_mild_severe, c.p_severe_critical, c.p_critical_death)]
    return p, sev


@pytest.mark.parametrize("launch", ["persistent", "per_day"])
def test_replicas_of_different_grid_points_equal_their_single_runs(gpu_engine_module, launch):
    """64 replicas over 8 parameter sets in ONE launch (seir_run_ex: per-replica T, switch days and table set) ==
    the 64 single runs, bit for bit."""
    import oracle
    c = util.Case("italy_policy")
    gs = np.array([len(g) for g in c.age_groups])
    sets, points = [], []
    for k in range(8):
        p, sev = _grid(c, k)
        tb = tables_mod.build_tables(c.contact_matrix, gs, tables_mod.params_to_dict(p), c.iso_factor, *sev)
        sets.append(tb)
        points.append((p, sev, tb))
    R = 64
    base = dict(points[0][0])
    base["T"] = max(pt[0]["T"] for pt in points)
    eng = gpu_engine_module.Engine(tables_mod.params_to_dict(base), c.households, c.age, c.age_groups, c.diabetes,
                                   c.hypertension, c.p_infect_household, sets[0], n_replicas=R, launch=launch)
    eng.set_table_bank(sets, [pt[0]["p_infect_given_contact"] for pt in points])
    seeds = [100 + r for r in range(R)]
    combo = [r % 8 for r in range(R)]
    picks = [tables_mod.prologue_keyed(s, c.age_groups, c.n, c.fraction_stay_home, base["initial_infected_fraction"]) for s in seeds]
    ov = [{"T": int(points[k][0]["T"]), "t_lockdown": int(points[k][0]["t_lockdown"]),
           "t_stayhome_start": int(points[k][0]["t_stayhome_start"]), "table_set": k} for k in combo]
    eng.run(seeds, keyed=(c.fraction_stay_home, base["initial_infected_fraction"]), overrides=ov)
    for r in range(0, R, 5):   # (every fifth replica against the oracle: all eight grid points are covered)
        p, sev, tb = points[combo[r]]
        args = list(c.model_args(p))
        args[6:9] = sev
        tup, ex = oracle.run_model(seeds[r], *args, provider="native", tracing_enabled=1, home_real=picks[r][0],
                                   initial_infected=picks[r][1], tables=tb)
        _compare(eng, tup, ex, c, int(p["T"]), replica=r)
    # and every replica against a plain single-replica engine run with seir_set_params per grid point
    one = gpu_engine_module.Engine(tables_mod.params_to_dict(base), c.households, c.age, c.age_groups, c.diabetes,
                                   c.hypertension, c.p_infect_household, sets[0], n_replicas=1, launch=launch)
    for r in range(R):
        p, sev, tb = points[combo[r]]
        one.set_params(p, tb)
        one.run([seeds[r]], [picks[r][0]], [picks[r][1]])
        assert np.array_equal(one.tallies(0), eng.tallies(r)), r
        assert np.array_equal(one.history("packed"), eng.history("packed", replica=r)), r
        assert np.array_equal(one.agent_vector("num_infected_by"), eng.agent_vector("num_infected_by", replica=r)), r
    one.close()
    eng.close()


def test_on_demand_history_equals_eager(gpu_engine_module):
    """SEIR_HISTORY_ON_DEMAND keeps no T x n buffer: the rows asked for are computed from the records."""
    c = util.Case("italy_tracing")
    params = dict(c.params)   # tracing as compiled: post-end quarantine marks included
    tup, ex, home, init = c.keyed_oracle("native", params)
    T = int(params["T"])
    eng = gpu_engine_module.Engine(tables_mod.params_to_dict(params), c.households, c.age, c.age_groups, c.diabetes,
                                   c.hypertension, c.p_infect_household, c.tables(params), n_replicas=2, history="on_demand")
    eng.run([c.seed, c.seed + 1], [home, home], [init, init])
    assert eng.counters()["launches"] >= 1
    want = util.pack_history(tup[:9])
    assert np.array_equal(eng.history("packed"), want)
    assert np.array_equal(eng.history("packed", 7, 23), want[7:23])
    assert np.array_equal(eng.history("Q", T - 1, T), np.asarray(tup[8])[T - 1:T])
    assert np.array_equal(eng.tallies(0), util.tallies_from_history(tup[:9], c.age, T))
    eng.close()
