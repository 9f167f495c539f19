"""The drop-in surface: `seir_individual.run_model` / `run_complete_simulation` with the
reference's signatures, the 20-tuple of :392, and the statements the drivers apply to it
(run_simulation.py:448-460, parameter_sweep.py:556-567) executed verbatim on the returned objects."""
import os
import pickle
import sys

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _module():
    """The drop-in file, loaded under a private name: `sys.modules['seir_individual']` must stay free for the
    tests that import the reference's own module of that name (test_reference_live.py)."""
    import importlib.util
    name = "b200_dropin_seir_individual"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "covid19-demography_b200", "seir_individual.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def test_module_has_the_reference_entry_points():
    import inspect
    m = _module()
    sig = inspect.signature(m.run_model)
    assert list(sig.parameters)[:15] == ["seed", "households", "age", "age_groups", "diabetes", "hypertension",
                                         "contact_matrix", "p_mild_severe", "p_severe_critical", "p_critical_death",
                                         "mean_time_to_isolate_factor", "lockdown_factor_age", "p_infect_household",
                                         "fraction_stay_home", "params"]  # seir_individual.py:91
    sig = inspect.signature(m.run_complete_simulation)
    assert list(sig.parameters) == ["seed", "country", "contact_matrix", "p_mild_severe", "p_severe_critical",
                                    "p_critical_death", "mean_time_to_isolate_factor", "lockdown_factor_age",
                                    "p_infect_household", "fraction_stay_home", "params", "load_population"]  # :11


@pytest.mark.gpu
def test_driver_statements_on_the_returned_tuple(gpu_engine_module):
    m = _module()
    c = util.Case("italy_seeded")
    params = dict(c.params)  # tracing as compiled (t_tracing_start = 37)
    n, T, n_ages = c.n, int(params["T"]), 101
    out = m.run_model(c.seed, *c.model_args(params))
    assert len(out) == 20
    S, E, Mild, Documented, Severe, Critical, R, D, Q, num_infected_by, time_documented, \
        time_to_activation, time_to_death, time_to_recovery, time_critical, time_exposed, num_infected_asympt, \
        age, time_infected, time_to_severe = out
    tup, ex, _, _ = c.keyed_oracle("native", params)
    # ---- run_simulation.py:448-460, verbatim
    r_0_over_time, cfr_over_time = np.zeros(T), np.zeros(T)
    fraction_over_70_time, median_age_time = np.zeros(T), np.zeros(T)
    dead_by_age = np.zeros(n_ages)
    with np.errstate(invalid="ignore", divide="ignore"):
        for t in range(T):
            if (time_exposed == t).sum() > 0:
                r_0_over_time[t] = num_infected_by[time_exposed == t].mean()
                cfr_over_time[t] = D[-1, time_exposed == t].sum() / (D[-1, time_exposed == t].sum() + R[-1, time_exposed == t].sum())
                fraction_over_70_time[t] = (age[time_exposed == t] >= 70).mean()
                median_age_time[t] = np.median(age[time_exposed == t])
    total_infections_time = (params['n'] - S.sum(axis=1) - E.sum(axis=1))
    total_documented_time = Documented.sum(axis=1)
    for patient_age in range(n_ages):
        dead_by_age[patient_age] = D[-1, age == patient_age].sum()
    total_deaths_time = D.sum(axis=1)
    all_final_s = S[-1]
    # ---- parameter_sweep.py:556-567
    Q_per_time = Q.sum(axis=1)
    r0_total = num_infected_by[np.logical_and(time_exposed <= 20, time_exposed > 0)].mean()
    susceptible_70 = np.logical_and(S[T // 2], age >= 70).sum()
    # ---- the same from the oracle's plain arrays
    oS, oE, oDoc, oR, oD, oQ = (np.asarray(tup[k]) for k in (0, 1, 3, 6, 7, 8))
    te, nib = tup[15], tup[9]
    assert np.array_equal(total_infections_time, n - oS.sum(axis=1) - oE.sum(axis=1))
    assert np.array_equal(total_documented_time, oDoc.sum(axis=1)) and np.array_equal(total_deaths_time, oD.sum(axis=1))
    assert np.array_equal(Q_per_time, oQ.sum(axis=1)) and np.array_equal(all_final_s, oS[-1])
    assert np.array_equal(dead_by_age, np.array([oD[-1, c.age == a].sum() for a in range(n_ages)]))
    assert r0_total == nib[np.logical_and(te <= 20, te > 0)].mean()
    assert susceptible_70 == np.logical_and(oS[T // 2], c.age >= 70).sum()
    for t in range(T):
        msk = te == t
        if msk.sum() > 0:
            assert r_0_over_time[t] == nib[msk].mean() and median_age_time[t] == np.median(c.age[msk])
            assert fraction_over_70_time[t] == (c.age[msk] >= 70).mean()
    assert np.array_equal(np.asarray(Mild), np.asarray(tup[2])) and Mild.shape == (T, n) and len(Severe) == T
    assert np.array_equal(age, c.age) and np.array_equal(time_documented, tup[10])


@pytest.mark.gpu
def test_run_complete_simulation_keeps_the_population_pickle(gpu_engine_module, tmp_path, monkeypatch):
    """seir_individual.py:20-41: the population is written to / read from {country}_population_{n}.pickle in the CWD."""
    m = _module()
    c = util.Case("italy_default")
    monkeypatch.chdir(tmp_path)
    params = dict(c.no_tracing_params(), n=2000.0, initial_infected_fraction=0.01, T=20.0)
    args = (c.contact_matrix, c.p_mild_severe, c.p_severe_critical, c.p_critical_death, c.iso_factor,
            c.lockdown_factor_age, np.full(2000, c.p_infect_household[0]), c.fraction_stay_home, params)
    a = m.run_complete_simulation(7, "Italy", *args, load_population=False)
    assert os.path.exists("Italy_population_2000.pickle")
    age, households, diabetes, hypertension, age_groups = pickle.load(open("Italy_population_2000.pickle", "rb"))
    assert len(age) == 2000 and households.shape[0] == 2000 and len(age_groups) == 101
    b = m.run_complete_simulation(7, "Italy", *args, load_population=True)
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True)
    assert np.asarray(a[0]).shape == (20, 2000)


@pytest.mark.gpu
def test_histories_of_earlier_calls_stay_valid_and_the_engine_is_reused(gpu_engine_module):
    """The drivers call run_model once per seed with the same population arrays (run_simulation.py:437-444): the
    device-resident population is reused (no second engine), and the lazy histories of EARLIER calls -- the
    reference hands out fresh arrays every time -- keep their own run's content while later runs overwrite the
    device buffer: one run back on the device (seir_history_preserve), older ones from a host copy."""
    m = _module()
    m.clear_engine_cache()
    c = util.Case("italy_seeded")
    params = c.no_tracing_params()
    outs, oracles = [], []
    for seed in (1, 2, 3, 4):
        outs.append(m.run_model(seed, *c.model_args(params)))
        oracles.append(c.keyed_oracle("native", params, seed=seed)[0])
    assert len(m._engine_cache) == 1 and m.last_engine.run_serial == 4
    for out, tup in zip(outs, oracles):   # every call's nine histories and sums, read AFTER all four runs
        for k in range(9):
            assert np.array_equal(np.asarray(out[k]), np.asarray(tup[k])), k
            assert np.array_equal(out[k].sum(axis=1), np.asarray(tup[k]).sum(axis=1))
        assert np.array_equal(out[7][-1], np.asarray(tup[7])[-1]) and np.array_equal(out[15], tup[15])
        for name, k in util.VEC_INDEX.items():   # the ten lazy vectors: set aside on the device / host copies
            assert np.array_equal(np.asarray(out[k]), tup[k], equal_nan=True), name
    # a different population gets its own engine; the cache keeps the two most recent
    c2 = util.Case("china_default")
    m.run_model(5, *c2.model_args(c2.no_tracing_params()))
    assert len(m._engine_cache) == 2
    m.clear_engine_cache()


@pytest.mark.gpu
@pytest.mark.parametrize("vectors", ["lazy", "eager"])
def test_result_vectors_lazy_and_eager_with_tracing(gpu_engine_module, vectors):
    """The ten per-agent vectors of :392 as lazy AgentVectors (default) and as plain arrays (SEIR_B200_VECTORS=eager),
    with tracing as compiled (time_documented then also depends on the tracer's words, which are set aside too), read
    only after later runs have overwritten the device state."""
    from covid19_demography_b200.history import AgentVector
    m = _module()
    m.clear_engine_cache()
    c = util.Case("italy_tracing")
    params = dict(c.params)
    outs, oracles = [], []
    for seed in (4, 5, 6):
        outs.append(m.run_model(seed, *c.model_args(params), vectors=vectors))
        oracles.append(c.keyed_oracle("native", params, seed=seed)[0])
    for out, tup in zip(outs, oracles):
        for name, k in util.VEC_INDEX.items():
            assert isinstance(out[k], AgentVector if vectors == "lazy" else np.ndarray)
            assert np.array_equal(np.asarray(out[k]), tup[k], equal_nan=True), name
        assert (np.asarray(out[10]) > 0).sum() > 50   # documented cases: the tracer ran
    # a used vector is a host array from then on, whatever runs later
    te = outs[2][15]
    first = te[te >= 0].copy()
    m.run_model(9, *c.model_args(params), vectors=vectors)
    m.run_model(10, *c.model_args(params), vectors=vectors)
    assert np.array_equal(te[te >= 0], first)
    m.clear_engine_cache()


@pytest.mark.gpu
def test_native_prologue_gives_the_numpy_replays_run(gpu_engine_module):
    """run_model's default prologue (seir_prologue_mt) and the NumPy replay (prologue="numpy") pick the same agents:
    the two runs are the same run."""
    m = _module()
    m.clear_engine_cache()
    c = util.Case("italy_policy")   # shelter in place for the 60+: Home_real matters
    params = c.no_tracing_params()
    a = m.run_model(c.seed, *c.model_args(params), prologue="native")
    b = m.run_model(c.seed, *c.model_args(params), prologue="numpy")
    for k in range(20):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True), k
    home, init = m.last_engine.prologue_picks()
    h0, i0 = __import__("covid19_demography_b200.tables", fromlist=["prologue"]).prologue(
        c.seed, c.age_groups, c.n, c.fraction_stay_home, params["initial_infected_fraction"])
    assert np.array_equal(home, h0) and np.array_equal(np.sort(init), np.sort(i0))
    m.clear_engine_cache()


@pytest.mark.gpu
def test_results_outlive_an_evicted_engine(gpu_engine_module):
    """The engine cache closes the least recently used population; lazy histories and vectors of its last two runs that
    somebody still holds take host copies first (the reference's arrays never expire)."""
    m = _module()
    m.clear_engine_cache()
    c = util.Case("italy_seeded")
    params = c.no_tracing_params()
    a = m.run_model(1, *c.model_args(params))
    b = m.run_model(2, *c.model_args(params))
    m.clear_engine_cache()            # closes the handle
    for out, seed in ((a, 1), (b, 2)):
        tup = c.keyed_oracle("native", params, seed=seed)[0]
        for k in (0, 3, 7, 8):
            assert np.array_equal(np.asarray(out[k]), np.asarray(tup[k])), k
        assert np.array_equal(out[7][-1], np.asarray(tup[7])[-1])
        for name, k in util.VEC_INDEX.items():
            assert np.array_equal(np.asarray(out[k]), tup[k], equal_nan=True), name
