"""Build-container only: the oracle against the reference run LIVE (numba), beyond the frozen cases."""
import pickle

import numpy as np
import pytest

import oracle
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present on this box")


def _numba_ok():
    try:
        import numba  # noqa: F401
        return True
    except Exception:
        return False


@pytest.mark.skipif(not _numba_ok(), reason="numba missing")
def test_oracle_equals_live_reference_new_seed():
    n = 2500
    ns = ref_shim.driver_namespace(n, "Italy")
    ref = ref_shim.load()
    params = ns["params"]
    params["initial_infected_fraction"] = 0.004
    args = (ns["contact_matrix"], ns["p_mild_severe"], ns["p_severe_critical"], ns["p_critical_death"],
            ns["mean_time_to_isolate_factor"], ns["lockdown_factor_age"], ns["p_infect_household"],
            ns["fraction_stay_home"])
    out = ref.run_complete_simulation(11, "Italy", *args, params, False)
    age, households, diabetes, hypertension, age_groups = pickle.load(open("Italy_population_%d.pickle" % n, "rb"))
    tup, _ = oracle.run_model(11, households, age, age_groups, diabetes, hypertension, *args, dict(params),
                              provider="mt", tracing_enabled=1)
    for k, a, b in zip(oracle.OUTPUT_NAMES, out, tup):
        assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), k
    # params['contact_tracing'] = 1 gives the same run (tracing is on as compiled)
    params["contact_tracing"] = 1.0
    out2 = ref.run_model(11, households, age, age_groups, diabetes, hypertension, *args, params)
    for k, a, b in zip(oracle.OUTPUT_NAMES, out, out2):
        assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), k


def test_dropin_population_can_be_the_references_own_sampler(monkeypatch):
    """BASELINE.json north_star: household / comorbidity sampling stay Python host code.  With the reference
    repository on sys.path and as CWD (the drop-in scenario) `SEIR_B200_POPULATION=reference` calls the reference's own
    sampler, so the same NumPy seed gives the same population; the default is the vectorised restatement."""
    from covid19_demography_b200 import population
    ref_shim.load()  # puts /root/reference on sys.path, chdir to a scratch dir with the data symlinks, numpy aliases
    import sample_households as ref_sh
    import sample_comorbidities as ref_sc
    np.random.seed(21)
    hh_ref, age_ref = ref_sh.sample_households_italy(700)
    d_ref, h_ref = ref_sc.sample_joint_comorbidities(age_ref, "Italy")
    monkeypatch.setenv("SEIR_B200_POPULATION", "reference")
    np.random.seed(21)
    hh, age = population.sample_households("Italy", 700)
    d, h = population.sample_comorbidities(age, "Italy")
    assert np.array_equal(hh, hh_ref) and np.array_equal(age, age_ref)
    assert np.array_equal(d, d_ref) and np.array_equal(h, h_ref)
    for src in ("census", "synthetic"):
        monkeypatch.setenv("SEIR_B200_POPULATION", src)
        hh2, _ = population.sample_households("Italy", 700)
        assert hh2.shape == (700, 6) and not np.array_equal(hh2, hh_ref)
