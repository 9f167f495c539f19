"""Parameter sweeps in process: the reference's SLURM-array index arithmetic
(parameter_sweep.py:107-126, :546-548), the dealing of jobs to ranks, and -- on the GPU -- a two-combination
sweep on one resident population against the oracle."""
import numpy as np
import pytest

import util
from covid19_demography_b200 import sweep, tables as tables_mod


def test_index_arithmetic_matches_the_reference():
    # parameter_sweep.py: INDEX -> (combo, trial, seed); seed = INDEX * N_SIMS_PER_JOB + offset, +1 before each run
    items = sweep.work_items(n_combos=3, n_sims_per_combo=8, n_sims_per_job=4, seed_offset=100)
    assert len(items) == 6
    for idx, it in enumerate(items):
        jobs_per_combo = 8 // 4
        assert it["combo"] == idx // jobs_per_combo and it["trial"] == idx % jobs_per_combo
        seed = idx * 4 + 100
        want = []
        for _ in range(4):
            seed += 1
            want.append(seed)
        assert it["seeds"] == want
    all_seeds = sum((it["seeds"] for it in items), [])
    assert len(set(all_seeds)) == len(all_seeds)
    with pytest.raises(ValueError):
        sweep.work_items(2, 5, 2)


def test_gather_sweep_checks_coverage():
    a = {0: {"x": 1}, 2: {"x": 3}}
    with pytest.raises(RuntimeError, match="no rank"):
        sweep.gather_sweep(a, 3)
    assert sweep.gather_sweep({0: 1, 1: 2}, 2) == {0: 1, 1: 2}


class _FakeEngine:
    """Records the order of parameter switches and runs; results are a function of (combo marker, seed)."""
    n_replicas, n, T = 2, 10, 4

    def __init__(self):
        self.log, self.marker = [], None

    def set_params(self, params=None, tables=None):
        self.marker = params["p_infect_given_contact"]
        self.log.append(("set", self.marker))

    def run(self, seeds, home, init):
        self._seeds = list(seeds)
        self.log.append(("run", tuple(seeds)))

    def tallies(self, r):
        return np.full((self.T, 9, 101), int(self._seeds[r] * 1000 + self.marker * 100), dtype=np.int64)

    def cohorts(self, r):
        return np.zeros((self.T, 4, 101), dtype=np.int64)

    def counters(self, r):
        return {}


def test_run_sweep_switches_parameters_only_between_combinations():
    c = util.Case("italy_default")
    base = tables_mod.params_to_dict(c.params)
    combos = [{"params": dict(base, p_infect_given_contact=float(k)), "tables": None} for k in (1, 2)]
    items = sweep.work_items(2, 6, 3)
    groups = tuple(np.arange(0) for _ in range(101))
    seen = {}
    for rank in range(2):
        eng = _FakeEngine()
        res = sweep.run_sweep(eng, combos, items, groups, np.zeros(101), rank=rank, world=2)
        sets = [x for x in eng.log if x[0] == "set"]
        assert len(sets) == 2, "one parameter switch per combination on each rank"
        for idx, r in res.items():
            assert idx not in seen
            seen[idx] = r
            marker = combos[items[idx]["combo"]]["params"]["p_infect_given_contact"]
            assert [int(t[0, 0, 0]) for t in r["tallies"]] == [int(s * 1000 + marker * 100) for s in items[idx]["seeds"]]
    assert sorted(seen) == list(range(len(items)))


@pytest.mark.gpu
def test_two_combination_sweep_on_one_resident_population(gpu_engine_module):
    """parameter_sweep.py-like: transmissibility and the start day (hence T and the lockdown day) change between
    combinations, the population stays on the device; every run == the oracle under the same keyed draws."""
    c = util.Case("italy_default")
    base = c.no_tracing_params()
    combos = []
    for pigc, T, t_lock in ((0.035, 50.0, 30.0), (0.025, 66.0, 46.0)):
        p = dict(base, p_infect_given_contact=pigc, T=T, t_lockdown=t_lock, initial_infected_fraction=0.004)
        combos.append({"params": p, "tables": c.tables(p)})
    items = sweep.work_items(2, 4, 2, seed_offset=40)
    eng = c.gpu_engine(gpu_engine_module, combos[0]["params"], "native", n_replicas=2)
    res = sweep.gather_sweep(sweep.run_sweep(eng, combos, items, c.age_groups, c.fraction_stay_home), len(items))
    for idx, r in res.items():
        p = combos[items[idx]["combo"]]["params"]
        for k, seed in enumerate(r["seeds"]):
            tup, _, _, _ = c.keyed_oracle("native", p, seed=seed)
            assert np.array_equal(r["tallies"][k], util.tallies_from_history(tup[:9], c.age, int(p["T"]))), (idx, seed)
    eng.close()


def test_job_files_have_the_reference_names_shapes_and_values(tmp_path):
    """parameter_sweep.py:617-733: 19 tab-separated files per job, read back the way combine_results.py does."""
    import pandas as pd
    import test_driver_tallies as tdt
    from covid19_demography_b200 import sweep_io
    c = util.Case("italy_seeded")
    params = c.no_tracing_params()
    T = int(params["T"])
    runs = [c.keyed_oracle("native", params, seed=s)[0] for s in (61, 62)]
    tallies = np.stack([util.tallies_from_history(tup[:9], c.age, T) for tup in runs])
    cohorts = np.stack([tdt.cohorts_from_outputs(tup, c.age, T) for tup in runs])
    obs = np.arange(5, dtype=np.float64)
    arrays = sweep_io.job_arrays(tallies, cohorts, actual_deaths=obs, first_day=40)
    stem = sweep_io.file_stem("sweepA", 10000000.0, 7, 1, 0.029, 4, 22)
    assert stem == "sweepA_paramsweep_n10000000.0_i7_N1_p0.029_m4_s22"
    paths = sweep_io.write_job_files(str(tmp_path / "sweepA"), stem, arrays)
    assert len(paths) == 19 and set(paths) == set(sweep_io.DATATYPES)
    want = [tdt.drivers_way(tup, c.age, c.n, T) for tup in runs]
    back = {k: pd.read_csv(p, sep="\t", header=None).values for k, p in paths.items()}
    assert back["susceptible"].shape == (2, T) and back["infected_age_time"].shape == (2, 5 * T) and back["r0_tot"].shape == (2, 1)
    for i, (tup, w) in enumerate(zip(runs, want)):
        assert np.array_equal(back["deaths"][i], np.asarray(tup[7]).sum(axis=1))
        assert np.array_equal(back["quarantine"][i], np.asarray(tup[8]).sum(axis=1))
        assert np.allclose(back["r0_time"][i], w["r_0_over_time"], equal_nan=True)
        assert np.allclose(back["cfr_time"][i], w["cfr_over_time"], equal_nan=True)
        assert np.allclose(back["median_age_time"][i], w["median_age_time"], equal_nan=True)
        assert np.allclose(back["infected_age_time"][i].reshape(5, T), w["infected_by_age_by_time"])
        assert np.allclose(back["cfr_age_time"][i].reshape(5, T), w["CFR_by_age_by_time"], equal_nan=True)
        assert np.allclose(back["dead_age_time"][i].reshape(5, T), w["dead_by_age_by_time"])
        assert back["r0_tot"][i, 0] == pytest.approx(w["r0_total"])
        d = np.asarray(tup[7]).sum(axis=1)[40:45]
        assert back["mse"][i, 0] == pytest.approx(np.mean((d - obs) ** 2))
    assert "NA" in open(paths["cfr_age_time"]).read()  # empty cohorts: 0/0, written as NA like the reference


def test_agepolicy_job_files(tmp_path):
    """simulate_agepolicies.py:645-716: twelve files per job."""
    import pandas as pd
    import test_driver_tallies as tdt
    from covid19_demography_b200 import sweep_io
    c = util.Case("italy_policy")
    params = c.no_tracing_params()
    T = int(params["T"])
    tup = c.keyed_oracle("native", params, seed=70)[0]
    tallies = util.tallies_from_history(tup[:9], c.age, T)[None]
    cohorts = tdt.cohorts_from_outputs(tup, c.age, T)[None]
    arrays = sweep_io.agepolicy_job_arrays(tallies, cohorts, c.n)
    stem = sweep_io.agepolicy_file_stem("pol", c.n, 3, 0, 0.029, 4, 22)
    assert stem == "pol_agepolicy_n3000.0_i3_N0_p0.029_m4_s22"
    paths = sweep_io.write_job_files(str(tmp_path), stem, arrays, datatypes=sweep_io.AGEPOLICY_DATATYPES)
    assert len(paths) == 12
    S, E, Doc = (np.asarray(tup[k]) for k in (0, 1, 3))
    back = {k: pd.read_csv(p, sep="\t", header=None).values for k, p in paths.items()}
    assert np.array_equal(back["infected"][0], c.n - S.sum(axis=1))  # :617
    assert np.array_equal(back["documented"][0], Doc.sum(axis=1))
    assert np.array_equal(back["susceptible"][0], S.sum(axis=1))


def test_tsv_writer_is_byte_identical_to_the_reference_to_csv_call():
    """parameter_sweep.py:716-729 writes every per-job array with pandas' to_csv(sep='\\t', index=False, header=False,
    na_rep='NA'); sweep_io formats the same bytes itself (pandas' per-call overhead dominated a sweep's wall time)."""
    import io

    import pandas as pd
    from covid19_demography_b200 import sweep_io

    def reference_bytes(a):
        b = io.StringIO()
        pd.DataFrame(a).to_csv(b, sep="\t", index=False, header=False, na_rep="NA")
        return b.getvalue()

    rng = np.random.default_rng(0)
    cases = [rng.random((3, 7)), rng.random(5), rng.integers(0, 100, (2, 5)), (rng.random((2, 365)) * 1e6).round(),
             np.array([1.0, np.nan, np.inf, -np.inf, 1e-5, 1e22, 123456789.123, 0.0, -0.0, 3.0, 1 / 3]),
             np.array([[np.nan] * 3] * 2), rng.random((4, 10)) / 3e-7, np.full(4, np.nan)]
    for a in cases:
        assert sweep_io._format_tsv(a) == reference_bytes(a)
        if np.asarray(a).dtype.kind == "f":
            assert sweep_io._format_tsv_py(a) == reference_bytes(a)


def test_native_tsv_formatter_equals_pythons_repr_on_random_doubles():
    """seir_format_tsv (host routine of the library: shortest round-trip digits from std::to_chars, laid out by Python's
    repr rules) against Python itself on every kind of double: random bit patterns (subnormals, huge exponents, NaN
    payloads), powers of ten around the fixed / scientific switch-over, counts, ratios."""
    from covid19_demography_b200 import engine, sweep_io
    rng = np.random.default_rng(5)
    special = np.array([0.0, -0.0, 1e-5, 1e-4, 9.999e-5, 1e15, 1e16, 1e17, 5e-324, 1.7976931348623157e308, np.inf, -np.inf,
                        np.nan, 0.30000000000000004, 1e22, 1e23, 9007199254740993.0, 123456789012345678.0, 1e100, 1e-100])
    pieces = [special, rng.integers(0, 2**64, size=200000, dtype=np.uint64).view(np.float64),
              rng.integers(0, 10**7, 50000).astype(np.float64), rng.random(50000),
              rng.integers(0, 5000, 50000) / rng.integers(1, 300, 50000),
              10.0 ** rng.integers(-30, 30, 50000) * rng.random(50000),
              np.concatenate([10.0 ** np.arange(-320, 309.0), -(10.0 ** np.arange(-20, 20.0))])]
    for a in pieces:
        a = a[: len(a) // 5 * 5].reshape(-1, 5) if len(a) >= 5 else a
        assert engine.format_tsv(a).decode("ascii") == sweep_io._format_tsv_py(a)
    assert engine.format_tsv(np.zeros((0, 3))) == b"" and engine.format_tsv(np.zeros(0)) == b""
