"""The sweep scheduler: workers pulling jobs through claim files, jobs of different
combinations batched into one launch, resumable by output-file existence (parameter_sweep.py:82-126,
job.parameter_sweep.sh:15, combine_results.py:91-92)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

import util
from covid19_demography_b200 import scheduler, sweep_io, tables as tables_mod


class FakeEngine:
    """Replica results are a function of (seed, table set, T): enough to check who ran what with which overrides."""
    n_replicas, n, T = 4, 50, 12

    def __init__(self):
        self.launches = []

    def set_table_bank(self, sets, pcs):
        self.n_sets = len(sets)

    def run(self, seeds, home=None, init=None, keyed=None, overrides=None):
        assert len(seeds) == self.n_replicas and all(o["table_set"] < self.n_sets for o in overrides)
        self._runs = list(zip(seeds, overrides))
        self.launches.append(self._runs)

    def tallies(self, r):
        seed, o = self._runs[r]
        t = np.zeros((self.T, 9, 101), dtype=np.int64)
        t[:o["T"], 0, 0] = self.n
        t[:o["T"], 7, 1] = seed * 10 + o["table_set"]
        return t[:o["T"]]

    def cohorts(self, r):
        return np.zeros((self.T, 4, 101), dtype=np.int64)

    def last_run_ms3(self):
        return (0.0, 0.0, 0.0)


def _plan(n_combos=5, per_combo=4, per_job=2):
    c = util.Case("italy_default")
    base = tables_mod.params_to_dict(c.params)
    combos = [{"params": dict(base, n=50.0, T=float(12 - (k % 3)), p_infect_given_contact=0.02 + 0.01 * k),
               "tables": {}, "label": (0.02 + 0.01 * k, 4, 20 + k)} for k in range(n_combos)]
    return scheduler.SweepPlan("sw", combos, per_combo, per_job, seed_offset=7)


def _deaths_file(directory, plan, job):
    return np.loadtxt(os.path.join(directory, plan.stem(job) + "_deaths.csv"), delimiter="\t", ndmin=2)


def test_jobs_of_different_combinations_share_a_launch_and_every_file_is_written(tmp_path):
    plan = _plan(per_combo=6)   # three jobs per combination, two jobs per launch: launches straddle combinations
    eng = FakeEngine()
    ran = scheduler.run_worker(eng, plan, str(tmp_path))
    assert ran == len(plan.items) == 15 and scheduler.missing_jobs(plan, str(tmp_path)) == []
    assert len(eng.launches) == 8 and any(len({o["table_set"] for _, o in l}) > 1 for l in eng.launches)
    for job in plan.items:
        d = _deaths_file(str(tmp_path), plan, job)
        T = int(plan.combos[job["combo"]]["params"]["T"])
        assert d.shape == (2, T) and d[:, -1].tolist() == [s * 10 + job["combo"] for s in job["seeds"]]
        for dt in sweep_io.DATATYPES:
            assert os.path.exists(os.path.join(str(tmp_path), "%s_%s.csv" % (plan.stem(job), dt)))
    assert not [f for f in os.listdir(str(tmp_path)) if f.endswith(".claim") or f.startswith(".tmp")]


def test_a_killed_sweep_resumes_without_recomputation(tmp_path):
    plan = _plan()
    first = FakeEngine()
    assert scheduler.run_worker(first, plan, str(tmp_path), max_jobs=4) == 4
    left = scheduler.missing_jobs(plan, str(tmp_path))
    assert len(left) == 6
    # a worker that died mid-job leaves a claim behind: its process is gone, the job is taken over
    stale = plan.stem(plan.items[left[0]])
    with open(os.path.join(str(tmp_path), stale + ".claim"), "w") as f:
        f.write("%s:%d" % (scheduler.socket.gethostname(), 2**22 + 12345))
    stamp = {f: os.path.getmtime(os.path.join(str(tmp_path), f)) for f in os.listdir(str(tmp_path)) if f.endswith(".csv")}
    second = FakeEngine()
    assert scheduler.run_worker(second, plan, str(tmp_path)) == 6
    ran_twice = {s for l in first.launches for s, _ in l[:4]} & {s for l in second.launches for s, _ in l[:4]}
    assert not ran_twice or len(second.launches[-1]) == 4   # (only the padding of a short last launch repeats a seed)
    for f, t in stamp.items():
        assert os.path.getmtime(os.path.join(str(tmp_path), f)) == t, "a finished job was rewritten"
    assert scheduler.missing_jobs(plan, str(tmp_path)) == []


def _worker(directory, out):
    eng = FakeEngine()
    out.put((os.getpid(), scheduler.run_worker(eng, _plan(n_combos=9, per_combo=8), directory),
             [s for l in eng.launches for s, _ in l]))


def test_concurrent_workers_split_the_queue(tmp_path):
    ctx = mp.get_context("fork")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(str(tmp_path), out)) for _ in range(3)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    plan = _plan(n_combos=9, per_combo=8)
    assert sum(r[1] for r in res) == len(plan.items) == 36
    assert scheduler.missing_jobs(plan, str(tmp_path)) == []
    for job in plan.items:
        assert _deaths_file(str(tmp_path), plan, job)[:, -1].tolist() == [s * 10 + job["combo"] for s in job["seeds"]]


@pytest.mark.gpu
def test_scheduler_on_the_device_equals_the_oracle(gpu_engine_module, tmp_path):
    """Four grid points x two jobs on one resident population, jobs of different grid points in one launch, picks of
    the prologue made by the device: the files hold the oracle's numbers."""
    import pandas as pd
    c = util.Case("italy_policy")
    gs = np.array([len(g) for g in c.age_groups])
    combos = []
    for k, (pigc, T, t_lock) in enumerate(((0.035, 50.0, 30.0), (0.025, 66.0, 46.0), (0.03, 58.0, 20.0), (0.04, 44.0, 25.0))):
        p = dict(c.no_tracing_params(), p_infect_given_contact=pigc, T=T, t_lockdown=t_lock, initial_infected_fraction=0.004,
                 t_stayhome_start=t_lock)
        combos.append({"params": p, "tables": tables_mod.build_tables(c.contact_matrix, gs, tables_mod.params_to_dict(p), c.iso_factor,
                                                                      c.p_mild_severe, c.p_severe_critical, c.p_critical_death),
                       "label": (pigc, 4, int(T))})
    plan = scheduler.SweepPlan("dev", combos, 4, 2, seed_offset=300, fraction_stay_home=c.fraction_stay_home)
    base = dict(combos[0]["params"], T=float(plan.T_max))
    eng = gpu_engine_module.Engine(tables_mod.params_to_dict(base), c.households, c.age, c.age_groups, c.diabetes, c.hypertension,
                                   c.p_infect_household, combos[0]["tables"], n_replicas=6, history="on_demand")
    assert scheduler.run_worker(eng, plan, str(tmp_path)) == 8
    eng.close()
    import oracle
    for job in plan.items:
        p = combos[job["combo"]]["params"]
        deaths = pd.read_csv(os.path.join(str(tmp_path), plan.stem(job) + "_deaths.csv"), sep="\t", header=None).values
        for k, seed in enumerate(job["seeds"]):
            home, init = tables_mod.prologue_keyed(seed, c.age_groups, c.n, c.fraction_stay_home, p["initial_infected_fraction"])
            tup, _ = oracle.run_model(seed, *c.model_args(p), provider="native", tracing_enabled=1, home_real=home,
                                      initial_infected=init, tables=combos[job["combo"]]["tables"])
            assert np.array_equal(deaths[k], np.asarray(tup[7]).sum(axis=1)), (job["index"], seed)
