"""Ensemble / sweep throughput on ONE B200 (the other GPUs of a box run their own share, no communication):
  C4-like: simulate_agepolicies.py semantics (T = 120, asymptomatic isolation never, no mild documentation,
           shelter-in-place of the 60+ from day 20) at n = 10 M, 8 seeds per launch;
  C5-like: parameter_sweep.py semantics (n = 1 M, T = 66, Italy defaults) at 64 seeds per launch.
    python profiles/c4_c5_ensembles.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from covid19_demography_b200 import engine, ensemble, population, tables


def fixture(name):
    fx = np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name))
    p = dict(zip(fx["param_keys"].tolist(), [float(v) for v in fx["param_vals"]]))
    return fx, p


def run(label, fx, p, n, R, batches, fsh):
    hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
    p = dict(p, n=float(n))
    tb = tables.build_tables(fx["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                             fx["mean_time_to_isolate_factor"], fx["p_mild_severe"], fx["p_severe_critical"],
                             fx["p_critical_death"], with_replay=False)
    eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, float(fx["p_infect_household_scalar"])), tb, n_replicas=R)
    T = int(p["T"])
    seeds = list(range(1000, 1000 + R * (batches + 1)))
    t_host0 = time.perf_counter()
    pro = [tables.prologue(s, groups, n, fsh, p["initial_infected_fraction"]) for s in seeds[:R]]
    t_pro = (time.perf_counter() - t_host0) / R
    dev = []
    wall0 = None
    for b in range(batches + 1):
        if b == 1:
            wall0 = time.perf_counter()
        eng.run(seeds[b * R:(b + 1) * R], [x[0] for x in pro], [x[1] for x in pro])
        for r in range(R):
            eng.tallies(r)
        dev.append(eng.last_run_ms()[1])
    wall = (time.perf_counter() - wall0) / batches
    tal = eng.tallies(0)
    ms = float(np.mean(dev[1:]))
    print("%s: n=%d T=%d, %d seeds per launch: %.2f ms device per launch (%.2f ms per run), %.2f ms wall per launch incl. tallies; "
          "%.3e agent-days/s, %.0f runs/hour per GPU (device), attack rate of seed 0 %.1f %%, host prologue %.0f ms per seed"
          % (label, n, T, R, ms, ms / R, wall * 1e3, n * (T - 1) * R / (ms / 1e3), 3600e3 * R / ms,
             100.0 * (1 - tal[-1, 0].sum() / n), t_pro * 1e3), flush=True)
    # end to end through ensemble.run_seeds: prologue picks on host threads overlapped with the device runs
    for fast in (False, True):
        t0 = time.perf_counter()
        res = ensemble.run_seeds(eng, list(range(5000, 5000 + R * batches)), groups, fsh, p["initial_infected_fraction"],
                                 fast_prologue=fast, host_threads=min(R, len(os.sched_getaffinity(0))))
        dt = time.perf_counter() - t0
        print("   end to end, %d runs, %s prologue on host threads: %.2f s -> %.0f runs/hour per GPU"
              % (R * batches, "fast" if fast else "reference-stream", dt, 3600.0 * R * batches / dt), flush=True)
    eng.close()


fx, p = fixture("italy_policy")
p.update(T=120.0, initial_infected_fraction=5.0 / 10_000_000, t_tracing_start=1.0e6)
run("C4-like", fx, p, 10_000_000, 8, 3, fx["fraction_stay_home"])
fx, p = fixture("italy_default")
p.update(T=66.0, initial_infected_fraction=5.0 / 1_000_000, t_tracing_start=1.0e6)
run("C5-like", fx, p, 1_000_000, 64, 3, fx["fraction_stay_home"])
