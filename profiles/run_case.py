"""One named workload, run a few times -- the unit ncu captures.   python profiles/run_case.py <case> [runs]

cases:  c2          C2 (10 M agents, T = 60, 5 initial cases, lockdown day 46)
        sparse      C2 cut at T = 30: every pass is a sparse one (<= 7 000 active agents)
        stress      SURVEY.md 8d stress variant: 0.1 % infected at t = 0, no lockdown
        c4like      10 M agents, T = 120, age-policy parameters (p_documented_in_mild = 0, asymptomatic isolation never)
Prints the device times of every run (CUDA events inside seir_run).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from covid19_demography_b200 import engine, population, tables


def build(case, n=10_000_000, replicas=1):
    T = {"sparse": 30, "c4like": 120}.get(case, 60)
    p, tabs = bench.workload_params(n, T)
    n_init = 5
    if case == "stress":
        p["initial_infected_fraction"] = 0.001
        p["t_lockdown"] = 1.0e6
        n_init = int(0.001 * n)
    if case == "c4like":  # simulate_agepolicies.py:269, :306
        p["mean_time_to_isolate_asympt"] = float("inf")
        p["p_documented_in_mild"] = 0.0
        p["t_lockdown"] = 1.0e6
        n_init = 500
    # POP=census: the census-driven population bench.py runs on (default here: the census-free generator of round 1,
    # so that the numbers compare with profiles/r01/)
    if os.environ.get("POP", "synthetic") == "census":
        hh, age, diab, hyp, groups = bench.make_population(n)
    else:
        hh, age, diab, hyp, groups = population.synthetic_population(n, seed=1)
    tb = tables.build_tables(tabs["contact_matrix"], np.array([len(g) for g in groups]), tables.params_to_dict(p),
                             tabs["mean_time_to_isolate_factor"], tabs["p_mild_severe"], tabs["p_severe_critical"],
                             tabs["p_critical_death"], with_replay=False)
    eng = engine.Engine(p, hh, age, groups, diab, hyp, np.full(n, tabs["p_infect_household"]), tb, n_replicas=replicas)
    return eng, n_init, T


if __name__ == "__main__":
    case = sys.argv[1]
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n = int(os.environ.get("N", "10000000"))
    eng, n_init, T = build(case, n)
    rng = np.random.default_rng(11)
    for k in range(runs):
        init = [rng.choice(n, size=n_init, replace=False).astype(np.int64)]
        eng.run([700 + k], None, init)
        c = eng.counters()
        print("%s run %d: init/loop/materialize ms = %.3f %.3f %.3f   active agent-days %d (visited %d), infections %d"
              % ((case, k) + eng.last_run_ms3() + (c["active_agent_days"], c["visits"], c["infections"])), flush=True)
    eng.close()
