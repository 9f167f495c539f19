"""Freeze golden fixtures from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference (`/root/reference`) holds no tests or golden vectors, so the pins of this
repo are outputs of the reference itself: for each case below the reference's own
`run_model` (numba) is run on a reference-sampled population and all 20 outputs are
stored.  The nine bool[T,n] histories are stored packed, one byte per agent-day:
    bits 0-2 compartment (0=S 1=E 2=Mild 3=Severe 4=Critical 5=R 6=D), bit 3 = Q, bit 4 = Documented
(the compartments are one-hot in the reference: asserted here).
The fixtures travel to the GPU box; `/root/reference` and numba do not.
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")
HIST = ("S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q")
VEC = ("num_infected_by", "time_documented", "time_to_activation", "time_to_death", "time_to_recovery",
       "time_critical", "time_exposed", "num_infected_asympt", "age", "time_infected", "time_to_severe")


def pack_history(out):
    S, E, Mild, Doc, Sev, Crit, R, D, Q = [np.asarray(a) for a in out[:9]]
    onehot = S.astype(np.int8) + E + Mild + Sev + Crit + R + D
    assert (onehot == 1).all(), "compartments are not one-hot"
    comp = (E * 1 + Mild * 2 + Sev * 3 + Crit * 4 + R * 5 + D * 6).astype(np.uint8)
    return comp | (Q.astype(np.uint8) << 3) | (Doc.astype(np.uint8) << 4)


def unpack_history(packed):
    comp = packed & 7
    return tuple([comp == 0, comp == 1, comp == 2, (packed >> 4) & 1 == 1, comp == 3, comp == 4,
                  comp == 5, comp == 6, (packed >> 3) & 1 == 1])


CASES = {
    # name: (country, n, seed, param overrides, fraction_stay_home spec)
    "italy_default": ("Italy", 3000, 0, {}, None),
    "italy_seeded": ("Italy", 4000, 3, {"initial_infected_fraction": 0.005}, None),
    "china_default": ("China", 3000, 1, {"initial_infected_fraction": 0.002}, None),
    # simulate_agepolicies.py-like: T=120, asymptomatic isolation never (inf), no mild documentation,
    # shelter-in-place for the 60+ from day 20, tracing never starts
    "italy_policy": ("Italy", 3000, 2, {"T": 80.0, "mean_time_to_isolate_asympt": float("inf"),
                                        "p_documented_in_mild": 0.0, "t_stayhome_start": 20.0,
                                        "t_tracing_start": 1000.0, "initial_infected_fraction": 0.004,
                                        "t_lockdown": 30.0, "lockdown_factor": 2.0}, (60, 0.8)),
    # heavy tracing: many mild cases documented, tracing from day 10
    "italy_tracing": ("Italy", 3000, 4, {"p_documented_in_mild": 0.2, "t_tracing_start": 10.0,
                                         "initial_infected_fraction": 0.005, "p_trace_household": 0.7,
                                         "p_trace_outside": 0.5}, None),
    # short isolation delays: exercises the draw-is-zero -> Q paths on infection
    "italy_fastiso": ("Italy", 3000, 5, {"mean_time_to_isolate_asympt": 1.5, "mean_time_to_isolate": 1.0,
                                         "initial_infected_fraction": 0.01, "t_tracing_start": 1000.0}, None),
}


def run_case(name):
    country, n, seed, over, stay = CASES[name]
    ns = ref_shim.driver_namespace(n, country)
    ref = ref_shim.load()
    params = ns["params"]
    for k, v in over.items():
        params[k] = float(v)
    fsh = np.array(ns["fraction_stay_home"], dtype=np.float64)
    if stay is not None:
        fsh[stay[0]:] = stay[1]
    args = (ns["contact_matrix"], ns["p_mild_severe"], ns["p_severe_critical"], ns["p_critical_death"],
            ns["mean_time_to_isolate_factor"], ns["lockdown_factor_age"], ns["p_infect_household"], fsh)
    out = ref.run_complete_simulation(seed, country, *args, params, False)
    age, households, diabetes, hypertension, age_groups = pickle.load(
        open("%s_population_%d.pickle" % (country, n), "rb"))
    assert np.array_equal(age, out[17])
    fx = {
        "country": np.array(country), "seed": np.int64(seed),
        "param_keys": np.array(list(params.keys())), "param_vals": np.array([params[k] for k in params.keys()]),
        "households": households.astype(np.int32), "age": age.astype(np.int16),
        "diabetes": diabetes.astype(np.int8), "hypertension": hypertension.astype(np.int8),
        "contact_matrix": args[0], "p_mild_severe": args[1], "p_severe_critical": args[2],
        "p_critical_death": args[3], "mean_time_to_isolate_factor": args[4].astype(np.int64),
        "lockdown_factor_age": args[5], "p_infect_household_scalar": np.float64(args[6][0]),
        "fraction_stay_home": fsh, "history_packed": pack_history(out),
    }
    assert (args[6] == args[6][0]).all()
    for k, v in zip(HIST + VEC, out):
        if k in VEC and k != "age":
            fx["out_" + k] = np.asarray(v)
    path = os.path.join(GOLDEN, "ref_%s.npz" % name)
    np.savez_compressed(path, **fx)
    hp = fx["history_packed"]
    print(name, "n", n, "T", hp.shape[0], "final S/E/M/Sev/C/R/D",
          [int(((hp[-1] & 7) == c).sum()) for c in range(7)], "Q", int(((hp[-1] >> 3) & 1).sum()),
          "Doc", int(((hp[-1] >> 4) & 1).sum()), "%.0f kB" % (os.path.getsize(path) / 1e3))


def run_ensemble(n=20000, seeds=64, t_tracing_start=1000.0, name="ref_ensemble_italy_n%d.npz"):
    """Reference ensemble for the statistical-parity tier: one fixed
    Italy population, `seeds` simulation seeds; stores the daily curves.  With t_tracing_start < T the
    reference traces contacts from that day (as compiled)."""
    ns = ref_shim.driver_namespace(n, "Italy")
    ref = ref_shim.load()
    params = ns["params"]
    params["initial_infected_fraction"] = 0.001
    params["t_tracing_start"] = float(t_tracing_start)
    params["t_lockdown"] = 40.0
    args = (ns["contact_matrix"], ns["p_mild_severe"], ns["p_severe_critical"], ns["p_critical_death"],
            ns["mean_time_to_isolate_factor"], ns["lockdown_factor_age"], ns["p_infect_household"],
            np.array(ns["fraction_stay_home"], dtype=np.float64))
    out = ref.run_complete_simulation(1234, "Italy", *args, params, False)  # builds the population
    age, households, diabetes, hypertension, age_groups = pickle.load(open("Italy_population_%d.pickle" % n, "rb"))
    T = int(params["T"])
    curves = np.zeros((seeds, T, 9), dtype=np.int32)
    r0 = np.zeros(seeds)
    multi = np.zeros(seeds)
    for s in range(seeds):
        out = ref.run_model(s, households, age, age_groups, diabetes, hypertension, *args, params)
        for k in range(9):
            curves[s, :, k] = out[k].sum(axis=1)
        te, nib = out[15], out[9]
        m = (te > 0) & (te <= 20)
        r0[s] = nib[m].mean() if m.any() else np.nan
        multi[s] = nib[nib > 0].sum() / max(1, (te > 0).sum())
    fx = {
        "param_keys": np.array(list(params.keys())), "param_vals": np.array([params[k] for k in params.keys()]),
        "households": households.astype(np.int32), "age": age.astype(np.int16),
        "diabetes": diabetes.astype(np.int8), "hypertension": hypertension.astype(np.int8),
        "contact_matrix": args[0], "p_mild_severe": args[1], "p_severe_critical": args[2],
        "p_critical_death": args[3], "mean_time_to_isolate_factor": args[4].astype(np.int64),
        "lockdown_factor_age": args[5], "p_infect_household_scalar": np.float64(args[6][0]),
        "fraction_stay_home": args[7], "curves": curves, "r0_early": r0, "hits_per_infection": multi,
    }
    path = os.path.join(GOLDEN, name % n)
    np.savez_compressed(path, **fx)
    print("ensemble", name % n, curves[:, -1, :].mean(axis=0), "%.0f kB" % (os.path.getsize(path) / 1e3))


def run_ensemble_large(n=20000, seeds=256, name="ref_ensemble256_italy_n%d.npz"):
    """A larger reference ensemble on the SAME population and parameters as run_ensemble() (tracing never on), for
    the tightened statistical tier: daily curves of 256 seeds, per-age attack counts at the end, the pooled
    distribution of num_infected_by (the reference's per-infector offspring numbers).  Only statistics are stored;
    the population is the one of ref_ensemble_italy_n20000.npz (checked here)."""
    ns = ref_shim.driver_namespace(n, "Italy")
    ref = ref_shim.load()
    params = ns["params"]
    params["initial_infected_fraction"] = 0.001
    params["t_tracing_start"] = 1000.0
    params["t_lockdown"] = 40.0
    args = (ns["contact_matrix"], ns["p_mild_severe"], ns["p_severe_critical"], ns["p_critical_death"],
            ns["mean_time_to_isolate_factor"], ns["lockdown_factor_age"], ns["p_infect_household"],
            np.array(ns["fraction_stay_home"], dtype=np.float64))
    ref.run_complete_simulation(1234, "Italy", *args, params, False)  # builds the population
    age, households, diabetes, hypertension, age_groups = pickle.load(open("Italy_population_%d.pickle" % n, "rb"))
    small = np.load(os.path.join(GOLDEN, "ref_ensemble_italy_n%d.npz" % n))
    assert np.array_equal(small["age"], age) and np.array_equal(small["households"], households), "not the fixture's population"
    T = int(params["T"])
    curves = np.zeros((seeds, T, 9), dtype=np.int32)
    attack_by_age = np.zeros((seeds, 101), dtype=np.int32)
    offspring = np.zeros((seeds, 16), dtype=np.int64)   # histogram of num_infected_by over infected agents, last bin = 15+
    r0 = np.zeros(seeds)
    for s in range(seeds):
        out = ref.run_model(10000 + s, households, age, age_groups, diabetes, hypertension, *args, params)
        for k in range(9):
            curves[s, :, k] = out[k].sum(axis=1)
        te, nib = out[15], out[9]
        attack_by_age[s] = np.bincount(age[te >= 0], minlength=101)
        offspring[s] = np.bincount(np.minimum(nib[nib >= 0].astype(np.int64), 15), minlength=16)
        m = (te > 0) & (te <= 20)
        r0[s] = nib[m].mean() if m.any() else np.nan
    path = os.path.join(GOLDEN, name % n)
    np.savez_compressed(path, curves=curves, attack_by_age=attack_by_age, offspring=offspring, r0_early=r0,
                        seeds=np.arange(10000, 10000 + seeds))
    print("ensemble256", curves[:, -1, :].mean(axis=0), "%.0f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    which = sys.argv[1:] or (list(CASES) + ["ensemble", "ensemble_tracing"])
    for name in which:
        if name == "ensemble256":
            run_ensemble_large()
        elif name == "ensemble":
            run_ensemble()
        elif name == "ensemble_tracing":  # tracing from day 25 of 60, the drivers' p_trace_household = 1, p_trace_outside = 0.95
            run_ensemble(t_tracing_start=25.0, name="ref_ensemble_tracing_italy_n%d.npz")
        else:
            run_case(name)
