"""Shared helpers of the parity tests: fixtures -> oracle inputs -> oracle / GPU runs."""
import os

import numpy as np

import oracle  # oracle/oracle.py (test infrastructure)
from covid19_demography_b200 import tables as tables_mod

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("italy_default", "italy_seeded", "china_default", "italy_policy", "italy_tracing", "italy_fastiso")
HIST = ("S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q")
VECS = ("num_infected_by", "time_documented", "time_to_activation", "time_to_death", "time_to_recovery",
        "time_critical", "time_exposed", "num_infected_asympt", "time_infected", "time_to_severe")
VEC_INDEX = {"num_infected_by": 9, "time_documented": 10, "time_to_activation": 11, "time_to_death": 12,
             "time_to_recovery": 13, "time_critical": 14, "time_exposed": 15, "num_infected_asympt": 16,
             "time_infected": 18, "time_to_severe": 19}


def unpack_history(packed):
    comp = packed & 7
    return tuple([comp == 0, comp == 1, comp == 2, ((packed >> 4) & 1) == 1, comp == 3, comp == 4,
                  comp == 5, comp == 6, ((packed >> 3) & 1) == 1])


def pack_history(hist9):
    S, E, Mild, Doc, Sev, Crit, R, D, Q = [np.asarray(a).astype(np.uint8) for a in hist9]
    return (E * 1 + Mild * 2 + Sev * 3 + Crit * 4 + R * 5 + D * 6) | (Q << 3) | (Doc << 4)


class Case:
    """A golden fixture unpacked into the reference's run_model arguments."""

    def __init__(self, name):
        fx = np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
        self.name = name
        self.fx = fx
        self.params = dict(zip(fx["param_keys"].tolist(), [float(v) for v in fx["param_vals"]]))
        self.seed = int(fx["seed"]) if "seed" in fx else 0
        self.households = fx["households"].astype(np.int64)
        self.age = fx["age"].astype(np.int64)
        self.n = len(self.age)
        self.diabetes = fx["diabetes"].astype(np.int64)
        self.hypertension = fx["hypertension"].astype(np.int64)
        self.age_groups = tuple(np.where(self.age == a)[0] for a in range(101))
        self.contact_matrix = fx["contact_matrix"]
        self.p_mild_severe = fx["p_mild_severe"]
        self.p_severe_critical = fx["p_severe_critical"]
        self.p_critical_death = fx["p_critical_death"]
        self.iso_factor = fx["mean_time_to_isolate_factor"]
        self.lockdown_factor_age = fx["lockdown_factor_age"]
        self.p_infect_household = np.full(self.n, float(fx["p_infect_household_scalar"]))
        self.fraction_stay_home = fx["fraction_stay_home"]

    def model_args(self, params=None):
        return (self.households, self.age, self.age_groups, self.diabetes, self.hypertension, self.contact_matrix,
                self.p_mild_severe, self.p_severe_critical, self.p_critical_death, self.iso_factor,
                self.lockdown_factor_age, self.p_infect_household, self.fraction_stay_home,
                self.params if params is None else params)

    def no_tracing_params(self, **over):
        p = dict(self.params)
        p["t_tracing_start"] = 100000.0  # tier-1 parity semantics: tracing never switches on
        p.update(over)
        return p

    def tables(self, params):
        gs = np.array([len(g) for g in self.age_groups])
        return tables_mod.build_tables(self.contact_matrix, gs, tables_mod.params_to_dict(params), self.iso_factor,
                                       self.p_mild_severe, self.p_severe_critical, self.p_critical_death)

    def keyed_oracle(self, provider, params, seed=None, tracing_enabled=1):
        seed = self.seed if seed is None else seed
        home, init = tables_mod.prologue(seed, self.age_groups, self.n, self.fraction_stay_home,
                                         params["initial_infected_fraction"])
        tup, ex = oracle.run_model(seed, *self.model_args(params), provider=provider,
                                   tracing_enabled=tracing_enabled, home_real=home, initial_infected=init,
                                   tables=self.tables(params))
        return tup, ex, home, init

    def gpu_engine(self, engine_mod, params, mode, launch="persistent", n_replicas=1, tracing_enabled=1):
        return engine_mod.Engine(tables_mod.params_to_dict(params), self.households, self.age, self.age_groups,
                                 self.diabetes, self.hypertension, self.p_infect_household, self.tables(params),
                                 n_replicas=n_replicas, mode=mode, launch=launch, tracing_enabled=tracing_enabled)


def tallies_from_history(hist9, age, T):
    """[T, 9, 101] per-day per-age counts from nine bool[T, n] arrays (what the drivers' column sums and
    age masks compute, run_simulation.py:448-460, parameter_sweep.py:572-603)."""
    out = np.zeros((T, 9, 101), dtype=np.int64)
    for k, h in enumerate(hist9):
        h = np.asarray(h)
        for t in range(T):
            out[t, k] = np.bincount(age[h[t]], minlength=101)
    return out


def check_invariants(tup, n, T, tracing_off=True):
    """SURVEY.md section 4, tier T3."""
    S, E, Mild, Doc, Sev, Crit, R, D, Q = [np.asarray(a) for a in tup[:9]]
    tot = S.astype(np.int32) + E + Mild + Sev + Crit + R + D
    assert (tot == 1).all(), "compartments are not a partition"
    assert (S[1:] <= S[:-1]).all(), "S must be non-increasing"
    assert (R[1:] >= R[:-1]).all() and (D[1:] >= D[:-1]).all(), "R, D must be non-decreasing"
    assert (Q[Sev | Crit]).all(), "Severe/Critical implies Q"
    assert (Doc[1:] >= Doc[:-1]).all(), "Documented is monotone"
    assert not (Doc & S).any(), "an S agent is never documented"
    te, nib = tup[15], tup[9]
    assert np.array_equal(te == -1, S[-1]), "time_exposed == -1 <=> S at the end"
    assert np.array_equal(nib == -1, te == -1), "num_infected_by == -1 <=> never exposed"
    infections = int((te > 0).sum())
    assert nib[nib > 0].sum() >= infections, "sum(num_infected_by) >= infections (multi-hit double count)"
    if tracing_off:
        assert not (Q[-1] & (R[-1] | D[-1])).any(), "without tracing nobody recovered/dead is quarantined"
