"""ctypes front end of the CPU oracle (oracle/oracle_seir.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product never imports it.

`run_model` takes the reference's argument list (`seir_individual.py:91`) and
returns the reference's 20-tuple (`:392`) plus a dict of extras.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_seir.so")

MT, KEYED_REPLAY, KEYED_NATIVE = 0, 1, 2
PROVIDERS = {"mt": MT, "replay": KEYED_REPLAY, "native": KEYED_NATIVE}


def build(force=False):
    src = os.path.join(HERE, "oracle_seir.c")
    hdr = os.path.join(HERE, "..", "include", "seir_draws.h")
    stale = (not os.path.exists(LIB_PATH)
             or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle_seir.so"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


class Params(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n", "T", "n_ages", "W", "t_lockdown", "t_tracing_start", "t_stayhome_start",
        "tracing_enabled", "provider", "n_iso_rows", "n_initial")] + [("seed", C.c_uint64)] + [
        (k, C.c_double) for k in (
            "lockdown_factor", "p_trace_household", "p_trace_outside", "p_documented_in_mild",
            "p_infect_given_contact", "asymptomatic_transmissibility",
            "mean_time_to_isolate", "mean_time_to_isolate_asympt", "mean_time_to_severe",
            "mean_time_mild_recovery", "mean_time_to_critical", "mean_time_severe_recovery",
            "mean_time_critical_recovery", "mean_time_to_death",
            "time_to_activation_mean", "time_to_activation_std", "initial_infected_fraction")]


_IN_FIELDS = ["households", "age", "diabetes", "hypertension", "group_offsets", "group_members",
              "contact_matrix", "p_mild_severe", "p_severe_critical", "p_critical_death", "iso_factor",
              "p_infect_household", "fraction_stay_home", "home_real", "initial_infected",
              "enlam_pre", "enlam_post", "nat_enlam", "nat_alias_prob", "nat_alias_idx"]
_OUT_FIELDS = ["S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q",
               "num_infected_by", "time_documented", "time_to_activation", "time_to_death",
               "time_to_recovery", "time_critical", "time_exposed", "num_infected_asympt",
               "time_infected", "time_to_severe", "time_to_isolate", "time_severe", "time_to_critical",
               "num_infected_by_outside", "home_real_out", "initial_infected_out", "traced", "draw_count"]


class Inputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _IN_FIELDS]


class Outputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _OUT_FIELDS]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_run_model.argtypes = [C.POINTER(Params), C.POINTER(Inputs), C.POINTER(Outputs)]
        _lib.oracle_run_model.restype = C.c_int
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def age_groups_csr(age_groups):
    """Flatten the reference's tuple of 101 index arrays (seir_individual.py:36) into CSR."""
    lens = np.array([len(g) for g in age_groups], dtype=np.int64)
    offsets = np.zeros(len(age_groups) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    members = (np.concatenate([np.asarray(g, dtype=np.int64) for g in age_groups])
               if offsets[-1] > 0 else np.zeros(0, np.int64))
    return offsets, np.ascontiguousarray(members, dtype=np.int64)


def run_model(seed, households, age, age_groups, diabetes, hypertension, contact_matrix, p_mild_severe,
              p_severe_critical, p_critical_death, mean_time_to_isolate_factor, lockdown_factor_age,
              p_infect_household, fraction_stay_home, params, provider="mt", tracing_enabled=1,
              home_real=None, initial_infected=None, tables=None):
    """Restated `run_model`.  Returns (20-tuple in the reference's order, extras dict).

    provider "mt": numba-compatible MT19937 stream, bit-identical to the reference.
    provider "replay"/"native": keyed Philox draws (include/seir_draws.h); the prologue results
    `home_real`, `initial_infected` and the draw tables `tables` (dict with enlam_pre, enlam_post,
    nat_enlam, nat_alias_prob, nat_alias_idx) come from the host code under test.
    """
    prov = PROVIDERS[provider]
    n = int(params["n"]); T = int(params["T"]); n_ages = int(params["n_ages"])
    hh = np.ascontiguousarray(households, dtype=np.int64)
    assert hh.shape[0] == n
    age_ = np.ascontiguousarray(age, dtype=np.int64)
    diab = np.ascontiguousarray(diabetes, dtype=np.int64)
    hyp = np.ascontiguousarray(hypertension, dtype=np.int64)
    offs, members = age_groups_csr(age_groups)
    cmat = np.ascontiguousarray(contact_matrix, dtype=np.float64)
    pms = np.ascontiguousarray(p_mild_severe, dtype=np.float64)
    psc = np.ascontiguousarray(p_severe_critical, dtype=np.float64)
    pcd = np.ascontiguousarray(p_critical_death, dtype=np.float64)
    iso = np.ascontiguousarray(mean_time_to_isolate_factor, dtype=np.int64)
    pih = np.ascontiguousarray(p_infect_household, dtype=np.float64)
    fsh = np.ascontiguousarray(fraction_stay_home, dtype=np.float64)

    P = Params()
    P.n, P.T, P.n_ages, P.W = n, T, n_ages, hh.shape[1]
    P.t_lockdown = int(params["t_lockdown"]); P.t_tracing_start = int(params["t_tracing_start"])
    P.t_stayhome_start = int(params["t_stayhome_start"])
    P.tracing_enabled = int(tracing_enabled); P.provider = prov; P.n_iso_rows = iso.shape[0]
    P.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    for k in ("lockdown_factor", "p_trace_household", "p_trace_outside", "p_documented_in_mild",
              "p_infect_given_contact", "asymptomatic_transmissibility", "mean_time_to_isolate",
              "mean_time_to_isolate_asympt", "mean_time_to_severe", "mean_time_mild_recovery",
              "mean_time_to_critical", "mean_time_severe_recovery", "mean_time_critical_recovery",
              "mean_time_to_death", "time_to_activation_mean", "time_to_activation_std",
              "initial_infected_fraction"):
        setattr(P, k, float(params[k]))

    keep = [hh, age_, diab, hyp, offs, members, cmat, pms, psc, pcd, iso, pih, fsh]
    I = Inputs()
    for name, arr in zip(_IN_FIELDS[:13], keep):
        setattr(I, name, _ptr(arr))
    n_init_mt = int(float(params["initial_infected_fraction"]) * n)
    if prov != MT:
        hr = np.ascontiguousarray(home_real, dtype=np.uint8)
        ii = np.ascontiguousarray(initial_infected, dtype=np.int64)
        assert hr.shape == (n,)
        P.n_initial = ii.shape[0]
        I.home_real, I.initial_infected = _ptr(hr), _ptr(ii)
        tb = {k: np.ascontiguousarray(tables[k], dtype=np.float64)
              for k in ("enlam_pre", "enlam_post", "nat_enlam", "nat_alias_prob")}
        tb["nat_alias_idx"] = np.ascontiguousarray(tables["nat_alias_idx"], dtype=np.int32)
        assert tb["enlam_pre"].shape == (n_ages, n_ages) and tb["nat_enlam"].shape == (2, n_ages, 2)
        assert tb["nat_alias_prob"].shape == (n_ages, n_ages) and tb["nat_alias_idx"].shape == (n_ages, n_ages)
        I.enlam_pre, I.enlam_post = _ptr(tb["enlam_pre"]), _ptr(tb["enlam_post"])
        I.nat_enlam = _ptr(tb["nat_enlam"])
        I.nat_alias_prob, I.nat_alias_idx = _ptr(tb["nat_alias_prob"]), _ptr(tb["nat_alias_idx"])
        keep += [hr, ii, tb]
        n_init = ii.shape[0]
    else:
        n_init = n_init_mt

    out = {}
    for k in _OUT_FIELDS[:9]:
        out[k] = np.empty((T, n), dtype=np.uint8)
    for k in _OUT_FIELDS[9:22]:
        out[k] = np.empty(n, dtype=np.float64)
    out["num_infected_by_outside"] = np.empty(n, dtype=np.int32)
    out["home_real_out"] = np.empty(n, dtype=np.uint8)
    out["initial_infected_out"] = np.empty(max(n_init, 1), dtype=np.int64)
    out["traced"] = np.empty(n, dtype=np.uint8)
    out["draw_count"] = np.zeros(1, dtype=np.int64)
    O = Outputs()
    for k in _OUT_FIELDS:
        setattr(O, k, _ptr(out[k]))

    rc = lib().oracle_run_model(C.byref(P), C.byref(I), C.byref(O))
    if rc != 0:
        raise RuntimeError("oracle_run_model failed rc=%d" % rc)
    hist = [out[k].view(np.bool_) for k in ("S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q")]
    tup = tuple(hist) + (out["num_infected_by"], out["time_documented"], out["time_to_activation"],
                         out["time_to_death"], out["time_to_recovery"], out["time_critical"],
                         out["time_exposed"], out["num_infected_asympt"], age_, out["time_infected"],
                         out["time_to_severe"])
    extras = {k: out[k] for k in ("time_to_isolate", "time_severe", "time_to_critical",
                                  "num_infected_by_outside", "traced", "draw_count")}
    extras["home_real"] = out["home_real_out"].view(np.bool_)
    extras["initial_infected"] = out["initial_infected_out"][:n_init]
    return tup, extras


OUTPUT_NAMES = ("S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q", "num_infected_by",
                "time_documented", "time_to_activation", "time_to_death", "time_to_recovery",
                "time_critical", "time_exposed", "num_infected_asympt", "age", "time_infected",
                "time_to_severe")


def perm_picks(seed, site, group, n, k):
    """First k entries of the keyed permutation of [0, n) (include/seir_draws.h, sd_perm) from the C header."""
    out = np.zeros(k, dtype=np.int64)
    L = lib()
    L.oracle_perm_picks.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p]
    L.oracle_perm_picks.restype = None
    L.oracle_perm_picks(int(seed) & 0xFFFFFFFFFFFFFFFF, site, group, n, k, out.ctypes.data)
    return out

