"""ctypes binding of libseir_b200.so (include/seir_b200.h) -- the only way Python reaches the GPU.

There is no CPU fallback: if the library cannot be built/loaded, or no B200 is present,
`Engine(...)` raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from . import tables as _tables

MODE_NATIVE, MODE_REPLAY = 0, 1
LAUNCH_PERSISTENT, LAUNCH_PER_DAY = 0, 1
MODES = {"native": MODE_NATIVE, "replay": MODE_REPLAY}
LAUNCHES = {"persistent": LAUNCH_PERSISTENT, "per_day": LAUNCH_PER_DAY}

VEC = {"num_infected_by": 0, "time_documented": 1, "time_to_activation": 2, "time_to_death": 3,
       "time_to_recovery": 4, "time_critical": 5, "time_exposed": 6, "num_infected_asympt": 7,
       "time_infected": 8, "time_to_severe": 9, "time_to_isolate": 10, "time_severe": 11,
       "time_to_critical": 12, "num_infected_by_outside": 13}
PLANE = {"S": 0, "E": 1, "Mild": 2, "Documented": 3, "Severe": 4, "Critical": 5, "R": 6, "D": 7, "Q": 8,
         "packed": 9}
PLANES_IN_ORDER = ("S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q")

EXPORTS = ("seir_create", "seir_set_params", "seir_run", "seir_last_run_ms", "seir_get_tallies",
           "seir_get_agent_vector", "seir_get_history", "seir_get_counters", "seir_destroy",
           "seir_last_error", "seir_device_count", "seir_build_info", "seir_probe_draws",
           "seir_host_register", "seir_host_unregister", "seir_last_run_ms3", "seir_get_day_times",
           "seir_get_stage_times", "seir_get_cohorts", "seir_shard_export", "seir_shard_attach")


class SeirParams(C.Structure):
    _fields_ = [("n", C.c_int64), ("T", C.c_int32), ("t_lockdown", C.c_int32),
                ("t_stayhome_start", C.c_int32), ("t_tracing_start", C.c_int32),
                ("tracing_enabled", C.c_int32), ("mode", C.c_int32), ("launch", C.c_int32),
                ("reserved0", C.c_int32)] + [(k, C.c_double) for k in (
                    "p_documented_in_mild", "p_infect_given_contact", "asymptomatic_transmissibility",
                    "mean_time_to_severe", "mean_time_mild_recovery", "mean_time_to_critical",
                    "mean_time_severe_recovery", "mean_time_critical_recovery", "mean_time_to_death",
                    "time_to_activation_mean", "time_to_activation_std", "p_trace_household",
                    "p_trace_outside")]


class SeirPopulation(C.Structure):
    _fields_ = [("n", C.c_int64), ("W", C.c_int32), ("reserved0", C.c_int32)] + [
        (k, C.c_void_p) for k in ("households", "age", "diabetes", "hypertension", "group_offsets",
                                  "group_members", "p_infect_household")]


class SeirTables(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("p_mild_severe", "p_severe_critical", "p_critical_death",
                                          "iso_mean", "iso_asympt_mean", "nat_enlam", "nat_alias_prob",
                                          "nat_alias_idx", "enlam_pre", "enlam_post")]


class SeirReplica(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("home_real", C.c_void_p), ("initial_infected", C.c_void_p),
                ("n_initial", C.c_int64)]


_lib = None


def load_library():
    """Build (if stale) and dlopen libseir_b200.so.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    lib = C.CDLL(path)
    H = C.c_void_p
    lib.seir_create.argtypes = [C.POINTER(SeirParams), C.POINTER(SeirPopulation), C.POINTER(SeirTables),
                                C.c_int, C.c_int, C.POINTER(H)]
    lib.seir_set_params.argtypes = [H, C.POINTER(SeirParams), C.POINTER(SeirTables)]
    lib.seir_run.argtypes = [H, C.POINTER(SeirReplica)]
    lib.seir_last_run_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.seir_last_run_ms3.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.seir_get_tallies.argtypes = [H, C.c_int, C.c_void_p]
    lib.seir_get_agent_vector.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.seir_get_history.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.seir_get_counters.argtypes = [H, C.c_int, C.c_void_p]
    lib.seir_get_cohorts.argtypes = [H, C.c_int, C.c_void_p]
    lib.seir_shard_export.argtypes = [H, C.c_void_p]
    lib.seir_shard_attach.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.seir_get_day_times.argtypes = [H, C.c_void_p]
    lib.seir_get_stage_times.argtypes = [H, C.c_void_p]
    lib.seir_destroy.argtypes = [H]
    lib.seir_destroy.restype = None
    lib.seir_last_error.restype = C.c_char_p
    lib.seir_build_info.restype = C.c_char_p
    lib.seir_device_count.restype = C.c_int
    lib.seir_host_register.argtypes = [C.c_void_p, C.c_uint64]
    lib.seir_host_unregister.argtypes = [C.c_void_p]
    lib.seir_probe_draws.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        msg = load_library().seir_last_error().decode()
        if "lambda < 0" in msg or "out of range" in msg or "invalid" in msg or "must be" in msg:
            raise ValueError(msg)
        raise RuntimeError(msg)


def _ptr(a):
    return None if a is None else a.ctypes.data


def make_params(p, mode="native", launch="persistent", tracing_enabled=1):
    """Pack the reference's `params` mapping into the POD struct of the C ABI."""
    d = _tables.params_to_dict(p)
    sp = SeirParams()
    sp.n = int(d["n"])
    sp.T = int(d["T"])
    clamp = lambda v: int(max(-2**31 + 1, min(2**31 - 1, v)))  # noqa: E731  (e.g. year-3000 stay-home dates)
    sp.t_lockdown = clamp(int(d["t_lockdown"]))
    sp.t_stayhome_start = clamp(int(d["t_stayhome_start"]))
    sp.t_tracing_start = clamp(int(d["t_tracing_start"]))
    sp.tracing_enabled = int(tracing_enabled)
    sp.mode = MODES[mode] if isinstance(mode, str) else int(mode)
    sp.launch = LAUNCHES[launch] if isinstance(launch, str) else int(launch)
    for k in ("p_documented_in_mild", "p_infect_given_contact", "asymptomatic_transmissibility",
              "mean_time_to_severe", "mean_time_mild_recovery", "mean_time_to_critical",
              "mean_time_severe_recovery", "mean_time_critical_recovery", "mean_time_to_death",
              "time_to_activation_mean", "time_to_activation_std", "p_trace_household", "p_trace_outside"):
        setattr(sp, k, d[k])
    return sp


class Engine:
    """One handle = one device-resident population + state for `n_replicas` concurrent runs."""

    def __init__(self, params, households, age, age_groups, diabetes, hypertension, p_infect_household,
                 tables, n_replicas=1, device=0, mode="native", launch="persistent", tracing_enabled=1):
        self.lib = load_library()
        if self.lib.seir_device_count() < 1:
            raise RuntimeError("no CUDA device: libseir_b200 has no CPU fallback")
        self._h = C.c_void_p()
        self.mode, self.launch, self.tracing_enabled = mode, launch, tracing_enabled
        self.sp = make_params(params, mode, launch, tracing_enabled)
        self.n, self.T, self.n_replicas = int(self.sp.n), int(self.sp.T), int(n_replicas)
        hh = np.ascontiguousarray(households, dtype=np.int64)
        if hh.ndim != 2 or hh.shape[0] != self.n:
            raise ValueError("households must be [n, W]")
        if isinstance(age_groups, tuple) and len(age_groups) == 2 and getattr(age_groups[0], "shape", None) == (102,):
            offs, members = age_groups  # already CSR
        else:
            lens = np.array([len(g) for g in age_groups], dtype=np.int64)
            offs = np.zeros(len(lens) + 1, dtype=np.int64)
            np.cumsum(lens, out=offs[1:])
            members = np.concatenate([np.asarray(g, dtype=np.int64) for g in age_groups])
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        members = np.ascontiguousarray(members, dtype=np.int64)
        if offs.shape[0] != _tables.N_AGES + 1:
            raise ValueError("age_groups must have %d entries" % _tables.N_AGES)
        self.group_sizes = np.diff(offs)
        arrs = [hh, np.ascontiguousarray(age, dtype=np.int64), np.ascontiguousarray(diabetes, dtype=np.int64),
                np.ascontiguousarray(hypertension, dtype=np.int64), offs, members,
                np.ascontiguousarray(p_infect_household, dtype=np.float64)]
        for a in arrs[1:4] + arrs[6:]:
            if a.shape != (self.n,):
                raise ValueError("per-agent inputs must have shape (n,)")
        pop = SeirPopulation()
        pop.n, pop.W = self.n, hh.shape[1]
        for k, a in zip(("households", "age", "diabetes", "hypertension", "group_offsets", "group_members",
                         "p_infect_household"), arrs):
            setattr(pop, k, _ptr(a))
        self._tables_keep = None
        tb = self._pack_tables(tables)
        _check(self.lib.seir_create(C.byref(self.sp), C.byref(pop), C.byref(tb), self.n_replicas, int(device),
                                    C.byref(self._h)))

    def _pack_tables(self, tables):
        tb = SeirTables()
        keep = {}
        for k, _ in SeirTables._fields_:
            v = tables.get(k)
            if v is not None:
                v = np.ascontiguousarray(v, dtype=np.int32 if k == "nat_alias_idx" else np.float64)
            keep[k] = v
            setattr(tb, k, _ptr(v))
        self._tables_keep = keep
        return tb

    def shard_export(self):
        """320 bytes: the CUDA IPC handles of this rank's arrays (include/seir_b200.h, seir_shard_export)."""
        out = np.zeros(5 * 64, dtype=np.uint8)
        _check(self.lib.seir_shard_export(self._h, _ptr(out)))
        return out

    def shard_attach(self, rank, world, all_handles, bounds):
        """all_handles uint8[world, 320] (rank-major), bounds int64[world + 1] (household starts)."""
        hd = np.ascontiguousarray(all_handles, dtype=np.uint8).reshape(world, 5 * 64)
        bd = np.ascontiguousarray(bounds, dtype=np.int64)
        if bd.shape != (world + 1,):
            raise ValueError("bounds must have world + 1 entries")
        _check(self.lib.seir_shard_attach(self._h, int(rank), int(world), _ptr(hd), _ptr(bd)))
        self.shard = (int(rank), int(world), bd.copy())

    def set_params(self, params=None, tables=None, mode=None, launch=None, tracing_enabled=None):
        if mode is not None:
            self.mode = mode
        if launch is not None:
            self.launch = launch
        if tracing_enabled is not None:
            self.tracing_enabled = tracing_enabled
        if params is not None:
            sp = make_params(params, self.mode, self.launch, self.tracing_enabled)
            if int(sp.n) != self.n:
                raise ValueError("params['n'] does not match the resident population")
        else:
            sp = self.sp
            sp.mode = MODES[self.mode]
            sp.launch = LAUNCHES[self.launch]
            sp.tracing_enabled = int(self.tracing_enabled)
        tb = self._pack_tables(tables) if tables is not None else None
        _check(self.lib.seir_set_params(self._h, C.byref(sp), C.byref(tb) if tb is not None else None))
        self.sp, self.T = sp, int(sp.T)

    def run(self, seeds, home_real=None, initial_infected=None):
        """seeds[r], home_real[r] (bool[n] or None), initial_infected[r] (int array) per replica."""
        R = self.n_replicas
        seeds = list(seeds)
        if len(seeds) != R:
            raise ValueError("need one seed per replica")
        home_real = list(home_real) if home_real is not None else [None] * R
        initial_infected = list(initial_infected) if initial_infected is not None else [np.zeros(0, np.int64)] * R
        reps = (SeirReplica * R)()
        keep = []
        for r in range(R):
            reps[r].seed = int(seeds[r]) & 0xFFFFFFFFFFFFFFFF
            hr = home_real[r]
            if hr is not None:
                hr = np.ascontiguousarray(hr, dtype=np.uint8)
                if hr.shape != (self.n,):
                    raise ValueError("home_real must have shape (n,)")
            ii = np.ascontiguousarray(initial_infected[r], dtype=np.int64)
            keep += [hr, ii]
            reps[r].home_real = _ptr(hr)
            reps[r].initial_infected = _ptr(ii) if ii.size else None
            reps[r].n_initial = ii.size
        _check(self.lib.seir_run(self._h, reps))

    def last_run_ms(self):
        a, b = C.c_float(), C.c_float()
        _check(self.lib.seir_last_run_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_run_ms3(self):
        """(state init, day loop, history materialisation) device milliseconds of the last run."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(self.lib.seir_last_run_ms3(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def tallies(self, replica=0):
        """int64[T, 9, 101]: per-day counts by age, planes S,E,Mild,Documented,Severe,Critical,R,D,Q."""
        out = np.empty((self.T, 9, _tables.N_AGES), dtype=np.int64)
        _check(self.lib.seir_get_tallies(self._h, replica, _ptr(out)))
        return out

    def cohorts(self, replica=0):
        """int64[T, 4, 101]: per exposure day and age -- agents exposed, sum of num_infected_by, dead at
        the end, recovered at the end (see driver_tallies.py)."""
        out = np.empty((self.T, 4, _tables.N_AGES), dtype=np.int64)
        _check(self.lib.seir_get_cohorts(self._h, replica, _ptr(out)))
        return out

    def agent_vector(self, which, replica=0):
        out = np.empty(self.n, dtype=np.float64)
        _check(self.lib.seir_get_agent_vector(self._h, replica, VEC[which] if isinstance(which, str) else which,
                                              _ptr(out)))
        return out

    def history(self, which, t0=0, t1=None, replica=0):
        t1 = self.T if t1 is None else t1
        out = np.empty((t1 - t0, self.n), dtype=np.uint8)
        _check(self.lib.seir_get_history(self._h, replica, PLANE[which] if isinstance(which, str) else which,
                                         t0, t1, _ptr(out)))
        return out if which in ("packed", 9) else out.view(np.bool_)

    def day_times_us(self):
        """Duration of every day pass of the last persistent run in microseconds (index t-1 = pass t)."""
        out = np.zeros(self.T + 2, dtype=np.int64)
        _check(self.lib.seir_get_day_times(self._h, _ptr(out)))
        return np.diff(out[1:]) / 1e3

    def stage_times_us(self):
        """[T+2, 8] stage stamps (us, relative to the start of each pass) of CTA 0's first round."""
        out = np.zeros((self.T + 2) * 8, dtype=np.int64)
        _check(self.lib.seir_get_stage_times(self._h, _ptr(out)))
        day = np.zeros(self.T + 2, dtype=np.int64)
        _check(self.lib.seir_get_day_times(self._h, _ptr(day)))
        st = out.reshape(self.T + 2, 8).astype(np.float64)
        rel = np.where(st > 0, (st - day[:, None]) / 1e3, np.nan)
        return rel

    def counters(self, replica=0):
        out = np.zeros(4, dtype=np.int64)
        _check(self.lib.seir_get_counters(self._h, replica, _ptr(out)))
        return {"launches": int(out[0]), "active_agent_days": int(out[1]), "infections": int(out[2]),
                "hits": int(out[3])}

    def close(self):
        if self._h:
            self.lib.seir_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pin(array):
    """Page-lock a NumPy buffer (per-run inputs of Engine.run then move by DMA)."""
    _check(load_library().seir_host_register(_ptr(array), array.nbytes))
    return array


def unpin(array):
    _check(load_library().seir_host_unregister(_ptr(array)))


def probe_draws(seed, m, mean, mu, sigma, enlam, log_in, device=0):
    """Device evaluation of the draw contract (include/seir_draws.h) for parity tests."""
    lib = load_library()
    log_in = np.ascontiguousarray(log_in, dtype=np.float64)
    assert log_in.shape == (m,)
    e, l, g = np.empty(m), np.empty(m), np.empty(m)
    p = np.empty(m, dtype=np.int64)
    _check(lib.seir_probe_draws(device, seed, m, mean, mu, sigma, enlam, _ptr(e), _ptr(l), _ptr(p), _ptr(g),
                                _ptr(log_in)))
    return e, l, p, g
