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
HISTORY = {"eager": 0, "on_demand": 1}
MODES = {"native": MODE_NATIVE, "replay": MODE_REPLAY}
LAUNCHES = {"persistent": LAUNCH_PERSISTENT, "per_day": LAUNCH_PER_DAY}

VEC = {"num_infected_by": 0, "time_documented": 1, "time_to_activation": 2, "time_to_death": 3,
       "time_to_recovery": 4, "time_critical": 5, "time_exposed": 6, "num_infected_asympt": 7,
       "time_infected": 8, "time_to_severe": 9, "time_to_isolate": 10, "time_severe": 11,
       "time_to_critical": 12, "num_infected_by_outside": 13}
PLANE = {"S": 0, "E": 1, "Mild": 2, "Documented": 3, "Severe": 4, "Critical": 5, "R": 6, "D": 7, "Q": 8,
         "packed": 9}
VEC_PREVIOUS = 256   # include/seir_b200.h SEIR_VEC_PREVIOUS
PLANES_IN_ORDER = ("S", "E", "Mild", "Documented", "Severe", "Critical", "R", "D", "Q")

EXPORTS = ("seir_create", "seir_set_params", "seir_run", "seir_last_run_ms", "seir_get_tallies",
           "seir_get_agent_vector", "seir_get_history", "seir_get_counters", "seir_destroy",
           "seir_last_error", "seir_device_count", "seir_build_info", "seir_probe_draws",
           "seir_host_register", "seir_host_unregister", "seir_last_run_ms3", "seir_get_day_times",
           "seir_get_stage_times", "seir_get_cohorts", "seir_shard_export", "seir_shard_attach", "seir_get_visits",
           "seir_history_preserve", "seir_run_ex", "seir_set_table_bank", "seir_get_prologue", "seir_abi_sizes",
           "seir_shard_set_threshold", "seir_prologue_mt", "seir_format_tsv")


SHARD_HANDLES = 6   # include/seir_b200.h SEIR_SHARD_HANDLES


class SeirParams(C.Structure):
    _fields_ = [("n", C.c_int64), ("T", C.c_int32), ("t_lockdown", C.c_int32),
                ("t_stayhome_start", C.c_int32), ("t_tracing_start", C.c_int32),
                ("tracing_enabled", C.c_int32), ("mode", C.c_int32), ("launch", C.c_int32),
                ("history_mode", C.c_int32)] + [(k, C.c_double) for k in (
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
                ("n_initial", C.c_int64), ("fraction_stay_home", C.c_void_p), ("initial_infected_fraction", C.c_double),
                ("keyed_prologue", C.c_int32), ("reserved0", C.c_int32)]


class SeirOverrides(C.Structure):
    _fields_ = [("T", C.c_int32), ("t_lockdown", C.c_int32), ("t_stayhome_start", C.c_int32),
                ("t_tracing_start", C.c_int32), ("table_set", C.c_int32), ("reserved0", C.c_int32)]


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
    lib.seir_run_ex.argtypes = [H, C.POINTER(SeirReplica), C.POINTER(SeirOverrides)]
    lib.seir_set_table_bank.argtypes = [H, C.c_int, C.POINTER(SeirTables), C.c_void_p]
    lib.seir_get_prologue.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.seir_last_run_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.seir_last_run_ms3.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.seir_get_tallies.argtypes = [H, C.c_int, C.c_void_p]
    lib.seir_get_agent_vector.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.seir_get_history.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.seir_get_counters.argtypes = [H, C.c_int, C.c_void_p]
    lib.seir_get_cohorts.argtypes = [H, C.c_int, C.c_void_p]
    lib.seir_get_visits.argtypes = [H, C.c_int, C.c_void_p]
    lib.seir_history_preserve.argtypes = [H]
    lib.seir_shard_export.argtypes = [H, C.c_void_p]
    lib.seir_shard_attach.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.seir_shard_set_threshold.argtypes = [H, C.c_int64]
    lib.seir_get_day_times.argtypes = [H, C.c_void_p]
    lib.seir_get_stage_times.argtypes = [H, C.c_void_p]
    lib.seir_destroy.argtypes = [H]
    lib.seir_destroy.restype = None
    lib.seir_last_error.restype = C.c_char_p
    lib.seir_build_info.restype = C.c_char_p
    lib.seir_device_count.restype = C.c_int
    lib.seir_host_register.argtypes = [C.c_void_p, C.c_uint64]
    lib.seir_host_unregister.argtypes = [C.c_void_p]
    lib.seir_prologue_mt.argtypes = [C.c_uint32, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.seir_format_tsv.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
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


def make_params(p, mode="native", launch="persistent", tracing_enabled=1, history="eager"):
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
    sp.history_mode = HISTORY[history]
    for k in ("p_documented_in_mild", "p_infect_given_contact", "asymptomatic_transmissibility",
              "mean_time_to_severe", "mean_time_mild_recovery", "mean_time_to_critical",
              "mean_time_severe_recovery", "mean_time_critical_recovery", "mean_time_to_death",
              "time_to_activation_mean", "time_to_activation_std", "p_trace_household", "p_trace_outside"):
        setattr(sp, k, d[k])
    return sp


class Engine:
    """One handle = one device-resident population + state for `n_replicas` concurrent runs."""

    def __init__(self, params, households, age, age_groups, diabetes, hypertension, p_infect_household,
                 tables, n_replicas=1, device=0, mode="native", launch="persistent", tracing_enabled=1,
                 history="eager"):
        self.lib = load_library()
        if self.lib.seir_device_count() < 1:
            raise RuntimeError("no CUDA device: libseir_b200 has no CPU fallback")
        self._h = C.c_void_p()
        self.run_serial = 0   # runs so far: a lazy history view knows which run it belongs to
        self.prev_serial = -1  # the run whose history has been set aside on the device (seir_history_preserve)
        self._views = []       # weak references to the lazy history views handed out for the last two runs
        self.mode, self.launch, self.tracing_enabled = mode, launch, tracing_enabled
        self.history_mode = history   # "eager": every run writes the T x n history; "on_demand": rows computed when asked for
        self.sp = make_params(params, mode, launch, tracing_enabled, history)
        self.n, self.T, self.n_replicas = int(self.sp.n), int(self.sp.T), int(n_replicas)
        self.T_rep = [self.T] * self.n_replicas   # per-replica horizons of the last run (overrides)
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
        self.group_csr = (offs, members)   # kept for prologue_mt (the reference's picks per seed, host only)
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

    def set_table_bank(self, table_sets, p_infect_given_contact):
        """Register table sets for per-replica overrides (`overrides=[{"table_set": k, ...}]` in run): one launch can
        then batch runs of different grid points of a sweep (parameter_sweep.py:54-56).  Set 0 replaces the engine's own."""
        arr = (SeirTables * len(table_sets))()
        keep = []
        for k, tables in enumerate(table_sets):
            for name, _ in SeirTables._fields_:
                v = tables.get(name)
                if v is not None:
                    v = np.ascontiguousarray(v, dtype=np.int32 if name == "nat_alias_idx" else np.float64)
                keep.append(v)
                setattr(arr[k], name, _ptr(v))
        pc = np.ascontiguousarray(p_infect_given_contact, dtype=np.float64)
        if pc.shape != (len(table_sets),):
            raise ValueError("need one p_infect_given_contact per table set")
        _check(self.lib.seir_set_table_bank(self._h, len(table_sets), arr, _ptr(pc)))
        self.n_table_sets = len(table_sets)

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
        """SHARD_HANDLES x 64 bytes: the CUDA IPC handles of this rank's arrays (include/seir_b200.h, seir_shard_export)."""
        out = np.zeros(SHARD_HANDLES * 64, dtype=np.uint8)
        _check(self.lib.seir_shard_export(self._h, _ptr(out)))
        return out

    def shard_attach(self, rank, world, all_handles, bounds, threshold=None):
        """all_handles uint8[world, SHARD_HANDLES * 64] (rank-major), bounds int64[world + 1] (household starts).  threshold: list
        entries of a day up to which the run stays replicated (None: the library's default; 0: sharded from day 1)."""
        hd = np.ascontiguousarray(all_handles, dtype=np.uint8).reshape(world, SHARD_HANDLES * 64)
        bd = np.ascontiguousarray(bounds, dtype=np.int64)
        if bd.shape != (world + 1,):
            raise ValueError("bounds must have world + 1 entries")
        _check(self.lib.seir_shard_attach(self._h, int(rank), int(world), _ptr(hd), _ptr(bd)))
        if threshold is not None:
            _check(self.lib.seir_shard_set_threshold(self._h, int(threshold)))
        self.shard = (int(rank), int(world), bd.copy())

    def set_params(self, params=None, tables=None, mode=None, launch=None, tracing_enabled=None):
        if mode is not None:
            self.mode = mode
        if launch is not None:
            self.launch = launch
        if tracing_enabled is not None:
            self.tracing_enabled = tracing_enabled
        if params is not None:
            sp = make_params(params, self.mode, self.launch, self.tracing_enabled, self.history_mode)
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

    def run(self, seeds, home_real=None, initial_infected=None, keyed=None, overrides=None):
        """seeds[r] per replica, and the prologue's picks (seir_individual.py:176-187) either from the caller --
        home_real[r] (bool[n] or None), initial_infected[r] (int array) -- or made on the device from the keyed
        contract: keyed[r] = (fraction_stay_home float[101] or None, initial_infected_fraction) (one tuple for all
        replicas is accepted too).  overrides[r]: dict with any of T, t_lockdown, t_stayhome_start, t_tracing_start,
        table_set (see set_table_bank) -- one launch over different grid points of a sweep."""
        R = self.n_replicas
        seeds = list(seeds)
        if len(seeds) != R:
            raise ValueError("need one seed per replica")
        home_real = list(home_real) if home_real is not None else [None] * R
        initial_infected = list(initial_infected) if initial_infected is not None else [np.zeros(0, np.int64)] * R
        if keyed is not None and isinstance(keyed, tuple) and len(keyed) == 2 and not isinstance(keyed[0], tuple):
            keyed = [keyed] * R
        reps = (SeirReplica * R)()
        keep = []
        for r in range(R):
            reps[r].seed = int(seeds[r]) & 0xFFFFFFFFFFFFFFFF
            if keyed is not None and keyed[r] is not None:
                fsh, frac = keyed[r]
                if fsh is not None:
                    fsh = np.ascontiguousarray(fsh, dtype=np.float64)
                    if fsh.shape != (_tables.N_AGES,):
                        raise ValueError("fraction_stay_home must have 101 entries")
                keep.append(fsh)
                reps[r].fraction_stay_home = _ptr(fsh)
                reps[r].initial_infected_fraction = float(frac)
                reps[r].keyed_prologue = 1
                continue
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
        ov = None
        T_rep = [self.T] * R
        if overrides is not None:
            if len(overrides) != R:
                raise ValueError("need one overrides entry per replica")
            ov = (SeirOverrides * R)()
            for r, o in enumerate(overrides):
                o = o or {}
                unknown = set(o) - {"T", "t_lockdown", "t_stayhome_start", "t_tracing_start", "table_set"}
                if unknown:
                    raise ValueError("unknown override(s): %s" % sorted(unknown))
                clamp = lambda v: int(max(-2**31 + 1, min(2**31 - 1, v)))  # noqa: E731
                ov[r].T = int(o.get("T", self.sp.T))
                ov[r].t_lockdown = clamp(o.get("t_lockdown", self.sp.t_lockdown))
                ov[r].t_stayhome_start = clamp(o.get("t_stayhome_start", self.sp.t_stayhome_start))
                ov[r].t_tracing_start = clamp(o.get("t_tracing_start", self.sp.t_tracing_start))
                ov[r].table_set = int(o.get("table_set", 0))
                T_rep[r] = int(ov[r].T)
        self._keep_live_histories()
        self.run_serial += 1
        _check(self.lib.seir_run_ex(self._h, reps, ov))
        self.T_rep = T_rep

    def prologue_picks(self, replica=0):
        """(home_real bool[n], initial_infected int64[k]) as the device holds them for the last run."""
        home = np.zeros(self.n, dtype=np.uint8)
        k = np.zeros(1, dtype=np.int64)
        _check(self.lib.seir_get_prologue(self._h, replica, _ptr(home), None, 0, _ptr(k)))
        init = np.zeros(int(k[0]), dtype=np.int64)
        _check(self.lib.seir_get_prologue(self._h, replica, None, _ptr(init), init.size, _ptr(k)))
        return home.view(np.bool_), init

    def register_view(self, view):
        import weakref
        self._views.append(weakref.ref(view))

    def _keep_live_histories(self):
        """The reference hands out fresh arrays per call; here the next run overwrites the device-resident history.
        Views of the last run that somebody still holds keep working: their history is set aside on the device (one
        device-to-device copy); views of the run before that, if still alive, take a host copy."""
        live = [v for v in (r() for r in self._views) if v is not None]
        older = [v for v in live if v._serial == self.prev_serial and v._host is None]
        planes = [v for v in older if v.ndim == 2]
        if planes:
            packed = self.history("packed", previous=True, rows=planes[0].shape[0])
            for v in planes:
                v._host = packed
        for v in older:
            if v.ndim == 1:       # a per-agent vector nobody has used yet: it takes its host copy now
                v._get()
        last = [v for v in live if v._serial == self.run_serial and v._host is None]
        if last and self.history_mode != "eager":   # (nothing is set aside on the device in this mode)
            for v in last:
                if v.ndim == 1:
                    v._get()
            last = [v for v in last if v._host is None]
        if last:
            _check(self.lib.seir_history_preserve(self._h))
            self.prev_serial = self.run_serial
        self._views = [r for r in self._views if r() is not None and r()._host is None]

    def keep_views_on_host(self):
        """Before the handle goes away while somebody still holds lazy results of its last two runs (the drop-in
        module's engine cache evicting a population): every live history takes a host copy of its run's packed
        history, every live vector becomes the host array it stands for."""
        live = [v for v in (r() for r in self._views) if v is not None and v._host is None]
        for serial, previous in ((self.run_serial, False), (self.prev_serial, True)):
            planes = [v for v in live if v.ndim == 2 and v._serial == serial]
            if planes and (not previous or serial != self.run_serial):
                packed = self.history("packed", previous=previous, rows=planes[0].shape[0] if previous else None)
                for v in planes:
                    v._host = packed
        for v in live:
            if v.ndim == 1:
                v._get()
        self._views = []

    def last_run_ms(self):
        a, b = C.c_float(), C.c_float()
        _check(self.lib.seir_last_run_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_run_ms3(self):
        """(state init, day loop, history materialisation) device milliseconds of the last run."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(self.lib.seir_last_run_ms3(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def tallies(self, replica=0, out=None):
        """int64[T, 9, 101]: per-day counts by age, planes S,E,Mild,Documented,Severe,Critical,R,D,Q.
        out: a C-contiguous int64[T, 9, 101] buffer of the caller's to fill (a loop over many runs then touches no
        fresh memory per run)."""
        if out is None:
            out = np.empty((self.T, 9, _tables.N_AGES), dtype=np.int64)
        elif out.shape != (self.T, 9, _tables.N_AGES) or out.dtype != np.int64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous int64[T, 9, 101] array")
        _check(self.lib.seir_get_tallies(self._h, replica, _ptr(out)))
        return out[:self.T_rep[replica]]

    def cohorts(self, replica=0):
        """int64[T, 4, 101]: per exposure day and age -- agents exposed, sum of num_infected_by, dead at
        the end, recovered at the end (see driver_tallies.py)."""
        out = np.empty((self.T, 4, _tables.N_AGES), dtype=np.int64)
        _check(self.lib.seir_get_cohorts(self._h, replica, _ptr(out)))
        return out

    def agent_vector(self, which, replica=0, previous=False):
        """float64[n]; previous: of the run set aside by the last seir_history_preserve (see _keep_live_histories)."""
        out = np.empty(self.n, dtype=np.float64)
        code = (VEC[which] if isinstance(which, str) else which) | (VEC_PREVIOUS if previous else 0)
        _check(self.lib.seir_get_agent_vector(self._h, replica, code, _ptr(out)))
        return out

    def history(self, which, t0=0, t1=None, replica=0, previous=False, rows=None):
        t1 = (self.T_rep[replica] if rows is None else rows) if t1 is None else t1
        out = np.empty((t1 - t0, self.n), dtype=np.uint8)
        code = (PLANE[which] if isinstance(which, str) else which) | (16 if previous else 0)
        _check(self.lib.seir_get_history(self._h, replica, code, t0, t1, _ptr(out)))
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
        vis = np.zeros(1, dtype=np.int64)
        _check(self.lib.seir_get_visits(self._h, replica, _ptr(vis)))
        return {"launches": int(out[0]), "active_agent_days": int(out[1]), "infections": int(out[2]),
                "hits": int(out[3]), "visits": int(vis[0])}

    def close(self):
        if self._h:
            self.lib.seir_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def group_csr(age_groups):
    """The reference's `age_groups` tuple (seir_individual.py:36) as (offsets int64[n_ages+1], members int64[n])."""
    lens = np.array([len(g) for g in age_groups], dtype=np.int64)
    offs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    members = (np.concatenate([np.asarray(g, dtype=np.int64) for g in age_groups]) if offs[-1] > 0
               else np.zeros(0, dtype=np.int64))
    return offs, np.ascontiguousarray(members, dtype=np.int64)


def prologue_mt(seed, csr, n, fraction_stay_home, initial_infected_fraction):
    """The reference's prologue picks for `seed` (seir_individual.py:163, 176-187) through the library's host routine
    `seir_prologue_mt`: the same (home_real bool[n], initial_infected int64[k]) as `tables.prologue` -- which replays them
    with two NumPy RandomState permutations, 0.3 s at n = 10 M -- without shuffling: the Mersenne-Twister draws are
    consumed in the reference's order and only the k picked positions are followed through the swaps.  `csr` is
    `group_csr(age_groups)` (an Engine keeps its own as `engine.group_csr`).  No device is touched."""
    lib = load_library()
    offs, members = csr
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    members = np.ascontiguousarray(members, dtype=np.int64)
    if members.shape != (n,) or int(offs[-1]) != n:
        raise ValueError("age_groups must partition the n agents")
    fsh = None
    if fraction_stay_home is not None:
        fsh = np.ascontiguousarray(fraction_stay_home, dtype=np.float64)
        if fsh.shape != (offs.shape[0] - 1,):
            raise ValueError("fraction_stay_home must have one entry per age")
        if not np.isfinite(fsh).all() or (fsh < 0).any():
            raise ValueError("fraction_stay_home must be finite and non-negative")   # (int(nan * len) raises in the reference, :181)
        if not fsh.any():
            fsh = None
    if not np.isfinite(initial_infected_fraction) or initial_infected_fraction < 0:
        raise ValueError("initial_infected_fraction must be finite and non-negative")
    k0 = int(initial_infected_fraction * n)
    home = np.empty(n, dtype=np.uint8)
    init = np.empty(max(k0, 1), dtype=np.int64)
    got = np.zeros(1, dtype=np.int64)
    _check(lib.seir_prologue_mt(int(seed) & 0xFFFFFFFF, n, offs.shape[0] - 1, _ptr(offs), _ptr(members), _ptr(fsh),
                                float(initial_infected_fraction), _ptr(home), _ptr(init), init.size, _ptr(got)))
    assert int(got[0]) == k0
    return home.view(np.bool_), init[:k0]


_tsv_buf = None


def format_tsv(a):
    """bytes: a float64 matrix as the reference's `to_csv(sep="\\t", index=False, header=False, na_rep="NA")` writes it
    (the library's host routine seir_format_tsv; sweep_io falls back on nothing -- this is its float path)."""
    global _tsv_buf
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    rows, cols = a.shape
    cap = rows * cols * 27 + rows + 1
    if _tsv_buf is None or len(_tsv_buf) < cap:
        _tsv_buf = C.create_string_buffer(max(cap, 1 << 16))
    n = C.c_int64()
    _check(load_library().seir_format_tsv(_ptr(a) if a.size else None, rows, cols, _tsv_buf, len(_tsv_buf), C.byref(n)))
    return C.string_at(_tsv_buf, n.value)


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
