"""Drop-in replacement of the reference's `seir_individual` module for its hot path.

Same two entry points, same positional arguments, same 20-tuple result
(/root/reference/seir_individual.py:11 `run_complete_simulation`, :91 `run_model`, return at :392),
so `run_simulation.py`, `simulate_agepolicies.py` and `parameter_sweep.py`-style drivers keep working
with this directory ahead of the reference on `sys.path`.  The day loop (:231-390) runs as CUDA
kernels on a B200 through libseir_b200.so; there is no CPU fallback.

Differences a caller can see (DESIGN.md has the full list):
  * the nine histories come back as lazy `HistoryPlane` objects over the device-resident packed
    history (`np.asarray(X)` materialises the reference's bool[T, n]), the ten per-agent vectors as lazy
    `AgentVector` objects that turn into host float64 arrays when first used (`SEIR_B200_VECTORS=eager`:
    plain ndarrays at once);
  * draws inside the day loop are keyed Philox draws (include/seir_draws.h), not the reference's
    single Mersenne-Twister stream, so trajectories agree with the reference in distribution,
    not per seed; Home_real and initial_infected ARE the reference's for the same seed;
  * contact tracing (`do_contact_tracing`, :60-88) runs on the device with the reference's sequential
    semantics.  As compiled by numba 0.65 the reference switches tracing on at t_tracing_start
    whatever params['contact_tracing'] says; that is the default here too
    (tracing="as_reference").  `SEIR_B200_TRACING=off` (or tracing="off") runs the paper's
    no-tracing semantics.
"""
import os
import pickle
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))

from covid19_demography_b200 import engine as _engine  # noqa: E402
from covid19_demography_b200 import tables as _tables  # noqa: E402
from covid19_demography_b200.history import AgentVector, HistoryPlane  # noqa: E402

OPTIONS = {
    "mode": os.environ.get("SEIR_B200_MODE", "native"),          # "native" | "replay"
    "launch": os.environ.get("SEIR_B200_LAUNCH", "persistent"),  # "persistent" | "per_day"
    "tracing": os.environ.get("SEIR_B200_TRACING", "as_reference"),  # "as_reference" | "off"
    "device": int(os.environ.get("SEIR_B200_DEVICE", "0")),
    # the ten float64[n] result vectors: "lazy" (computed and copied when first used, history.AgentVector) | "eager"
    "vectors": os.environ.get("SEIR_B200_VECTORS", "lazy"),
    # the reference's Mersenne-Twister picks of Home_real / initial_infected (:176-187): "native" (the library's host
    # routine seir_prologue_mt) | "numpy" (tables.prologue: two RandomState permutations) -- the same picks either way
    "prologue": os.environ.get("SEIR_B200_PROLOGUE", "native"),
}
last_engine = None
"""The Engine of the most recent run_model call (its histories stay valid until the next call on the same population)."""
_engine_cache = {}
ENGINE_CACHE_SIZE = int(os.environ.get("SEIR_B200_ENGINE_CACHE", "2"))


def _population_key(households, age, diabetes, hypertension, p_infect_household, opt):
    """Identity of a device-resident population: shapes plus a strided checksum of every array (the drivers pass the
    same arrays seed after seed -- run_simulation.py:437-444 -- or reload the same pickle, parameter_sweep.py:29).
    A few hundred kB are read: microseconds against the 1.5 s an engine takes to create at n = 10 M."""
    import zlib

    def sig(a, samples):
        a = np.asarray(a)
        flat = a.reshape(-1)
        step = max(1, flat.shape[0] // samples)
        return (a.shape, str(a.dtype), zlib.crc32(np.ascontiguousarray(flat[::step]).tobytes()),
                zlib.crc32(np.ascontiguousarray(flat[-4096:]).tobytes()))
    arrays = (households, age, diabetes, hypertension, p_infect_household)
    # the same array objects at the same addresses as in the last call (the usual case): a sixteenth of the samples
    # confirms the key that call computed
    ident = tuple((id(a), np.asarray(a).__array_interface__["data"][0], np.asarray(a).shape) for a in arrays)
    light = tuple(sig(a, 4096) for a in arrays)
    global _last_population
    if _last_population is not None and _last_population[0] == ident and _last_population[1] == light:
        full = _last_population[2]
    else:
        full = tuple(sig(a, 65536) for a in arrays)
        _last_population = (ident, light, full)
    return full + (opt["mode"], opt["launch"], opt["device"])


_last_population = None


_tables_cache = {}


def _tables_for(contact_matrix, group_sizes, p, iso_factor, p_ms, p_sc, p_cd, with_replay):
    """tables.build_tables, remembered by the CONTENT of its inputs: the drivers call run_model seed after seed with the
    same matrices and parameters (run_simulation.py:437-444), and the alias tables are a Python loop of ten milliseconds."""
    import zlib
    arrays = [np.ascontiguousarray(a) for a in (contact_matrix, group_sizes, iso_factor, p_ms, p_sc, p_cd)]
    key = (tuple((a.shape, str(a.dtype), zlib.crc32(a.tobytes()), zlib.adler32(a.tobytes())) for a in arrays),
           tuple(sorted(p.items())), with_replay)
    tb = _tables_cache.get(key)
    if tb is None:
        if len(_tables_cache) >= 8:
            _tables_cache.pop(next(iter(_tables_cache)))
        tb = _tables_cache[key] = _tables.build_tables(contact_matrix, group_sizes, p, iso_factor, p_ms, p_sc, p_cd,
                                                       with_replay=with_replay)
    return tb


def _engine_for(key, build):
    """Least-recently-used cache of engines (an engine owns O(100 B/agent) of HBM: keep few)."""
    eng = _engine_cache.pop(key, None)
    fresh = eng is None
    if fresh and ENGINE_CACHE_SIZE <= 0:
        return build(), True   # no cache: every call owns its engine (earlier histories stay readable)
    if fresh:
        while len(_engine_cache) >= max(1, ENGINE_CACHE_SIZE):
            old = _engine_cache.pop(next(iter(_engine_cache)))
            old.keep_views_on_host()   # results of that population somebody still holds stay readable
            old.close()
        eng = build()
    _engine_cache[key] = eng
    return eng, fresh


def clear_engine_cache():
    global last_engine
    for eng in _engine_cache.values():
        eng.keep_views_on_host()
        eng.close()
    _engine_cache.clear()
    last_engine = None


def run_complete_simulation(seed, country, contact_matrix, p_mild_severe, p_severe_critical, p_critical_death,
                            mean_time_to_isolate_factor, lockdown_factor_age, p_infect_household,
                            fraction_stay_home, params, load_population=False):
    """seir_individual.py:11-44: build or load the population, then `run_model`.

    Keeps the side effect drivers rely on: `{country}_population_{n}.pickle` in the CWD holds
    `(age, households, diabetes, hypertension, age_groups)` (:22, :39-41).
    """
    n = int(params['n'])
    n_ages = int(params['n_ages'])
    np.random.seed(seed)
    fname = '{}_population_{}.pickle'.format(country, n)
    if load_population:
        print('loading')
        with open(fname, 'rb') as f:
            age, households, diabetes, hypertension, age_groups = pickle.load(f)
    else:
        from covid19_demography_b200 import population as _population
        households, age = _population.sample_households(country, n)
        age_groups = tuple([np.where(age == i)[0] for i in range(0, n_ages)])
        diabetes, hypertension = _population.sample_comorbidities(age, country)
        with open(fname, 'wb') as f:
            pickle.dump((age, households, diabetes, hypertension, age_groups), f)
    print('starting simulation')
    return run_model(seed, households, age, age_groups, diabetes, hypertension, contact_matrix, p_mild_severe,
                     p_severe_critical, p_critical_death, mean_time_to_isolate_factor, lockdown_factor_age,
                     p_infect_household, fraction_stay_home, params)


def run_model(seed, households, age, age_groups, diabetes, hypertension, contact_matrix, p_mild_severe,
              p_severe_critical, p_critical_death, mean_time_to_isolate_factor, lockdown_factor_age,
              p_infect_household, fraction_stay_home, params, **options):
    """seir_individual.py:90-392 on the GPU.  Returns the reference's 20-tuple:
    S, E, Mild, Documented, Severe, Critical, R, D, Q (HistoryPlane, bool[T, n] on demand),
    num_infected_by, time_documented, time_to_activation, time_to_death, time_to_recovery,
    time_critical, time_exposed, num_infected_asympt, age, time_infected, time_to_severe."""
    global last_engine
    opt = dict(OPTIONS)
    opt.update(options)
    p = _tables.params_to_dict(params)
    n, T = int(p["n"]), int(p["T"])
    age = np.asarray(age)
    if len(age_groups) != _tables.N_AGES or int(p["n_ages"]) != _tables.N_AGES:
        raise ValueError("n_ages must be %d" % _tables.N_AGES)
    tracing_enabled = 0 if opt["tracing"] == "off" else 1
    group_sizes = np.array([len(g) for g in age_groups], dtype=np.int64)
    tb = _tables_for(contact_matrix, group_sizes, p, mean_time_to_isolate_factor, p_mild_severe, p_severe_critical,
                     p_critical_death, opt["mode"] == "replay")
    # the population stays resident on the device between calls (the drivers call once per seed with the same
    # arrays): only the scalars, the tables and the per-run picks travel
    key = _population_key(households, age, diabetes, hypertension, p_infect_household, opt)
    eng, fresh = _engine_for(key, lambda: _engine.Engine(
        p, households, age, age_groups, diabetes, hypertension, p_infect_household, tb, n_replicas=1,
        device=opt["device"], mode=opt["mode"], launch=opt["launch"], tracing_enabled=tracing_enabled))
    if not fresh:
        eng.set_params(p, tb, tracing_enabled=tracing_enabled)
    if opt["prologue"] == "numpy":
        home_real, initial_infected = _tables.prologue(seed, age_groups, n, fraction_stay_home,
                                                       p["initial_infected_fraction"])
    else:
        home_real, initial_infected = _engine.prologue_mt(seed, eng.group_csr, n, fraction_stay_home,
                                                          p["initial_infected_fraction"])
    print('run_model')
    eng.run([seed], [home_real], [initial_infected])
    last_engine = eng
    tal = eng.tallies(0)
    planes = tuple(HistoryPlane(eng, k, tallies=tal) for k in _engine.PLANES_IN_ORDER)
    if opt["vectors"] == "eager":
        v = eng.agent_vector
    else:
        v = lambda which: AgentVector(eng, which)  # noqa: E731
    return planes + (v("num_infected_by"), v("time_documented"), v("time_to_activation"), v("time_to_death"),
                     v("time_to_recovery"), v("time_critical"), v("time_exposed"), v("num_infected_asympt"),
                     age, v("time_infected"), v("time_to_severe"))
