"""CPU tier: `history.AgentVector`, the lazy stand-in of the ten float64[n] result vectors (seir_individual.py:392),
behaves like the ndarray it becomes -- the drivers' statements (run_simulation.py:448-453, parameter_sweep.py:556-567),
NumPy functions, operators, methods, pickling -- fetches at most once, and follows the engine's hand-over protocol
(current run -> set aside on the device -> host copy) driven here by a stand-in engine."""
import pickle

import numpy as np
import pytest

from covid19_demography_b200 import engine as engine_mod
from covid19_demography_b200.history import AgentVector


class FakeEngine:
    """The part of engine.Engine an AgentVector touches; `vectors[serial]` is what run `serial` produced."""
    _keep_live_histories = engine_mod.Engine._keep_live_histories
    register_view = engine_mod.Engine.register_view
    keep_views_on_host = engine_mod.Engine.keep_views_on_host

    def __init__(self, n, history_mode="eager"):
        self.n, self.T = n, 5
        self.run_serial, self.prev_serial, self._views = 0, -1, []
        self.history_mode = history_mode
        self.vectors, self.set_aside, self.fetches, self.preserves = {}, None, [], 0
        outer = self

        class _Lib:
            @staticmethod
            def seir_history_preserve(h):
                outer.set_aside = outer.vectors[outer.run_serial]
                outer.preserves += 1
                return 0
        self.lib, self._h = _Lib, None

    def run(self, values):
        self._keep_live_histories()
        self.run_serial += 1
        self.vectors[self.run_serial] = values

    def agent_vector(self, which, replica=0, previous=False):
        self.fetches.append((which, previous))
        src = self.set_aside if previous else self.vectors[self.run_serial]
        return np.array(src[which], dtype=np.float64)

    def history(self, *a, **k):
        raise AssertionError("no history plane is alive in these tests")


def _values(seed, n):
    rng = np.random.default_rng(seed)
    te = np.where(rng.random(n) < 0.3, rng.integers(0, 20, n), -1).astype(np.float64)
    nib = np.where(te >= 0, rng.integers(0, 5, n), -1).astype(np.float64)
    return {"time_exposed": te, "num_infected_by": nib}


def test_behaves_like_the_array_it_becomes():
    n = 500
    e = FakeEngine(n)
    vals = _values(1, n)
    e.run(vals)
    te, nib = AgentVector(e, "time_exposed"), AgentVector(e, "num_infected_by")
    assert e.fetches == [] and te.shape == (n,) and len(te) == n and te.dtype == np.float64 and te.ndim == 1
    assert "device-resident" in repr(te)
    want_te, want_nib = vals["time_exposed"], vals["num_infected_by"]
    age = np.random.default_rng(0).integers(0, 101, n)
    # the drivers' statements
    for t in range(20):
        if (te == t).sum() > 0:
            assert nib[te == t].mean() == want_nib[want_te == t].mean()
            assert np.median(age[te == t]) == np.median(age[want_te == t])
    assert nib[np.logical_and(te <= 20, te > 0)].mean() == want_nib[np.logical_and(want_te <= 20, want_te > 0)].mean()
    assert e.fetches == [("time_exposed", False), ("num_infected_by", False)]   # one fetch each, however often used
    # operators, reflected operators, ufuncs, NumPy functions, methods, attributes
    assert np.array_equal(te + 1, want_te + 1) and np.array_equal(2 * te, 2 * want_te) and np.array_equal(-te, -want_te)
    assert np.array_equal(te >= 0, want_te >= 0) and np.array_equal(np.maximum(te, nib), np.maximum(want_te, want_nib))
    assert np.array_equal(np.where(te >= 0)[0], np.where(want_te >= 0)[0]) and np.sum(te) == want_te.sum()
    assert te.mean() == want_te.mean() and te.max() == want_te.max() and te.astype(np.int64).dtype == np.int64
    assert te.T.shape == (n,) and te.nbytes == n * 8 and te.tolist() == want_te.tolist()
    assert np.array_equal(np.asarray(te), want_te) and np.asarray(te, dtype=np.float32).dtype == np.float32
    assert np.array_equal(te[::7], want_te[::7]) and te[3] == want_te[3] and list(te)[:5] == list(want_te[:5])
    assert np.array_equal(np.concatenate([te, nib]), np.concatenate([want_te, want_nib]))
    assert np.array_equal(np.bincount(age[te >= 0], minlength=101), np.bincount(age[want_te >= 0], minlength=101))
    assert np.array_equal(pickle.loads(pickle.dumps(te)), want_te) and isinstance(pickle.loads(pickle.dumps(te)), np.ndarray)
    te[0] = 42.0   # the reference's vectors are the caller's to modify
    assert te[0] == 42.0 and np.asarray(te)[0] == 42.0
    out = np.empty(n)
    np.add(te, 1.0, out=out)
    assert out[0] == 43.0
    with pytest.raises(AttributeError):
        te._no_such_private_name


def test_vectors_of_earlier_runs_keep_their_own_content():
    n = 300
    e = FakeEngine(n)
    runs, held = {}, {}
    for serial in (1, 2, 3, 4):
        runs[serial] = _values(serial, n)
        e.run(runs[serial])
        held[serial] = (AgentVector(e, "time_exposed"), AgentVector(e, "num_infected_by"))
        if serial == 2:
            assert np.array_equal(held[2][0], runs[2]["time_exposed"])   # used while current: plain fetch
    # nothing was fetched that nobody used, except what a second run would have overwritten
    assert e.preserves == 3
    for serial in (1, 2, 3, 4):
        assert np.array_equal(held[serial][0], runs[serial]["time_exposed"]), serial
        assert np.array_equal(held[serial][1], runs[serial]["num_infected_by"]), serial
    assert ("num_infected_by", True) in e.fetches and ("time_exposed", False) in e.fetches
    # vectors nobody holds cost nothing
    e2 = FakeEngine(n)
    for serial in (1, 2, 3):
        e2.run(_values(serial, n))
        AgentVector(e2, "time_exposed")
    assert e2.fetches == [] and e2.preserves == 0


def test_history_on_demand_engines_take_host_copies_before_the_next_run():
    n = 100
    e = FakeEngine(n, history_mode="on_demand")
    a = _values(1, n)
    e.run(a)
    v = AgentVector(e, "time_exposed")
    e.run(_values(2, n))
    assert e.preserves == 0 and e.fetches == [("time_exposed", False)]
    assert np.array_equal(v, a["time_exposed"])


def test_vectors_outlive_their_engine():
    n = 200
    e = FakeEngine(n)
    a, b = _values(1, n), _values(2, n)
    e.run(a)
    va = AgentVector(e, "time_exposed")
    e.run(b)
    vb = AgentVector(e, "num_infected_by")
    e.keep_views_on_host()          # what the drop-in module does before it closes an evicted engine
    e.vectors, e.set_aside = None, None
    assert np.array_equal(va, a["time_exposed"]) and np.array_equal(vb, b["num_infected_by"])
