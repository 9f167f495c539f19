"""Lazy stand-ins for the nine bool[T, n] history arrays and the ten float64[n] vectors `run_model` returns
(seir_individual.py:392).

The reference materialises 9 x T x n bools on the host (5.4 GB at n = 10 M, T = 60).  Here the
history stays on the device as ONE packed byte per agent-day; a `HistoryPlane` fetches only what is
asked for.  It supports what the reference's drivers do with these arrays
(run_simulation.py:448-460, simulate_agepolicies.py:617-631, parameter_sweep.py:556-603):
`X.sum(axis=1)`, `X[-1]`, `X[t]`, `X[-1, mask]`, `np.logical_and(X[t], m)`, `np.asarray(X)`.
"""
import numpy as np


class HistoryPlane:
    __array_priority__ = 100

    def __init__(self, engine, plane, replica=0, tallies=None):
        self._e, self._plane, self._rep = engine, plane, replica
        self._tallies = tallies   # int64[T, 9, 101] of the run (the per-day sums never touch the history)
        self._host = None         # packed uint8[T, n] once the run's device history had to give way
        self.shape = (engine.T, engine.n)
        self.dtype = np.dtype(np.bool_)
        self.ndim = 2
        self._row_t, self._row = None, None   # the last single row fetched (the drivers index X[-1, mask] in loops)
        self._serial = getattr(engine, "run_serial", 0)
        if hasattr(engine, "register_view"):
            engine.register_view(self)

    def __len__(self):
        return self.shape[0]

    def _rows(self, t0, t1):
        if self._host is not None:   # (two or more runs ago: served from the host copy taken when it was overwritten)
            p = self._host[t0:t1]
            k = _PLANE_INDEX[self._plane]
            comp = {0: 0, 1: 1, 2: 2, 4: 3, 5: 4, 6: 5, 7: 6}.get(k)
            return (p & 7) == comp if comp is not None else ((p >> 4) & 1) == 1 if k == 3 else ((p >> 3) & 1) == 1
        cur = getattr(self._e, "run_serial", 0)
        if cur == self._serial:
            return self._e.history(self._plane, t0, t1, self._rep)
        if getattr(self._e, "prev_serial", -1) == self._serial:
            return self._e.history(self._plane, t0, t1, self._rep, previous=True, rows=self.shape[0])
        raise RuntimeError("this history was overwritten by a later run on the same device-resident population")

    def _one_row(self, t):
        if self._row_t != t:
            self._row, self._row_t = self._rows(t, t + 1)[0], t
            self._row.setflags(write=False)
        return self._row

    def __array__(self, dtype=None, copy=None):
        a = self._rows(0, self.shape[0])
        return a if dtype is None else a.astype(dtype)

    def sum(self, axis=None, **kw):
        """Per-day counts come from the per-age tallies fused into the day step -- no history read."""
        tal = self._tallies if self._tallies is not None else self._e.tallies(self._rep)
        per_day = tal[:, _PLANE_INDEX[self._plane], :].sum(axis=1)
        if axis == 1 or axis == -1:
            return per_day
        if axis is None:
            return per_day.sum()
        return np.asarray(self).sum(axis=axis, **kw)

    def __getitem__(self, key):
        T = self.shape[0]
        if isinstance(key, tuple):
            row, rest = key[0], key[1:]
        else:
            row, rest = key, ()
        if isinstance(row, (int, np.integer)):
            t = int(row) + (T if row < 0 else 0)
            if not 0 <= t < T:
                raise IndexError("day index out of range")
            out = self._one_row(t)
            return out[rest] if rest else out
        if isinstance(row, slice):
            idx = range(*row.indices(T))
            if len(idx) == 0:
                out = np.zeros((0, self.shape[1]), dtype=np.bool_)
            elif idx.step == 1:
                out = self._rows(idx.start, idx.stop)
            else:
                out = np.stack([self._rows(t, t + 1)[0] for t in idx])
            return out[(slice(None),) + rest] if rest else out   # the rest indexes the AGENT axis
        return np.asarray(self)[key]                              # (array / mask row keys: NumPy's own semantics)

    def __iter__(self):
        for t in range(self.shape[0]):
            yield self[t]

    def __repr__(self):
        return "HistoryPlane(%s, shape=%s, device-resident)" % (self._plane, self.shape)


class AgentVector(np.lib.mixins.NDArrayOperatorsMixin):
    """Lazy stand-in for one of the ten float64[n] vectors of the reference's result tuple (seir_individual.py:392).

    The reference hands back ten fresh n-vectors per call (0.8 GB at n = 10 M: a fifth of a second of PCIe and page
    faults per seed) of which a driver reads two or three (`time_exposed`, `num_infected_by`: run_simulation.py:448-453).
    Here a vector is computed from the device-resident records and copied when it is first USED: arithmetic and
    comparisons (NumPy ufuncs), indexing, `np.asarray`, any ndarray method or attribute (`.mean()`, `.astype`, ...),
    iteration, pickling.  After that it IS a host float64 array.  Like the histories it survives the next run on the same
    device-resident population (the records of the last run are set aside on the device) and takes its host copy when
    a second run would overwrite it."""
    __array_priority__ = 100

    def __init__(self, engine, which, replica=0):
        self._e, self._which, self._rep = engine, which, replica
        self._host = None
        self.shape = (engine.n,)
        self.dtype = np.dtype(np.float64)
        self.ndim = 1
        self.size = engine.n
        self._serial = getattr(engine, "run_serial", 0)
        if hasattr(engine, "register_view"):
            engine.register_view(self)

    def _get(self):
        if self._host is None:
            cur = getattr(self._e, "run_serial", 0)
            if cur == self._serial:
                self._host = self._e.agent_vector(self._which, self._rep)
            elif getattr(self._e, "prev_serial", -1) == self._serial:
                self._host = self._e.agent_vector(self._which, self._rep, previous=True)
            else:
                raise RuntimeError("this vector was overwritten by a later run on the same device-resident population")
            self._e = None
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self._get()
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        unwrap = lambda x: x._get() if isinstance(x, AgentVector) else x  # noqa: E731
        if "out" in kwargs:
            kwargs["out"] = tuple(unwrap(x) for x in kwargs["out"])
        return getattr(ufunc, method)(*[unwrap(x) for x in inputs], **kwargs)

    def __getitem__(self, key):
        return self._get()[key]

    def __setitem__(self, key, value):
        self._get()[key] = value

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        return iter(self._get())

    def __getattr__(self, name):   # (only reached for names not set above: the ndarray's own methods and attributes)
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self._get(), name)

    def __reduce__(self):
        return (np.asarray, (self._get(),))

    def __repr__(self):
        if self._host is None:
            return "AgentVector(%s, shape=%s, device-resident)" % (self._which, self.shape)
        return repr(self._host)


_PLANE_INDEX = {"S": 0, "E": 1, "Mild": 2, "Documented": 3, "Severe": 4, "Critical": 5, "R": 6, "D": 7, "Q": 8}
