"""CPU tier: the C-ABI library loads and exports every symbol include/seir_b200.h declares, the ctypes
structs match the header, and the host-side table / history / parameter logic behaves like the
reference's (`seir_individual.py:46-51`, `:135-158`).  No compute call is made (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from covid19_demography_b200 import build, engine, tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "seir_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(seir_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 18
    lib = C.CDLL(build.build())
    for name in names:
        assert hasattr(lib, name), "libseir_b200.so does not export %s" % name
    assert set(engine.EXPORTS) <= set(names)
    lib.seir_build_info.restype = C.c_char_p
    assert b"sm_100a" in lib.seir_build_info()


def test_struct_sizes_match_the_header():
    # the ctypes mirrors against sizeof() as the library was compiled, and against the header by hand:
    # seir_params: int64 + 8 x int32 + 13 doubles ; seir_population: int64 + 2 x int32 + 7 pointers ;
    # seir_tables: 10 pointers ; seir_replica: uint64 + 2 pointers + int64 + pointer + double + 2 x int32 ;
    # seir_overrides: 6 x int32
    lib = engine.load_library()
    sizes = (C.c_int64 * 5)()
    lib.seir_abi_sizes(sizes)
    mine = [C.sizeof(x) for x in (engine.SeirParams, engine.SeirPopulation, engine.SeirTables, engine.SeirReplica,
                                  engine.SeirOverrides)]
    assert list(sizes) == mine == [8 + 8 * 4 + 13 * 8, 8 + 8 + 7 * 8, 10 * 8, 8 + 8 + 8 + 8 + 8 + 8 + 8, 6 * 4]


def test_shard_handle_count_matches_the_header():
    # seir_shard_export writes SEIR_SHARD_HANDLES x 64 bytes; the ctypes side sizes its buffers from its own constant
    m = re.search(r"#define\s+SEIR_SHARD_HANDLES\s+(\d+)", open(HEADER).read())
    assert m and int(m.group(1)) == engine.SHARD_HANDLES


def test_flag_constants_match_the_header():
    src = open(HEADER).read()
    m = re.search(r"#define\s+SEIR_VEC_PREVIOUS\s+(\d+)", src)
    assert m and int(m.group(1)) == engine.VEC_PREVIOUS
    m = re.search(r"SEIR_VEC_COUNT\s*=\s*(\d+)", src)
    assert m and int(m.group(1)) == len(engine.VEC) and engine.VEC_PREVIOUS > max(engine.VEC.values())
    m = re.search(r"#define\s+SEIR_H_PREVIOUS\s+(\d+)", src)
    assert m and int(m.group(1)) == 16


def test_no_device_is_an_error_not_a_fallback():
    lib = engine.load_library()
    if lib.seir_device_count() > 0:
        pytest.skip("a CUDA device is present")
    c = util.Case("italy_default")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.gpu_engine(engine, c.no_tracing_params(), "native")


def test_isolation_factor_lut_is_first_match():
    rows = np.array([[0, 14, 1], [14, 24, 3], [25, 39, 5], [40, 69, 7], [70, 100, 9]])
    lut = tables.isolation_factor_lut(rows)
    assert lut[14] == 1 and lut[15] == 3 and lut[24] == 3 and lut[25] == 5 and lut[100] == 9  # :46-51
    assert tables.isolation_factor_lut(np.array([[5, 10, 2]]))[[4, 5, 10, 11]].tolist() == [1, 2, 2, 1]


def test_params_accepts_any_mapping_and_names_the_missing_key():
    c = util.Case("italy_default")
    d = tables.params_to_dict(c.params)
    assert set(d) == set(tables.PARAM_KEYS) and d["n"] == c.n
    bad = dict(c.params)
    del bad["t_lockdown"]
    with pytest.raises(KeyError, match="t_lockdown"):
        tables.params_to_dict(bad)


def test_alias_table_reproduces_the_row_distribution():
    c = util.Case("italy_default")
    row = c.contact_matrix[40].copy()
    row[[3, 99]] = 0.0
    prob, alias = tables.alias_table(row)
    K = len(row)
    mass = np.zeros(K)
    for j in range(K):
        mass[j] += prob[j] / K
        mass[alias[j]] += (1.0 - prob[j]) / K
    assert np.allclose(mass, row / row.sum(), atol=1e-12)
    assert mass[3] == 0 and mass[99] == 0


def test_negative_contact_rate_raises_like_numba():
    c = util.Case("italy_default")
    cm = c.contact_matrix.copy()
    cm[0, 0] = -1.0
    with pytest.raises(ValueError, match="lambda < 0"):
        tables.build_tables(cm, np.ones(101), tables.params_to_dict(c.params), c.iso_factor, c.p_mild_severe,
                            c.p_severe_critical, c.p_critical_death)


def test_history_plane_indexing_matches_numpy():
    from covid19_demography_b200.history import HistoryPlane

    class FakeEngine:
        T, n = 7, 5
        rng = np.random.default_rng(0)
        full = rng.random((7, 5)) < 0.4

        def history(self, plane, t0, t1, rep):
            return self.full[t0:t1].copy()

        def tallies(self, rep):
            out = np.zeros((7, 9, 101), dtype=np.int64)
            out[:, 2, 0] = self.full.sum(axis=1)
            return out

    e = FakeEngine()
    X = HistoryPlane(e, "Mild")
    mask = np.array([True, False, True, True, False])
    assert np.array_equal(np.asarray(X), e.full)
    assert np.array_equal(X[-1], e.full[-1]) and np.array_equal(X[3], e.full[3])
    assert np.array_equal(X[-1, mask], e.full[-1, mask])
    assert np.array_equal(X.sum(axis=1), e.full.sum(axis=1))
    assert np.array_equal(np.logical_and(X[2], mask), np.logical_and(e.full[2], mask))
    assert np.array_equal(X[1:4], e.full[1:4]) and len(X) == 7
    # 2-D keys whose row part is a slice or an array index the AGENT axis with the rest, as an ndarray does
    idx = np.array([4, 0, 2])
    assert np.array_equal(X[:, mask], e.full[:, mask]) and np.array_equal(X[1:6:2, idx], e.full[1:6:2, idx])
    assert np.array_equal(X[2:5, 1], e.full[2:5, 1]) and np.array_equal(X[[0, 6], :], e.full[[0, 6], :])
    assert np.array_equal(X[np.array([1, 3]), 2], e.full[np.array([1, 3]), 2])
    # the last row is cached: the drivers' loop `X[-1, mask_k]` for k in ... fetches it once
    calls = []
    orig = e.history
    e.history = lambda *a: (calls.append(a), orig(*a))[1]
    for k in range(5):
        X[-1, mask]
        X[-1]
    assert len(calls) <= 1
    with pytest.raises(IndexError):
        X[7]


def test_fast_prologue_draws_the_same_counts():
    c = util.Case("italy_policy")
    home_ref, init_ref = tables.prologue(5, c.age_groups, c.n, c.fraction_stay_home, 0.004)
    home, init = tables.prologue_fast(5, c.age_groups, c.n, c.fraction_stay_home, 0.004)
    assert len(init) == len(init_ref) == int(0.004 * c.n) and len(set(init.tolist())) == len(init)
    for a, g in enumerate(c.age_groups):
        assert home[g].sum() == home_ref[g].sum() == int(c.fraction_stay_home[a] * len(g))
    h2, i2 = tables.prologue_fast(5, c.age_groups, c.n, c.fraction_stay_home, 0.004)
    assert np.array_equal(home, h2) and np.array_equal(init, i2)
    assert not np.array_equal(tables.prologue_fast(6, c.age_groups, c.n, c.fraction_stay_home, 0.004)[1], init)
