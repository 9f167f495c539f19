"""The keyed-draw contract (include/seir_draws.h) evaluated on the CPU: known answers, accuracy of the
deterministic log/exp, and distribution of the variates against the reference's definitions
(functions.py:16-25, numba/cpython/randomimpl.py:966-972, 1672-1691)."""
import ctypes as C

import numpy as np
import scipy.stats as st

import oracle


def _philox(k0, k1, c):
    out = np.zeros(4, dtype=np.uint32)
    oracle.lib().oracle_philox(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(c[0]), C.c_uint32(c[1]),
                               C.c_uint32(c[2]), C.c_uint32(c[3]), C.c_void_p(out.ctypes.data))
    return [int(x) for x in out]


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _philox(0, 0, (0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _philox(f, f, (f, f, f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox(0xa4093822, 0x299f31d0, (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _apply(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    getattr(oracle.lib(), fn)(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_int64(len(x)))
    return y


def test_sd_log_accuracy():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(200000), 1.0 - rng.random(1000) * 1e-9, rng.random(1000) * 1e-300,
                        np.exp(rng.uniform(-700, 700, 20000)), [1.0, 0.5, 2.0, 2.0 ** -53, 5e-324]])
    y = _apply("oracle_sd_log", x)
    ref = np.log(x)
    err = np.abs(y - ref) / np.maximum(np.abs(ref), 1e-300)
    assert err[ref != 0].max() < 4e-16
    assert y[-5] == 0.0


def test_sd_exp_accuracy():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-20, 20, 200000), rng.uniform(-740, 709, 20000), [0.0, 1.0, -1.0]])
    y = _apply("oracle_sd_exp", x)
    ref = np.exp(x)
    normal = ref > 1e-300
    assert (np.abs(y - ref)[normal] / ref[normal]).max() < 5e-16
    assert y[-3] == 1.0


def test_exponential_days_distribution():
    rng = np.random.default_rng(2)
    u = rng.random(400000)
    for mean in (4.6, 14.0, 1e4):
        y = np.empty_like(u)
        oracle.lib().oracle_sd_exp_days(C.c_void_p(u.ctypes.data), C.c_double(mean), C.c_void_p(y.ctypes.data),
                                        C.c_int64(len(u)))
        want = np.round(-np.log(1.0 - u) * mean)  # functions.py:17 with numba's exponential
        assert (y != want).mean() < 1e-9  # identical except possibly at exact rounding ties of the last ulp
    inf = np.empty(10)
    uu = np.linspace(0.1, 0.9, 10)
    oracle.lib().oracle_sd_exp_days(C.c_void_p(uu.ctypes.data), C.c_double(np.inf), C.c_void_p(inf.ctypes.data),
                                    C.c_int64(10))
    assert np.isinf(inf).all()


def test_lognormal_days_distribution():
    m = 200000
    y = np.empty(m)
    oracle.lib().oracle_sd_lognormal_days(C.c_uint64(7), C.c_int64(m), C.c_double(1.621), C.c_double(0.418),
                                          C.c_void_p(y.ctypes.data))
    # functions.py:19-25 : round(lognormal); compare with the exact cell probabilities
    ks = np.arange(0, 40)
    cdf = st.norm.cdf((np.log(ks + 0.5) - 1.621) / 0.418)
    p = np.diff(np.concatenate([[0.0], cdf]))
    obs = np.bincount(y.astype(int), minlength=40)[:40]
    keep = p * m > 20
    chi2 = ((obs[keep] - p[keep] * m) ** 2 / (p[keep] * m)).sum()
    assert chi2 < st.chi2.ppf(1 - 1e-6, keep.sum() - 1)
    assert y.min() >= 1


def test_poisson_distribution():
    m = 300000
    for lam in (0.02, 0.4, 2.18):
        y = np.empty(m, dtype=np.int64)
        oracle.lib().oracle_sd_poisson(C.c_uint64(9), C.c_int64(m), C.c_double(np.exp(-lam)),
                                       C.c_void_p(y.ctypes.data))
        ks = np.arange(0, 12)
        p = st.poisson.pmf(ks, lam)
        obs = np.bincount(y, minlength=12)[:12]
        keep = p * m > 20
        chi2 = ((obs[keep] - p[keep] * m) ** 2 / (p[keep] * m)).sum()
        assert chi2 < st.chi2.ppf(1 - 1e-6, max(1, keep.sum() - 1)), lam
