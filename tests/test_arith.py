"""include/vksift_arith.h as compiled by gcc: accuracy against numpy (float64) in ULP, exact integer helpers."""
import numpy as np


def _ulp_err(got, exact64):
    got = np.asarray(got, np.float64)
    ulp = np.spacing(np.abs(exact64).astype(np.float32)).astype(np.float64)
    return np.abs(got - exact64) / ulp


def test_expf(oracle_mod):
    a = oracle_mod.arith
    xs = np.concatenate([np.linspace(-87, 0, 4001), np.linspace(-2, 2, 2001), [-0.0, 0.0, -1e-8, -86.99]]).astype(np.float32)
    got = np.array([a.expf(x) for x in xs])
    err = _ulp_err(got, np.exp(xs.astype(np.float64)))
    assert err.max() <= 2.0, err.max()
    assert a.expf(-87.5) == 0.0 and a.expf(-1000.0) == 0.0
    assert a.expf(0.0) == 1.0


def test_exp2f(oracle_mod):
    a = oracle_mod.arith
    xs = np.linspace(-3, 4, 7001).astype(np.float32)
    got = np.array([a.exp2f(x) for x in xs])
    assert _ulp_err(got, np.exp2(xs.astype(np.float64))).max() <= 2.0
    for k in range(-5, 6):
        assert a.exp2f(float(k)) == 2.0 ** k


def test_atan2f(oracle_mod):
    a = oracle_mod.arith
    rng = np.random.default_rng(0)
    ys = rng.uniform(-1, 1, 6000).astype(np.float32)
    xs = rng.uniform(-1, 1, 6000).astype(np.float32)
    got = np.array([a.atan2f(y, x) for y, x in zip(ys, xs)])
    exact = np.arctan2(ys.astype(np.float64), xs.astype(np.float64))
    assert np.abs(got - exact).max() < 4e-7  # ~ 2 ULP at pi
    assert a.atan2f(0.0, 0.0) == 0.0
    assert a.atan2f(0.0, 1.0) == 0.0
    assert abs(a.atan2f(1.0, 0.0) - np.pi / 2) < 1e-7
    assert abs(a.atan2f(0.0, -1.0) - np.pi) < 3e-7
    assert abs(a.atan2f(-1.0, -1.0) + 3 * np.pi / 4) < 3e-7


def test_sincosf(oracle_mod):
    a = oracle_mod.arith
    ts = np.linspace(0, 6.4, 6401).astype(np.float32)
    s = np.array([a.sincosf(t) for t in ts])
    assert np.abs(s[:, 0] - np.sin(ts.astype(np.float64))).max() < 2e-7
    assert np.abs(s[:, 1] - np.cos(ts.astype(np.float64))).max() < 2e-7
    assert a.sincosf(0.0) == (0.0, 1.0)


def test_ceil_log2_and_mirror(oracle_mod):
    a = oracle_mod.arith
    for m, e in [(1.0, 0), (1.0000001, 1), (2.0, 1), (3.0, 2), (4.0, 2), (1000.0, 10), (1024.0, 10), (1025.0, 11), (0.5, -1), (0.75, 0)]:
        assert a.ceil_log2(m) == e, m
    n = 5
    assert [a.mirror(i, n) for i in range(-7, 13)] == [3, 4, 4, 3, 2, 1, 0, 0, 1, 2, 3, 4, 4, 3, 2, 1, 0, 0, 1, 2]
