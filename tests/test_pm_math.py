"""Accuracy and special values of the portable math header used by the bit-exact tier
(clode_b200/csrc/device/pm_math.h) against numpy/glibc.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def pm():
    out = os.path.join(REPO, "oracle", "_build", "pm_probe.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", f"-I{REPO}/clode_b200/csrc/device",
                    os.path.join(HERE, "emu", "pm_probe.c"), "-o", out], check=True)
    return ctypes.CDLL(out)


def _call1(fn, x):
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty_like(x)
    fn(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), x.size)
    return y


def _ulp_err(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_exp_log_sin_cos_accuracy(pm):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-1, 1, 100000), rng.normal(0, 1e-8, 1000)])
    assert _ulp_err(_call1(pm.probe_exp, x), np.exp(x)).max() <= 2.0
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0.5, 2.0, 100000), [5e-324, 1e-310]])
    want = np.log(x)
    err = np.abs(_call1(pm.probe_log, x) - want) / np.maximum(np.spacing(np.abs(want)), 1e-300)
    assert err.max() <= 2.0
    x = np.concatenate([rng.uniform(-50, 50, 200000), rng.uniform(-1e4, 1e4, 50000)])
    # absolute-error bound near zeros of sin/cos (Cody-Waite reduction keeps ~1e-16 * |n| absolute error)
    for fn, ref in ((pm.probe_sin, np.sin), (pm.probe_cos, np.cos)):
        got, want = _call1(fn, x), ref(x)
        assert np.max(np.abs(got - want)) <= 4e-16
        big = np.abs(want) > 0.1
        assert _ulp_err(got[big], want[big]).max() <= 2.0


def test_pow_accuracy_inside_opencl_bound(pm):
    """pow(x, y) error <= 2 + |y ln x| ulp; OpenCL C 1.2 allows 16 ulp for pow"""
    rng = np.random.default_rng(1)
    x = 10.0 ** rng.uniform(-12, 12, 200000)
    y = rng.uniform(-3, 3, x.size)
    got = np.empty_like(x)
    pm.probe_pow(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p),
                 got.ctypes.data_as(ctypes.c_void_p), x.size)
    want = np.power(x, y)
    err = _ulp_err(got, want)
    assert np.all(err <= 2.0 + np.abs(y * np.log(x)))
    ctrl = np.abs(y * np.log(x)) < 5  # the step-size controller's operating range
    assert err[ctrl].max() <= 7.0


def test_controller_root_path(pm):
    """pow(x, 1/5) and pow(x, 1/3) — the step-size controller's only calls (adaptive_explicit_step.clh:51,66) — take the
    multiplication-only root (pm_rootq): <= 11 ulp / <= 5 ulp over every binade, exact at perfect powers, the general
    exp(y log x) path for every other exponent"""
    rng = np.random.default_rng(2)
    x = np.concatenate([2.0 ** rng.uniform(-1070, 1023, 400000), 10.0 ** rng.uniform(-8, 8, 100000), [5e-324, 1e-310, 1.0, 32.0, 243.0]])
    for y, bound in ((0.2, 11.0), (1.0 / 3.0, 5.0)):
        got = np.empty_like(x)
        yy = np.full(x.size, y)
        pm.probe_pow(x.ctypes.data_as(ctypes.c_void_p), yy.ctypes.data_as(ctypes.c_void_p), got.ctypes.data_as(ctypes.c_void_p), x.size)
        want = np.power(x.astype(np.longdouble), np.longdouble(1) / np.longdouble(round(1 / y))).astype(np.float64)
        assert _ulp_err(got, want).max() <= bound
    one = lambda a, b: (lambda o: (pm.probe_pow(np.array([a]).ctypes.data_as(ctypes.c_void_p), np.array([b]).ctypes.data_as(ctypes.c_void_p),
                                               o.ctypes.data_as(ctypes.c_void_p), 1), o[0])[1])(np.empty(1))
    assert one(32.0, 0.2) == 2.0 and one(27.0, 1.0 / 3.0) == 3.0 and one(1.0, 0.2) == 1.0
    assert one(np.inf, 0.2) == np.inf and one(0.0, 1.0 / 3.0) == 0.0 and np.isnan(one(-8.0, 1.0 / 3.0))


def test_special_values(pm):
    inf, nan = np.inf, np.nan
    assert np.array_equal(_call1(pm.probe_exp, [0.0, -inf, inf, 710.0, -750.0]), [1.0, 0.0, inf, inf, 0.0])
    assert np.isnan(_call1(pm.probe_exp, [nan]))[0]
    got = _call1(pm.probe_log, [1.0, 0.0, -0.0, inf])
    assert np.array_equal(got, [0.0, -inf, -inf, inf])
    assert np.isnan(_call1(pm.probe_log, [-1.0, nan])).all()

    def p(a, b):
        a, b = np.array([a], np.float64), np.array([b], np.float64)
        out = np.empty(1)
        pm.probe_pow(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                     out.ctypes.data_as(ctypes.c_void_p), 1)
        return out[0]

    assert p(2.0, 0.0) == 1.0 and p(nan, 0.0) == 1.0 and p(1.0, nan) == 1.0
    assert p(inf, 0.2) == inf and p(0.0, 0.2) == 0.0 and p(0.0, -1.0) == inf
    assert p(-2.0, 3.0) == pytest.approx(-8.0) and p(-2.0, 2.0) == pytest.approx(4.0) and np.isnan(p(-2.0, 0.5))
    assert p(4.0, 0.5) == pytest.approx(2.0, rel=1e-15) and p(2.0, 10.0) == pytest.approx(1024.0, rel=1e-15)


def test_step_floor_exponent_identity():
    """steppers.cuh step_floor(), production build: for 0 < t <= t_end, t normal and >= 2^-974,
    16*| |nextafter(t, 1.1 t_end)| - t |  (adaptive_explicit_step.clh:17) == 2^(E-1071), E = biased exponent of t,
    built as hi' = (hi & 0x7ff00000) - (48 << 20), lo' = 0."""
    rng = np.random.default_rng(7)
    t = np.concatenate([
        np.exp(rng.uniform(np.log(2.0 ** -974), np.log(1e300), 20000)),
        2.0 ** rng.integers(-974, 1000, 2000).astype(np.float64),          # exact powers of two
        np.nextafter(2.0 ** rng.integers(-973, 1000, 2000).astype(np.float64), 0.0),  # all-ones mantissas
        [2.0 ** -974, 1e-3, 0.01, 100.0, 1e4],
    ])
    t_end = t * np.concatenate([rng.uniform(1.0, 3.0, t.size - 5), [1.0, 1.0, 1.0, 1.0, 1.0]])
    t_end = np.maximum(t_end, t)
    want = 16.0 * np.abs(np.abs(np.nextafter(t, 1.1 * t_end)) - t)
    bits = t.view(np.uint64)
    hi = (bits >> np.uint64(32)).astype(np.int64)
    assert np.all((hi >= (49 << 20)) & (hi < 0x7FF00000))
    got = (((hi & 0x7FF00000) - (48 << 20)).astype(np.uint64) << np.uint64(32)).view(np.float64)
    assert np.array_equal(got, want)


def test_attempt_clamp_on_high_words_identity():
    """steppers.cuh adaptive_attempt(), production double: with hmin = 16 ulp(t) = {hi', 0} (a power of two, zero low word),
    h <= dtmax (loop invariant) and hi(t) < floor_hi_max (which encodes 16 ulp(t) <= dtmax),
        clamp(h, hmin, dtmax) = fmin(fmax(h, hmin), dtmax)   (adaptive_explicit_step.clh:31)
    equals  `hi(h) < hi' ? hmin : h`  — an integer comparison of high words."""
    rng = np.random.default_rng(21)
    n = 200000
    t = np.exp(rng.uniform(np.log(1e-6), np.log(1e12), n))
    dtmax = np.exp(rng.uniform(np.log(1e-9), np.log(1e6), n))
    # h anywhere from far below hmin to dtmax, negative and zero values included; a share sits within a few ulp of hmin
    hi_t = (t.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    f_hi = (hi_t & 0x7FF00000) - (48 << 20)
    hmin = (f_hi.astype(np.uint64) << np.uint64(32)).view(np.float64)
    h = np.minimum(np.exp(rng.uniform(np.log(1e-30), np.log(1e6), n)), dtmax)
    near = rng.random(n) < 0.3
    h[near] = np.minimum((hmin * (1.0 + rng.integers(-4, 5, n) * 2.0 ** -52))[near], dtmax[near])
    h[rng.random(n) < 0.02] = 0.0
    h[rng.random(n) < 0.02] *= -1.0
    d_hi = (dtmax.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    floor_hi_max = np.minimum(0x7FF00000, (d_hi & 0x7FF00000) + (49 << 20))
    fast = (hi_t >= (49 << 20)) & (hi_t < floor_hi_max)
    assert fast.mean() > 0.5 and (~fast).sum() > 100      # both sides of the 16 ulp(t) <= dtmax condition are sampled
    assert np.all(hmin[fast] <= dtmax[fast])
    want = np.minimum(np.maximum(h, hmin), dtmax)
    hi_h = (h.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    hi_h = np.where(hi_h >= 1 << 31, hi_h - (1 << 32), hi_h)  # as the signed 32-bit word the device compares
    got = np.where(hi_h < f_hi, hmin, h)
    assert np.array_equal(got[fast], want[fast])
