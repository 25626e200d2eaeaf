"""Accuracy of the product's own FP64 device math (airice_math.cuh): table-driven log, MUFU-seeded sqrt / reciprocal /
division.  The parity tolerances (1e-9 relative on distances and times, 1e-7 deg on angles, identical solution flags)
rest on these staying within ~1 ulp."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("math") / "libmathprobe.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler",
                           "-fPIC", "-I" + os.path.join(ROOT, "airiceraytracing_b200", "csrc"), "-o", so,
                           os.path.join(ROOT, "tests", "compat", "math_probe.cu")])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.math_probe.argtypes = [C.c_int, C.c_long, dp, dp, dp]

    def run(op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(a if b is None else b, dtype=np.float64)
        out = np.empty_like(a)
        assert lib.math_probe(op, a.size, a.ctypes.data_as(dp), b.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
        return out
    return run


def ulps(got, exact):
    """error in units of the last place of the exact value (exact: longdouble, 64-bit mantissa)"""
    ex = exact.astype(np.float64)
    return np.abs((got.astype(np.longdouble) - exact) / np.spacing(np.abs(ex)).astype(np.longdouble)).astype(np.float64)


@pytest.mark.gpu
def test_device_log(probe):
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-40, 40, 200000)), rng.uniform(0.5, 2.0, 200000), 1.0 + rng.uniform(-0.02, 0.02, 100000),
                        1.0 + 10.0 ** rng.uniform(-15, -3, 50000) * rng.choice([-1, 1], 50000),
                        [1.0, 0.6875, 1.375, np.nextafter(1.0, 0), np.nextafter(1.0, 2), 2.0, 0.5, 1e-300, 1e300]])
    got = probe(0, x)
    exact = np.log(x.astype(np.longdouble))
    assert got[x == 1.0][0] == 0.0                      # exact zero: zero-thickness segments rely on it
    far = np.abs(x - 1.0) > 0.02
    assert ulps(got[far], exact[far]).max() <= 1.5
    near = ~far
    # next to 1 the result is small; what the ray integrals need there is absolute accuracy (and relative accuracy in
    # the two table cells that touch 1)
    assert np.abs(got[near].astype(np.longdouble) - exact[near]).max() <= 1e-17
    touch = (x >= 0.99609375) & (x < 1.0078125)
    assert ulps(got[touch], exact[touch]).max() <= 1.0
    assert np.isnan(probe(0, np.array([-1.0, 0.0, np.nan]))).all()


@pytest.mark.gpu
def test_device_sqrt_rcp_div(probe):
    rng = np.random.default_rng(2)
    x = np.concatenate([np.exp(rng.uniform(-200, 200, 200000)), rng.uniform(0.5, 2.0, 100000), [1.0, 4.0, 2.0, 1e-300, 1e300]])
    s = probe(1, x)
    assert ulps(s, np.sqrt(x.astype(np.longdouble))).max() <= 1.0
    assert probe(1, np.array([0.0]))[0] == 0.0 and np.isnan(probe(1, np.array([-1.0]))[0])
    xr = np.concatenate([np.exp(rng.uniform(-200, 200, 200000)) * rng.choice([-1, 1], 200000), [1.0, -2.0, 3.0]])
    assert ulps(probe(2, xr), 1.0 / xr.astype(np.longdouble)).max() <= 1.0
    a = np.exp(rng.uniform(-100, 100, 200000)) * rng.choice([-1, 1], 200000)
    b = np.exp(rng.uniform(-100, 100, 200000)) * rng.choice([-1, 1], 200000)
    assert ulps(probe(3, a, b), a.astype(np.longdouble) / b.astype(np.longdouble)).max() <= 1.0


@pytest.mark.gpu
def test_device_atan_and_div100(probe):
    rng = np.random.default_rng(3)
    n = 300000
    # atan(y / x): tangents of every incidence angle (0 .. 1e6), both signs, the range switches, x != 1 (receive angle)
    y = np.concatenate([np.tan(rng.uniform(0, np.pi / 2, n)), 10.0 ** rng.uniform(-12, 8, n) * rng.choice([-1, 1], n),
                        [0.0, 1.0, 0.41421356237309503, 2.4142135623730951, 1e300, np.inf, -np.inf, -1.0]])
    x = np.ones_like(y)
    got = probe(4, y, x)
    exact = np.arctan(y.astype(np.longdouble))
    nz = y != 0
    assert ulps(got[nz], exact[nz]).max() <= 2.5      # what the solver needs is ~1e-13 deg, i.e. ~1e3 ulp
    assert got[y == 0][0] == 0.0 and got[-3] == np.pi / 2 and got[-2] == -np.pi / 2
    yy, xx = rng.uniform(0, 1.8, n), rng.uniform(1e-3, 1.8, n)
    got = probe(4, yy, xx)
    assert ulps(got, np.arctan2(yy.astype(np.longdouble), xx.astype(np.longdouble))).max() <= 2.5
    assert probe(4, np.array([0.7]), np.array([0.0]))[0] == np.pi / 2          # grazing: asin(1)
    assert np.isnan(probe(4, np.array([np.nan, 0.5]), np.array([1.0, np.nan]))).all()
    # x / 100 must be THE IEEE quotient (the reference's cm -> m conversion feeds layer tests and exp())
    v = np.concatenate([rng.uniform(1, 1.5e7, n), 10.0 ** rng.uniform(-3, 9, n) * rng.choice([-1, 1], n), [300000.0, 20000.0, 0.0]])
    assert np.array_equal(probe(5, v), v / 100)


@pytest.mark.gpu
def test_device_glibc_exp_log_pow_equal_host_libm(probe, tmp_path):
    """The device build of airice_glibc_math.cuh against the libm of the host this test runs on: bit for bit (the in-ice
    solver's branch flags depend on it).  Same argument sets as the CPU-only test of the host build."""
    from test_glibc_math import build_libm_ref, glibc_math_cases, has_fma, same_bits
    if not has_fma():
        pytest.skip("host libm runs its non-FMA build")
    lib, run = build_libm_ref(tmp_path)
    rng = np.random.default_rng(6)
    for op, a, b in glibc_math_cases(rng, 1_000_000):
        want, got = run(lib.libm_v, op, a, b), probe(6 + op, a, b)
        ok = same_bits(want, got)
        # outside the restated fast paths the device falls back to CUDA's own function (|x| >= 512, subnormal results)
        fast = np.abs(a) < 512 if op == 0 else (np.abs(b * np.log(np.maximum(a, 1e-300))) < 512 if op == 2 else np.ones_like(ok))
        assert ok[fast].all(), (op, a[fast & ~ok][:3], want[fast & ~ok][:3], got[fast & ~ok][:3])
