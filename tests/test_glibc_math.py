"""airice_glibc_math.cuh restates glibc's exp / log / pow (FMA build) so that the in-ice kernels stop their iterations on
the same iterate as the reference's x86 build.  Here: the HOST build of that header against the libm of this machine, bit
for bit, over the argument ranges the solver produces and well beyond.  (The device build is checked against the same
libm in tests/test_gpu_math.py.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


def build_libm_ref(tmpdir):
    so = os.path.join(str(tmpdir), "liblibmref.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I" + os.path.join(ROOT, "airiceraytracing_b200", "csrc"), os.path.join(ROOT, "tests", "compat", "libm_ref.cc"),
                           "-o", so, "-lm"])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    for f in (lib.libm_v, lib.glibc_host_v):
        f.argtypes = [C.c_int, C.c_long, dp, dp, dp]
    lib.libm_sin64.restype = C.c_double
    lib.model_sin64.restype = C.c_double

    def run(fn, op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(a if b is None else b, dtype=np.float64)
        out = np.empty_like(a)
        fn(op, a.size, a.ctypes.data_as(dp), b.ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out
    return lib, run


def glibc_math_cases(rng, n):
    """argument sets: (op, a, b).  exp: -C z for depths to 5000 m and the whole fast-path range; log: the T and n+R terms
    (0.1 .. 10), the near-1 branch and its edges, wide; pow: the cube root of gsl_deriv_central's round/(2 trunc), wide."""
    ex = np.concatenate([-0.0132 * rng.uniform(0, 5000, n), rng.uniform(-511, 511, n), rng.uniform(-1, 1, n) * 1e-3,
                         rng.uniform(-1, 1, n // 10) * 2e-16, [0.0, -0.0, 1.0, -1.0, 511.9, -511.9, 5e-324, 1e-17, -1e-17]])
    lg = np.concatenate([rng.uniform(0.05, 10, n), 1 + rng.uniform(-0.07, 0.07, n), np.exp(rng.uniform(-700, 700, n)),
                         1 + rng.uniform(-1, 1, n // 4) * 10.0 ** rng.uniform(-16, -2, n // 4), rng.uniform(0, 1, n // 10) * 1e-310,
                         [1.0, 0.9375, np.nextafter(0.9375, 0), 1.0647, 1.064697265625, np.nextafter(1.064697265625, 0), 0.0, -0.0,
                          -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308]])
    px = np.concatenate([rng.uniform(0, 0.5, n), 10.0 ** rng.uniform(-30, 2, n), np.exp(rng.uniform(-50, 50, n)),
                         1 + rng.uniform(-0.01, 0.01, n // 4)])
    py = np.concatenate([np.full(2 * n, 1.0 / 3.0), rng.uniform(-2, 2, n), rng.uniform(-100, 100, n // 4)])
    return [(0, ex, None), (1, lg, None), (2, px, py)]


def same_bits(a, b):
    return (a.view(np.int64) == b.view(np.int64)) | (np.isnan(a) & np.isnan(b))


@pytest.mark.skipif(not has_fma(), reason="glibc dispatches to its non-FMA build on this CPU; the header restates the FMA build")
def test_host_build_equals_this_machines_libm(tmp_path):
    lib, run = build_libm_ref(tmp_path)
    rng = np.random.default_rng(5)
    for op, a, b in glibc_math_cases(rng, 2_000_000):
        want, got = run(lib.libm_v, op, a, b), run(lib.glibc_host_v, op, a, b)
        ok = same_bits(want, got)
        assert ok.all(), (op, a[~ok][:3], None if b is None else b[~ok][:3], want[~ok][:3], got[~ok][:3])


def test_sin64_constant(tmp_path):
    lib, _ = build_libm_ref(tmp_path)
    assert lib.libm_sin64() == lib.model_sin64() == float.fromhex("0x1.cc2ebbb5639ecp-1")
