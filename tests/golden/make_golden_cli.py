"""Generates tests/golden/cli_solve.npz by RUNNING THE UNMODIFIED REFERENCE command-line solver Air2IceRayTracing.C
(oracle/_ref/libcli_ref.so: its main() called as a function, stdout captured at 17 digits).

    python tests/golden/make_golden_cli.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcli_ref.so"))
    lib.cliref_air2ice.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_double)]
    rng = np.random.default_rng(2026)
    n = 400
    h = rng.uniform(3010, 60000, n)
    ang = np.concatenate([rng.uniform(106.5, 179.5, n - 120), rng.uniform(90.3, 106.0, 120)])   # a third in the clamped zone
    depth = rng.choice([200.0, 100.0, 10.0, 57.5], n)
    ice = rng.choice([3000.0, 2800.0], n)
    d = (h - ice + depth) * np.tan((180 - ang) * 3.1415927 / 180)
    h[:2], d[:2], ice[:2], depth[:2] = [5000.0, 20000.0], [1000.0, 3018.9072284385093], [3000.0, 3000.0], [200.0, 200.0]   # README example
    out = np.zeros((n, 9))
    found = np.zeros(n, dtype=np.int32)
    cwd = os.getcwd()
    os.chdir(HERE)                      # ./Atmosphere.dat
    try:
        fd = os.dup(1)                  # the CLI also printf()s nothing, but keep our stdout clean anyway
        for i in range(n):
            found[i] = lib.cliref_air2ice(h[i], d[i], ice[i], depth[i], out[i].ctypes.data_as(C.POINTER(C.c_double)))
        os.close(fd)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "cli_solve.npz"), h=h, d=d, ice=ice, depth=depth, out=out, found=found)
    print("written cli_solve.npz:", n, "calls,", int((found == 9).sum()), "complete; first:", out[0])


if __name__ == "__main__":
    main()
