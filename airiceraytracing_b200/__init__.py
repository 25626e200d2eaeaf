"""airiceraytracing_b200 -- B200-native batched air->ice radio ray solver.

Hot path of uzairlatif90/AirIceRayTracing (forward table build, per-pair launch-angle solve, table lookup) as
hand-written sm_100a CUDA kernels behind a C ABI (include/airice_b200.h).  This Python package is the thin host
side used by the tests and the benchmark: device memory and streams come from torch, every number comes from
libairice_b200.so.  There is no CPU fallback.
"""
from ._capi import (LIB_PATH, UNITS_CM_RAD, UNITS_M_DEG, VARIANT_CLI, VARIANT_MULTIRAY, VARIANT_PYWRAP, AirIceError)
from .solver import AirIceSolver, Table, TABLE_COLUMNS64, TABLE_COLUMNS32, SOLVE_COLUMNS_M_DEG, SOLVE_COLUMNS_CM_RAD

__all__ = ["AirIceSolver", "Table", "AirIceError", "LIB_PATH", "UNITS_CM_RAD", "UNITS_M_DEG", "VARIANT_MULTIRAY",
           "VARIANT_PYWRAP", "VARIANT_CLI", "TABLE_COLUMNS64", "TABLE_COLUMNS32", "SOLVE_COLUMNS_M_DEG", "SOLVE_COLUMNS_CM_RAD"]
