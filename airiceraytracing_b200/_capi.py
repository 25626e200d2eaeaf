"""ctypes binding of libairice_b200.so (the C ABI declared in include/airice_b200.h).

There is no CPU implementation behind this module: if the CUDA library has not been built, or no GPU is
visible, the calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AIRICE_LIB: development override (tools/*_probe.py time alternative builds of the library)
LIB_PATH = os.environ.get("AIRICE_LIB") or os.path.join(_HERE, "lib", "libairice_b200.so")

TABLE_COLS64 = 17
TABLE_COLS32 = 11
SOLVE_COLS = 13
SOLVE_COLS_CM_RAD = 9
LOOKUP_COLS = 9
INICE_COLS = 29
INICE_RAYS_COLS = 10
UNITS_M_DEG = 0
UNITS_CM_RAD = 1
VARIANT_MULTIRAY = 0
VARIANT_PYWRAP = 1
VARIANT_CLI = 2

# every symbol include/airice_b200.h declares
EXPORTS = [
    "airice_create", "airice_destroy", "airice_last_error", "airice_device_count", "airice_get_medium",
    "airice_set_ice_model", "airice_table_dims", "airice_table_build_device", "airice_forward_device",
    "airice_forward_host", "airice_table_create", "airice_table_create_multi", "airice_table_wrap", "airice_table_destroy", "airice_table_info",
    "airice_table_copy_column", "airice_table_column_ptr", "airice_table_copy_row_ranges", "airice_solve_device",
    "airice_solve_multi_device", "airice_solve_host", "airice_solve_host_columns", "airice_lookup_device", "airice_lookup_host", "airice_lookup_host_columns", "airice_inice_solve_device",
    "airice_inice_solve_host", "airice_inice_two_rays_device", "airice_inice_two_rays_host", "airice_ray_path_device", "airice_ray_path_host", "airice_fp64_peak_tflops", "airice_sync", "airice_trim",
    "airice_inice_two_rays_att_device", "airice_inice_two_rays_att_host", "airice_inice_attenuation_device", "airice_inice_attenuation_host",
    "airice_inice_quadrature_stats", "airice_inice_ladder_stats", "airice_inice_focusing_device", "airice_inice_focusing_host", "airice_inice_table_create",
    "airice_inice_table_destroy", "airice_inice_table_info", "airice_inice_table_copy_column", "airice_inice_table_copy_positions",
    "airice_inice_table_interp_device", "airice_inice_table_interp_host", "airice_table_save", "airice_table_load", "airice_oldtable_create", "airice_oldtable_wrap_host", "airice_oldtable_destroy", "airice_oldtable_info",
    "airice_oldtable_copy_column", "airice_oldtable_copy_positions", "airice_oldtable_interp_device", "airice_oldtable_interp_host",
    "airice_host_register", "airice_host_unregister", "airice_host_alloc", "airice_host_free",
    "airice_peer_alloc", "airice_peer_free", "airice_peer_open", "airice_peer_close", "airice_peer_copy",
]

_lib = None


def load():
    """Load libairice_b200.so, failing loudly when it has not been built (see __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("%s is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                      "g.build()'); this package has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, d, i, i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
    pp = C.POINTER(C.c_void_p)
    lib.airice_last_error.restype = C.c_char_p
    lib.airice_create.argtypes = [C.c_char_p, i, i, pp]
    lib.airice_destroy.argtypes = [vp]
    lib.airice_destroy.restype = None
    lib.airice_get_medium.argtypes = [vp, C.POINTER(d)]
    lib.airice_set_ice_model.argtypes = [vp, d, d, d]
    lib.airice_table_dims.argtypes = [vp, d, d, d, d, d, d, d, C.POINTER(i64), C.POINTER(i64)]
    lib.airice_table_build_device.argtypes = [vp, d, d, d, d, d, d, d, i64, i64, pp, pp, vp]
    lib.airice_forward_device.argtypes = [vp, i64, vp, vp, d, d, pp, vp]
    lib.airice_table_create.argtypes = [vp, d, d, d, d, d, d, d, pp]
    lib.airice_table_create_multi.argtypes = [vp, i, C.POINTER(d), d, d, d, d, d, d, pp]
    lib.airice_table_wrap.argtypes = [vp, pp, i64, i64, d, d, pp]
    lib.airice_table_destroy.argtypes = [vp]
    lib.airice_table_destroy.restype = None
    lib.airice_table_info.argtypes = [vp, C.POINTER(i64)]
    lib.airice_table_copy_column.argtypes = [vp, i, vp]
    lib.airice_table_column_ptr.argtypes = [vp, i, pp]
    lib.airice_table_copy_row_ranges.argtypes = [vp, vp, vp]
    lib.airice_solve_device.argtypes = [vp, i64, vp, vp, vp, d, d, i, pp, vp, vp, vp]
    lib.airice_solve_multi_device.argtypes = [vp, i64, i, vp, vp, C.POINTER(d), d, i, pp, vp, vp]
    lib.airice_solve_host.argtypes = [vp, i64, vp, vp, vp, d, d, i, vp, vp]
    lib.airice_solve_host_columns.argtypes = [vp, i64, vp, vp, vp, d, d, i, pp, vp]
    lib.airice_forward_host.argtypes = [vp, i64, vp, vp, d, d, vp]
    lib.airice_lookup_device.argtypes = [vp, vp, i64, vp, vp, pp, vp, vp]
    lib.airice_lookup_host.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.airice_lookup_host_columns.argtypes = [vp, vp, i64, vp, vp, pp, vp]
    lib.airice_inice_solve_device.argtypes = [vp, i64, vp, vp, vp, pp, vp, vp]
    lib.airice_inice_solve_host.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.airice_inice_two_rays_device.argtypes = [vp, i64, vp, vp, vp, pp, pp, pp, vp]
    lib.airice_inice_two_rays_host.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.airice_ray_path_device.argtypes = [vp, i64, vp, vp, d, d, i64, vp, vp, vp, vp]
    lib.airice_ray_path_host.argtypes = [vp, i64, vp, vp, d, d, i64, vp, vp, vp]
    lib.airice_fp64_peak_tflops.argtypes = [vp, C.POINTER(d)]
    lib.airice_sync.argtypes = [vp]
    lib.airice_trim.argtypes = [vp]
    lib.airice_inice_two_rays_att_device.argtypes = [vp, i64, vp, vp, vp, d, d, pp, pp, pp, pp, vp]
    lib.airice_inice_two_rays_att_host.argtypes = [vp, i64, vp, vp, vp, d, d, vp, vp, vp]
    lib.airice_inice_attenuation_device.argtypes = [vp, i64, i, d, d, vp, vp, vp, vp, vp, vp]
    lib.airice_inice_attenuation_host.argtypes = [vp, i64, i, d, d, vp, vp, vp, vp, vp]
    lib.airice_inice_ladder_stats.argtypes = [vp, C.POINTER(i64)]
    lib.airice_inice_quadrature_stats.argtypes = [vp, C.POINTER(i64)]
    lib.airice_inice_focusing_device.argtypes = [vp, i64, vp, vp, vp, pp, vp]
    lib.airice_inice_focusing_host.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.airice_inice_table_create.argtypes = [vp, d, d, d, d, d, d, d, pp]
    lib.airice_inice_table_destroy.argtypes = [vp]
    lib.airice_inice_table_destroy.restype = None
    lib.airice_inice_table_info.argtypes = [vp, C.POINTER(i64)]
    lib.airice_inice_table_copy_column.argtypes = [vp, i, vp]
    lib.airice_inice_table_copy_positions.argtypes = [vp, vp, vp]
    lib.airice_inice_table_interp_device.argtypes = [vp, vp, i64, vp, vp, i, vp, vp]
    lib.airice_inice_table_interp_host.argtypes = [vp, vp, i64, vp, vp, i, vp]
    lib.airice_table_save.argtypes = [vp, C.c_char_p]
    lib.airice_table_load.argtypes = [vp, C.c_char_p, pp]
    lib.airice_oldtable_create.argtypes = [vp, d, d, d, d, d, d, pp]
    lib.airice_oldtable_wrap_host.argtypes = [vp, d, d, d, d, d, vp, pp]
    lib.airice_oldtable_destroy.argtypes = [vp]
    lib.airice_oldtable_destroy.restype = None
    lib.airice_oldtable_info.argtypes = [vp, C.POINTER(i64)]
    lib.airice_oldtable_copy_column.argtypes = [vp, i, vp]
    lib.airice_oldtable_copy_positions.argtypes = [vp, vp, vp]
    lib.airice_oldtable_interp_device.argtypes = [vp, vp, i64, vp, vp, i, vp, vp]
    lib.airice_oldtable_interp_host.argtypes = [vp, vp, i64, vp, vp, i, vp]
    lib.airice_host_register.argtypes = [vp, C.c_size_t]
    lib.airice_host_unregister.argtypes = [vp]
    lib.airice_host_alloc.argtypes = [C.c_size_t, pp]
    lib.airice_host_free.argtypes = [vp]
    lib.airice_peer_alloc.argtypes = [vp, C.c_size_t, pp, C.c_char_p]
    lib.airice_peer_free.argtypes = [vp, vp]
    lib.airice_peer_open.argtypes = [vp, C.c_char_p, pp]
    lib.airice_peer_close.argtypes = [vp, vp]
    lib.airice_peer_copy.argtypes = [vp, vp, vp, C.c_size_t, vp]
    _lib = lib
    return lib


class AirIceError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().airice_last_error()
        raise AirIceError("libairice_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def ptr_array(ptrs):
    """A C array of void* from a list of ints/None."""
    arr = (C.c_void_p * len(ptrs))()
    for k, p in enumerate(ptrs):
        arr[k] = p if p else None
    return arr
