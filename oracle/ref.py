"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the parity oracles under oracle/_ref/.

* ``Oracle``    -- oracle/_ref/liboracle.so, our plain-C restatement (oracle/airice_oracle.c).
* ``Reference`` -- oracle/_ref/libmultiray_ref.so, the UNMODIFIED reference translation unit
  MultiRayAirIceRefraction.cc behind a thin extern "C" shim (oracle/ref_driver/multiray_ref.cc).
* ``PyWrapReference`` -- oracle/_ref/pywrap/libAirIceRayTracing.so, the unmodified python-wrapper library.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import contextlib
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
ATMOSPHERE = os.path.join(os.path.dirname(HERE), "tests", "golden", "Atmosphere.dat")

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_ubyte_p = C.POINTER(C.c_ubyte)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def build():
    """(Re)build the oracles with oracle/Makefile (reference targets only if /root/reference exists)."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", HERE])


class OracleAtm(C.Structure):
    _fields_ = [("variant", C.c_int), ("max_layers", C.c_int), ("npoints", C.c_int),
                ("atmlay_cm", C.c_double * 5), ("abc", (C.c_double * 3) * 5),
                ("B_air", C.c_double * 5), ("C_air", C.c_double * 5), ("A_air", C.c_double),
                ("A_ice", C.c_double), ("B_ice", C.c_double), ("C_ice", C.c_double),
                ("pi", C.c_double), ("c", C.c_double), ("n0", C.c_double)]


class OracleTable(C.Structure):
    _fields_ = [("n_h", C.c_int), ("n_th", C.c_int), ("cells", C.c_long),
                ("angle_step", C.c_double), ("angle_start", C.c_double), ("angle_stop", C.c_double),
                ("height_step", C.c_double), ("loop_start_h", C.c_double), ("loop_stop_h", C.c_double),
                ("depth_m", C.c_double), ("ice_m", C.c_double), ("col", c_float_p * 11)]


class Oracle:
    """Plain-C restatement of the reference algorithm (variant 0 = MultiRay, 1 = python wrapper)."""

    def __init__(self, atmosphere=ATMOSPHERE, variant=0):
        self.lib = C.CDLL(os.path.join(REFDIR, "liboracle.so"))
        L = self.lib
        L.oracle_nz_air.restype = C.c_double
        L.oracle_nz_ice.restype = C.c_double
        L.oracle_rootfn.restype = C.c_double
        L.oracle_nz_air.argtypes = [C.c_void_p, C.c_double]
        L.oracle_nz_ice.argtypes = [C.c_void_p, C.c_double]
        L.oracle_layer_of.argtypes = [C.c_void_p, C.c_double]
        L.oracle_rootfn.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.oracle_forward.argtypes = [C.c_void_p] + [C.c_double] * 4 + [C.c_int, c_double_p]
        L.oracle_air2ice.argtypes = [C.c_void_p] + [C.c_double] * 5 + [c_double_p]
        L.oracle_solve_cm.argtypes = [C.c_void_p] + [C.c_double] * 4 + [c_double_p]
        L.oracle_solve_cm_batch.argtypes = [C.c_void_p, C.c_long, c_double_p, c_double_p, C.c_double, C.c_double,
                                            c_double_p, c_ubyte_p]
        L.oracle_pywrap_solution.argtypes = [C.c_void_p] + [C.c_double] * 4 + [c_double_p]
        L.oracle_py_trace.argtypes = [C.c_void_p] + [C.c_double] * 4 + [c_double_p]
        L.oracle_table_build.restype = C.POINTER(OracleTable)
        L.oracle_table_build.argtypes = [C.c_void_p] + [C.c_double] * 6
        L.oracle_table_free.argtypes = [C.POINTER(OracleTable)]
        L.oracle_find_rows.argtypes = [C.POINTER(OracleTable), C.c_double, c_int_p, c_double_p]
        L.oracle_find_thd.argtypes = [C.POINTER(OracleTable), C.c_double, C.c_int, C.c_int, c_int_p, c_double_p]
        L.oracle_lookup_cm.argtypes = [C.c_void_p, C.POINTER(OracleTable)] + [C.c_double] * 4 + [c_double_p]
        L.oracle_lookup_cm_batch.argtypes = [C.c_void_p, C.POINTER(OracleTable), C.c_long, c_double_p, c_double_p,
                                             C.c_double, C.c_double, c_double_p, c_ubyte_p]
        self.atm = OracleAtm()
        rc = L.oracle_atm_load(atmosphere.encode(), variant, C.byref(self.atm))
        if rc != 0:
            raise OSError("oracle_atm_load(%s) failed: %d" % (atmosphere, rc))
        self._a = C.byref(self.atm)

    def constants(self):
        a = self.atm
        return dict(max_layers=a.max_layers, atmlay_cm=list(a.atmlay_cm), B_air=list(a.B_air), C_air=list(a.C_air),
                    A_ice=a.A_ice, B_ice=a.B_ice, C_ice=a.C_ice, pi=a.pi, n0=a.n0, npoints=a.npoints)

    def nz_air(self, z):
        return self.lib.oracle_nz_air(self._a, z)

    def nz_ice(self, z):
        return self.lib.oracle_nz_ice(self._a, z)

    def rootfn(self, theta, h, ice, depth_pos, d):
        return self.lib.oracle_rootfn(self._a, theta, h, ice, depth_pos, d)

    def forward(self, theta, h, ice, depth, inice=True):
        out = np.zeros(18)
        self.lib.oracle_forward(self._a, theta, h, ice, depth, int(inice), _dp(out))
        return out

    def forward_batch(self, theta, h, ice, depth, inice=True):
        theta = np.asarray(theta, dtype=np.float64)
        h = np.asarray(h, dtype=np.float64)
        out = np.zeros((theta.size, 18))
        for i in range(theta.size):
            self.lib.oracle_forward(self._a, theta[i], h[i], ice, depth, int(inice), _dp(out[i]))
        return out

    def air2ice(self, h, d, ice, depth, thR):
        out = np.zeros(17)
        nev = self.lib.oracle_air2ice(self._a, h, d, ice, depth, thR, _dp(out))
        return out, nev

    def solve_cm(self, h_cm, d_cm, depth_cm, ice_cm):
        out = np.zeros(9)
        ok = self.lib.oracle_solve_cm(self._a, h_cm, d_cm, depth_cm, ice_cm, _dp(out))
        return bool(ok), out

    def solve_cm_batch(self, h_cm, d_cm, depth_cm, ice_cm):
        h_cm = np.ascontiguousarray(h_cm, dtype=np.float64)
        d_cm = np.ascontiguousarray(d_cm, dtype=np.float64)
        out = np.zeros((h_cm.size, 9))
        ok = np.zeros(h_cm.size, dtype=np.uint8)
        self.lib.oracle_solve_cm_batch(self._a, h_cm.size, _dp(h_cm), _dp(d_cm), depth_cm, ice_cm, _dp(out),
                                       ok.ctypes.data_as(c_ubyte_p))
        return ok.astype(bool), out

    def pywrap_solution(self, h, d, depth, ice):
        out = np.zeros(8)
        ok = self.lib.oracle_pywrap_solution(self._a, h, d, depth, ice, _dp(out))
        return bool(ok), out

    def py_trace(self, depth, ice, h, d):
        out = np.zeros(10)
        self.lib.oracle_py_trace(self._a, depth, ice, h, d, _dp(out))
        return out

    def table_build(self, depth_cm, ice_cm, angle_step=0.1, angle_start=90.1, angle_stop=180.0, height_step=10.0):
        return OracleTableHandle(self, self.lib.oracle_table_build(self._a, depth_cm, ice_cm, angle_step, angle_start,
                                                                   angle_stop, height_step))


class OracleTableHandle:
    def __init__(self, oracle, ptr):
        self.o, self.ptr = oracle, ptr
        t = ptr.contents
        self.n_h, self.n_th, self.cells = t.n_h, t.n_th, t.cells
        self.loop_stop_h, self.height_step = t.loop_stop_h, t.height_step

    def columns(self):
        t = self.ptr.contents
        return np.stack([np.ctypeslib.as_array(t.col[k], shape=(self.cells,)).copy() for k in range(11)])

    def set_columns(self, cols):
        t = self.ptr.contents
        for k in range(11):
            np.ctypeslib.as_array(t.col[k], shape=(self.cells,))[:] = cols[k]

    def find_rows(self, h):
        idx = (C.c_int * 4)()
        cv = (C.c_double * 2)()
        self.o.lib.oracle_find_rows(self.ptr, h, idx, cv)
        return list(idx), list(cv)

    def find_thd(self, d, start, end):
        idx = (C.c_int * 2)()
        cv = C.c_double()
        self.o.lib.oracle_find_thd(self.ptr, d, start, end, idx, C.byref(cv))
        return list(idx), cv.value

    def lookup_cm(self, h_cm, d_cm, depth_cm, ice_cm):
        out = np.zeros(9)
        ok = self.o.lib.oracle_lookup_cm(self.o._a, self.ptr, h_cm, d_cm, depth_cm, ice_cm, _dp(out))
        return bool(ok), out

    def lookup_cm_batch(self, h_cm, d_cm, depth_cm, ice_cm):
        h_cm = np.ascontiguousarray(h_cm, dtype=np.float64)
        d_cm = np.ascontiguousarray(d_cm, dtype=np.float64)
        out = np.zeros((h_cm.size, 9))
        ok = np.zeros(h_cm.size, dtype=np.uint8)
        self.o.lib.oracle_lookup_cm_batch(self.o._a, self.ptr, h_cm.size, _dp(h_cm), _dp(d_cm), depth_cm, ice_cm,
                                          _dp(out), ok.ctypes.data_as(c_ubyte_p))
        return ok.astype(bool), out

    def free(self):
        if self.ptr:
            self.o.lib.oracle_table_free(self.ptr)
            self.ptr = None


def reference_available(name="libmultiray_ref.so"):
    return os.path.exists(os.path.join(REFDIR, name))


class Reference:
    """The unmodified reference MultiRayAirIceRefraction.cc (process-global state, not re-entrant).

    The reference opens "Atmosphere.dat" from the current directory (M.cc:27,80), so every call that
    (re)reads it runs inside the directory holding the fixture."""

    def __init__(self, atmosphere=ATMOSPHERE, opt="O2"):
        name = "libmultiray_ref.so" if opt == "O2" else "libmultiray_ref_O0.so"
        self.lib = C.CDLL(os.path.join(REFDIR, name))
        self.dir = os.path.dirname(os.path.abspath(atmosphere))
        assert os.path.basename(atmosphere) == "Atmosphere.dat"
        L = self.lib
        L.ref_nz_air.restype = C.c_double
        L.ref_nz_ice.restype = C.c_double
        L.ref_rootfn.restype = C.c_double
        L.ref_nz_air.argtypes = [C.c_double]
        L.ref_nz_ice.argtypes = [C.c_double]
        L.ref_rootfn.argtypes = [C.c_double] * 5
        L.ref_forward.argtypes = [C.c_double] * 4 + [C.c_int, c_double_p]
        L.ref_forward_batch.argtypes = [C.c_long, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int, c_double_p]
        L.ref_air2ice.argtypes = [C.c_double] * 5 + [c_double_p]
        L.ref_solve_cm.argtypes = [C.c_double] * 4 + [c_double_p]
        L.ref_solve_cm_batch.argtypes = [C.c_long, c_double_p, c_double_p, C.c_double, C.c_double, c_double_p, c_ubyte_p]
        L.ref_set_grid.argtypes = [C.c_double] * 4
        L.ref_make_table.argtypes = [C.c_double, C.c_double]
        L.ref_table_info.argtypes = [C.c_int, c_long_p]
        L.ref_table_col.argtypes = [C.c_int, C.c_int, c_float_p]
        L.ref_table_set_col.argtypes = [C.c_int, C.c_int, c_float_p]
        L.ref_lookup_cm.argtypes = [C.c_double] * 4 + [C.c_int, c_double_p]
        L.ref_lookup_cm_batch.argtypes = [C.c_long, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int,
                                          c_double_p, c_ubyte_p]
        L.ref_find_rows.argtypes = [C.c_double, C.c_int, c_int_p, c_double_p]
        L.ref_find_thd.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, c_int_p, c_double_p]
        L.ref_set_ice.argtypes = [C.c_double] * 3
        L.ref_quiet(1)
        with _cwd(self.dir):
            L.ref_make_atmosphere()

    def constants(self):
        c = np.zeros(21)
        self.lib.ref_constants(_dp(c))
        return dict(max_layers=int(c[0]), atmlay_cm=list(c[1:6]), B_air=list(c[6:11]), C_air=list(c[11:16]),
                    A_ice=c[16], B_ice=c[17], C_ice=c[18], pi=c[19], n0=c[20])

    def nz_air(self, z):
        return self.lib.ref_nz_air(z)

    def nz_ice(self, z):
        return self.lib.ref_nz_ice(z)

    def rootfn(self, theta, h, ice, depth_pos, d):
        return self.lib.ref_rootfn(theta, h, ice, depth_pos, d)

    def forward(self, theta, h, ice, depth, inice=True):
        out = np.zeros(18)
        self.lib.ref_forward(theta, h, ice, depth, int(inice), _dp(out))
        return out

    def forward_batch(self, theta, h, ice, depth, inice=True):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        h = np.ascontiguousarray(h, dtype=np.float64)
        out = np.zeros((theta.size, 18))
        self.lib.ref_forward_batch(theta.size, _dp(theta), _dp(h), ice, depth, int(inice), _dp(out))
        return out

    def air2ice(self, h, d, ice, depth, thR):
        out = np.zeros(17)
        self.lib.ref_air2ice(h, d, ice, depth, thR, _dp(out))
        return out

    def solve_cm(self, h_cm, d_cm, depth_cm, ice_cm):
        out = np.zeros(9)
        ok = self.lib.ref_solve_cm(h_cm, d_cm, depth_cm, ice_cm, _dp(out))
        return bool(ok), out

    def solve_cm_batch(self, h_cm, d_cm, depth_cm, ice_cm):
        h_cm = np.ascontiguousarray(h_cm, dtype=np.float64)
        d_cm = np.ascontiguousarray(d_cm, dtype=np.float64)
        out = np.zeros((h_cm.size, 9))
        ok = np.zeros(h_cm.size, dtype=np.uint8)
        self.lib.ref_solve_cm_batch(h_cm.size, _dp(h_cm), _dp(d_cm), depth_cm, ice_cm, _dp(out),
                                    ok.ctypes.data_as(c_ubyte_p))
        return ok.astype(bool), out

    def set_grid(self, angle_step=0.1, angle_start=90.1, angle_stop=180.0, height_step=10.0):
        self.lib.ref_set_grid(angle_step, angle_start, angle_stop, height_step)

    def clear_tables(self):
        self.lib.ref_clear_tables()

    def make_table(self, depth_cm, ice_cm):
        with _cwd(self.dir):
            return self.lib.ref_make_table(depth_cm, ice_cm)

    def table_info(self, ant=0):
        info = (C.c_long * 4)()
        self.lib.ref_table_info(ant, info)
        return dict(n_h=info[0], n_th=info[1], cells=info[2], ncols=info[3])

    def table_columns(self, ant=0):
        info = self.table_info(ant)
        cols = np.zeros((info["ncols"], info["cells"]), dtype=np.float32)
        for k in range(info["ncols"]):
            self.lib.ref_table_col(ant, k, cols[k].ctypes.data_as(c_float_p))
        return cols

    def set_table_columns(self, cols, ant=0):
        cols = np.ascontiguousarray(cols, dtype=np.float32)
        for k in range(cols.shape[0]):
            self.lib.ref_table_set_col(ant, k, cols[k].ctypes.data_as(c_float_p))

    def find_rows(self, h, ant=0):
        idx = (C.c_int * 4)()
        cv = (C.c_double * 2)()
        self.lib.ref_find_rows(h, ant, idx, cv)
        return list(idx), list(cv)

    def find_thd(self, d, start, end, ant=0):
        idx = (C.c_int * 2)()
        cv = C.c_double()
        self.lib.ref_find_thd(d, start, end, ant, idx, C.byref(cv))
        return list(idx), cv.value

    def lookup_cm(self, h_cm, d_cm, depth_cm, ice_cm, ant=0):
        out = np.zeros(9)
        ok = self.lib.ref_lookup_cm(h_cm, d_cm, depth_cm, ice_cm, ant, _dp(out))
        return bool(ok), out

    def lookup_cm_batch(self, h_cm, d_cm, depth_cm, ice_cm, ant=0):
        h_cm = np.ascontiguousarray(h_cm, dtype=np.float64)
        d_cm = np.ascontiguousarray(d_cm, dtype=np.float64)
        out = np.zeros((h_cm.size, 9))
        ok = np.zeros(h_cm.size, dtype=np.uint8)
        self.lib.ref_lookup_cm_batch(h_cm.size, _dp(h_cm), _dp(d_cm), depth_cm, ice_cm, ant, _dp(out),
                                     ok.ctypes.data_as(c_ubyte_p))
        return ok.astype(bool), out


def _old_table_api(ref):
    L = ref.lib
    L.ref_old_grid.argtypes = [C.c_double] * 4
    L.ref_old_make_table.argtypes = [C.c_double, C.c_double]
    L.ref_old_info.argtypes = [c_long_p]
    L.ref_old_col.argtypes = [C.c_int, c_double_p]
    L.ref_old_interp.restype = C.c_double
    L.ref_old_interp.argtypes = [C.c_double, C.c_double, C.c_int]


def old_make_table(ref, ice_cm, depth_cm, start_th=90.05, stop_th=179.95, step_h=25.0, step_th=0.01):
    """Reference MakeTable (M.cc:1618-1696) on a chosen grid -> (info dict, [9, n] columns)."""
    _old_table_api(ref)
    ref.lib.ref_old_grid(start_th, stop_th, step_h, step_th)
    with _cwd(ref.dir):
        ref.lib.ref_old_make_table(ice_cm, depth_cm)
    info = (C.c_long * 4)()
    ref.lib.ref_old_info(info)
    n = info[3]
    cols = np.zeros((9, n))
    for k in range(9):
        ref.lib.ref_old_col(k, _dp(cols[k]))
    return dict(n_h=info[0], n_th=info[1], points=info[2]), cols


def old_interp(ref, h, th, par):
    _old_table_api(ref)
    return ref.lib.ref_old_interp(h, th, par)


class PyWrapReference:
    """The unmodified reference pythonwrapper library (Py_TraceIceToAir re-parses ./Atmosphere.dat per call)."""

    def __init__(self, atmosphere=ATMOSPHERE):
        self.lib = C.CDLL(os.path.join(REFDIR, "pywrap", "libAirIceRayTracing.so"))
        self.dir = os.path.dirname(os.path.abspath(atmosphere))
        self.path = atmosphere
        L = self.lib
        L.Py_TraceIceToAir.argtypes = [C.c_double] * 4 + [C.c_double * 10]
        L.pyref_air2ice.argtypes = [C.c_double] * 5 + [c_double_p]
        L.pyref_solution.argtypes = [C.c_double] * 4 + [c_double_p]
        L.pyref_make_atmosphere.argtypes = [C.c_char_p]
        L.pyref_quiet(1)
        L.pyref_make_atmosphere(atmosphere.encode())

    def py_trace(self, depth, ice, h, d):
        arr = (C.c_double * 10)(*([1.0] * 10))
        with _cwd(self.dir):
            self.lib.Py_TraceIceToAir(depth, ice, h, d, arr)
        return np.array(list(arr))

    def solution(self, h, d, depth, ice):
        out = np.zeros(8)
        ok = self.lib.pyref_solution(h, d, depth, ice, _dp(out))
        return bool(ok), out

    def air2ice(self, h, d, ice, depth, thR):
        out = np.zeros(15)
        self.lib.pyref_air2ice(h, d, ice, depth, thR, _dp(out))
        return out

    def set_constant_air(self, on, ice=3000.0):
        """UseConstantRefractiveIndex / A_const / A_air as TraceIceToAir.C:27-29 would set them (commented out there)."""
        self.lib.pyref_set_constant_air.argtypes = [C.c_int, C.c_double]
        self.lib.pyref_set_constant_air(int(on), float(ice))

    def rootfn3(self, h, d, ice, depth, thetas):
        self.lib.pyref_rootfn3.argtypes = [C.c_double] * 4 + [c_double_p, c_double_p]
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        out = np.zeros(3)
        self.lib.pyref_rootfn3(h, d, ice, depth, _dp(th), _dp(out))
        return out


class InIceOracle:
    """Our plain-C restatement of the in-ice solver (oracle/inice_oracle.c)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(REFDIR, "liboracle_inice.so"))
        self.lib.inice_oracle_solve_batch.argtypes = [C.c_long, c_double_p, c_double_p, c_double_p, c_double_p]
        self.lib.inice_oracle_set_model.argtypes = [C.c_double] * 3
        self.lib.inice_oracle_two_rays_batch.argtypes = [C.c_long, c_double_p, c_double_p, c_double_p, c_double_p,
                                                         C.POINTER(C.c_int)]

    def set_model(self, A, B, Cc):
        self.lib.inice_oracle_set_model(A, B, Cc)

    def two_rays(self, rx_depth, distance, tx_depth):
        return _two_rays(self.lib.inice_oracle_two_rays_batch, rx_depth, distance, tx_depth)

    def solve_batch(self, z0, x1, z1):
        z0 = np.ascontiguousarray(z0, dtype=np.float64)
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        z1 = np.ascontiguousarray(z1, dtype=np.float64)
        out = np.zeros((z0.size, 29))
        self.lib.inice_oracle_solve_batch(z0.size, _dp(z0), _dp(x1), _dp(z1), _dp(out))
        return out


def _two_rays(fn, rx, dist, tx):
    rx = np.ascontiguousarray(rx, dtype=np.float64)
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    tx = np.ascontiguousarray(tx, dtype=np.float64)
    out = np.zeros((rx.size, 10))
    ig = np.zeros((rx.size, 2), dtype=np.int32)
    fn(rx.size, _dp(rx), _dp(dist), _dp(tx), _dp(out), ig.ctypes.data_as(C.POINTER(C.c_int)))
    return out, ig


class IceRayReference:
    """The unmodified reference IceRayTracing.cc (in-ice direct / reflected / refracted solver)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(REFDIR, "libiceray_ref.so"))
        self.lib.iceref_solve_batch.argtypes = [C.c_long, c_double_p, c_double_p, c_double_p, c_double_p]
        self.lib.iceref_two_rays_batch.argtypes = [C.c_long, c_double_p, c_double_p, c_double_p, c_double_p,
                                                   C.POINTER(C.c_int)]

    def two_rays(self, rx_depth, distance, tx_depth):
        """GetRayTracingSolutions(RxDepth, Distance, TxDepth, ...) -> (out[n,10], IgnoreCh[n,2])"""
        return _two_rays(self.lib.iceref_two_rays_batch, rx_depth, distance, tx_depth)

    # ---- attenuation (through the stand-in's QAGS), focusing, in-ice table
    def two_rays_att(self, rx_depth, distance, tx_depth, A0, frequency):
        """-> (out[n,10], AttRay[n,2], IgnoreCh[n,2])"""
        rx = np.ascontiguousarray(rx_depth, dtype=np.float64)
        dist = np.ascontiguousarray(distance, dtype=np.float64)
        tx = np.ascontiguousarray(tx_depth, dtype=np.float64)
        out, att, ig = np.zeros((rx.size, 10)), np.zeros((rx.size, 2)), np.zeros((rx.size, 2), dtype=np.int32)
        f = self.lib.iceref_two_rays_att_batch
        f.argtypes = [C.c_long, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, c_double_p, c_double_p, C.POINTER(C.c_int)]
        f(rx.size, _dp(rx), _dp(dist), _dp(tx), A0, frequency, _dp(out), _dp(att), ig.ctypes.data_as(C.POINTER(C.c_int)))
        return out, att, ig

    def attenuation(self, kind, A0, frequency, z0, z1, zmax, L):
        f = self.lib.iceref_attenuation
        f.restype = C.c_double
        f.argtypes = [C.c_int] + [C.c_double] * 6
        return f(kind, A0, frequency, z0, z1, zmax, L)

    def focusing(self, zT, xR, zR):
        zT = np.ascontiguousarray(zT, dtype=np.float64)
        xR = np.ascontiguousarray(xR, dtype=np.float64)
        zR = np.ascontiguousarray(zR, dtype=np.float64)
        out = np.zeros((zT.size, 2))
        f = self.lib.iceref_focusing_batch
        f.argtypes = [C.c_long, c_double_p, c_double_p, c_double_p, c_double_p]
        f(zT.size, _dp(zT), _dp(xR), _dp(zR), _dp(out))
        return out

    def make_table(self, hit_distance, shower_depth, zR, step_x=0.1, step_z=0.1, width_x=40.0, width_z=20.0, ant=0, n_ant=1):
        """MakeTable on the given grid -> (columns [13, points], pos_x float32, pos_z float32)"""
        L = self.lib
        L.iceref_set_grid.argtypes = [C.c_double] * 4
        L.iceref_make_table.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
        L.iceref_set_grid(step_x, step_z, width_x, width_z)
        L.iceref_make_table(n_ant, hit_distance, shower_depth, zR, ant)
        info = (C.c_long * 3)()
        L.iceref_table_info(ant, info)
        cols = np.zeros((13, info[2]))
        L.iceref_table_column.argtypes = [C.c_int, C.c_int, c_double_p]
        for k in range(13):
            L.iceref_table_column(ant, k, _dp(cols[k]))
        px, pz = np.zeros(info[0], dtype=np.float32), np.zeros(info[1], dtype=np.float32)
        L.iceref_table_positions.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.iceref_table_positions(ant, px.ctypes.data, pz.ctypes.data)
        return cols, px, pz

    def interp(self, x, z, par, ant=0):
        f = self.lib.iceref_interp
        f.restype = C.c_double
        f.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        return np.array([f(float(a), float(b), par, ant) for a, b in zip(x, z)])

    def solve_batch(self, z0, x1, z1):
        z0 = np.ascontiguousarray(z0, dtype=np.float64)
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        z1 = np.ascontiguousarray(z1, dtype=np.float64)
        out = np.zeros((z0.size, 29))
        self.lib.iceref_solve_batch(z0.size, _dp(z0), _dp(x1), _dp(z1), _dp(out))
        return out
