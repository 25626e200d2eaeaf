"""Host-side handle on the C ABI: one AirIceSolver per GPU (mirrors the reference's per-process state:
MakeAtmosphere() once, then tables / solves / lookups).  All arrays are torch CUDA tensors (device API) or
numpy / pinned torch CPU tensors (host API); nothing is computed in Python."""
import ctypes as C
import weakref

import numpy as np
import torch

from . import _capi
from ._capi import check, ptr_array

# dummy[1..17] of GetRayTracingSolutions (MultiRayAirIceRefraction.cc:1999-2016)
TABLE_COLUMNS64 = ["h", "x", "x_air", "x_ice", "opt", "opt_air", "opt_ice", "t_ns", "t_air_ns", "t_ice_ns", "launch",
                   "incident", "received", "trans_s", "trans_p", "geo_air", "geo_ice"]
# README.md:8 "13 columns": entry number (implicit cell index) + these twelve
README_COLUMNS = ["h", "x", "x_air", "x_ice", "t_ns", "t_air_ns", "t_ice_ns", "launch", "incident", "received",
                  "trans_s", "trans_p"]
# AllTableAllAntData order (MultiRayAirIceRefraction.cc:2101-2111)
TABLE_COLUMNS32 = ["h", "x", "opt_ice", "opt_air", "launch", "x_air", "trans_s", "trans_p", "geo_air", "geo_ice",
                   "received"]
SOLVE_COLUMNS_M_DEG = ["x", "x_air", "x_ice", "t_air", "t_ice", "launch", "received", "trans_s", "trans_p", "geo_air",
                       "geo_ice", "incident", "refracted"]
# by-reference arguments of GetHorizontalDistanceToIntersectionPoint (MultiRayAirIceRefraction.h:170)
SOLVE_COLUMNS_CM_RAD = ["opt_ice", "opt_air", "geo_ice", "geo_air", "launch", "x_air", "trans_s", "trans_p", "received"]

REFERENCE_GRID = dict(h_top=100000.0, h_step=10.0, th_start=90.1, th_step=0.1, th_stop=180.0)  # M.cc:12-18,2044
README_GRID = dict(h_top=100000.0, h_step=20.0, th_start=92.0, th_step=0.5, th_stop=180.0)    # README.md:7-8


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class Table:
    """Device-resident float table (one entry of the reference's AllTableAllAntData) plus its row trim ranges."""

    def __init__(self, solver, handle, keepalive=None):
        self.solver, self.handle, self._keep = solver, handle, keepalive
        solver._tables.add(self)          # closed with the solver: a table must not outlive its context
        info = (C.c_int64 * 4)()
        check(solver.lib.airice_table_info(handle, info))
        self.n_h, self.n_th, self.cells = int(info[0]), int(info[1]), int(info[2])

    def columns(self):
        out = np.empty((_capi.TABLE_COLS32, self.cells), dtype=np.float32)
        for k in range(_capi.TABLE_COLS32):
            check(self.solver.lib.airice_table_copy_column(self.handle, k, out[k].ctypes.data))
        return out

    def row_ranges(self):
        first = np.empty(self.n_h, dtype=np.int32)
        last = np.empty(self.n_h, dtype=np.int32)
        check(self.solver.lib.airice_table_copy_row_ranges(self.handle, first.ctypes.data, last.ctypes.data))
        return first, last

    def save(self, path):
        """Write the 11 float columns + grid description to a versioned file (airice_table_save)."""
        check(self.solver.lib.airice_table_save(self.handle, str(path).encode()))

    def close(self):
        if self.handle:
            if getattr(self.solver, "handle", None):     # the context is gone: its tables went with it
                self.solver.lib.airice_table_destroy(self.handle)
            self.handle = None
            self.solver._tables.discard(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OldTable:
    """Device-resident old solve-per-cell grid (GridZValue[0..8]) with batched GetInterpolatedValue."""

    def __init__(self, solver, handle):
        self.solver, self.handle = solver, handle
        info = (C.c_int64 * 3)()
        check(solver.lib.airice_oldtable_info(handle, info))
        self.n_h, self.n_th, self.points = int(info[0]), int(info[1]), int(info[2])

    def columns(self):
        out = np.empty((9, self.points), dtype=np.float64)
        for k in range(9):
            check(self.solver.lib.airice_oldtable_copy_column(self.handle, k, out[k].ctypes.data))
        return out

    def positions(self):
        ph, pt = np.empty(self.n_h), np.empty(self.n_th)
        check(self.solver.lib.airice_oldtable_copy_positions(self.handle, ph.ctypes.data, pt.ctypes.data))
        return ph, pt

    def interp(self, h, th, par):
        s = self.solver
        h = h.to(s.torch_device, torch.float64).contiguous()
        th = th.to(s.torch_device, torch.float64).contiguous()
        out = torch.empty_like(h)
        check(s.lib.airice_oldtable_interp_device(s.handle, self.handle, h.numel(), h.data_ptr(), th.data_ptr(), int(par),
                                                  out.data_ptr(), _stream_ptr(s.torch_device)))
        return out

    def interp_host(self, h, th, par):
        h = np.ascontiguousarray(h, dtype=np.float64)
        th = np.ascontiguousarray(th, dtype=np.float64)
        out = np.empty_like(h)
        check(self.solver.lib.airice_oldtable_interp_host(self.solver.handle, self.handle, h.size, h.ctypes.data, th.ctypes.data,
                                                          int(par), out.ctypes.data))
        return out

    def close(self):
        if self.handle:
            if getattr(self.solver, "handle", None):
                self.solver.lib.airice_oldtable_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AirIceSolver:
    def __init__(self, atmosphere="Atmosphere.dat", variant=_capi.VARIANT_MULTIRAY, device=0):
        self.lib = _capi.load()
        self._tables = weakref.WeakSet()
        if not torch.cuda.is_available():
            raise _capi.AirIceError("no CUDA device visible: airiceraytracing_b200 has no CPU path")
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        h = C.c_void_p()
        # atmosphere=None: an ice-only context (in-ice entry points only, no GDAS file needed)
        check(self.lib.airice_create(str(atmosphere).encode() if atmosphere else None, int(variant), self.device, C.byref(h)))
        self.handle = h
        self.variant = variant

    def close(self):
        if getattr(self, "handle", None):
            for t in list(self._tables):
                t.close()
            self.lib.airice_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ medium
    def medium(self):
        out = (C.c_double * 24)()
        check(self.lib.airice_get_medium(self.handle, out))
        return dict(max_layers=int(out[0]), atmlay_cm=list(out[1:6]), B_air=list(out[6:11]), C_air=list(out[11:16]),
                    A_ice=out[16], B_ice=out[17], C_ice=out[18], pi=out[19], n0=out[20], npoints=int(out[21]))

    def set_ice_model(self, A, B, C_):
        check(self.lib.airice_set_ice_model(self.handle, A, B, C_))

    def fp64_peak_tflops(self):
        v = C.c_double()
        check(self.lib.airice_fp64_peak_tflops(self.handle, C.byref(v)))
        return v.value

    def sync(self):
        check(self.lib.airice_sync(self.handle))

    def trim(self):
        """Return the context's cached device memory (buffers of closed tables, solve scratch) to the driver."""
        check(self.lib.airice_trim(self.handle))

    # ------------------------------------------------------------------ peer memory (multi-GPU reassembly, dist.PeerGather)
    def peer_alloc(self, nbytes):
        """-> (device pointer, 64-byte handle another process on this node can open)"""
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        check(self.lib.airice_peer_alloc(self.handle, int(nbytes), C.byref(p), h))
        return p.value, h.raw

    def peer_open(self, handle):
        p = C.c_void_p()
        check(self.lib.airice_peer_open(self.handle, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def peer_close(self, ptr):
        check(self.lib.airice_peer_close(self.handle, ptr))

    def peer_free(self, ptr):
        check(self.lib.airice_peer_free(self.handle, ptr))

    def peer_copy(self, dst, src, nbytes, stream=None):
        check(self.lib.airice_peer_copy(self.handle, dst, src, int(nbytes),
                                        _stream_ptr(self.torch_device) if stream is None else stream))

    def wrap_device_memory(self, ptr, shape, dtype):
        """A torch view of device memory of THIS device that the library allocated (no copy, no ownership)."""
        typestr = {torch.float64: "<f8", torch.uint8: "|u1", torch.int32: "<i4", torch.float32: "<f4"}[dtype]

        class _Mem:
            __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                        "strides": None}
        return torch.as_tensor(_Mem(), device=self.torch_device)

    def solve_into(self, h, d, depth, ice, units, out_ptrs, ok_ptr, straight=None):
        """airice_solve_device with raw output pointers (columns anywhere this GPU can store to, peer memory included)."""
        h = h.to(self.torch_device, torch.float64).contiguous()
        d = d.to(self.torch_device, torch.float64).contiguous()
        check(self.lib.airice_solve_device(self.handle, h.numel(), h.data_ptr(), d.data_ptr(),
                                           straight.data_ptr() if straight is not None else None, depth, ice, units,
                                           ptr_array(out_ptrs), ok_ptr, None, _stream_ptr(self.torch_device)))

    # ------------------------------------------------------------------ kernel 1
    def table_dims(self, depth_m, ice_m, h_top=100000.0, h_step=10.0, th_start=90.1, th_step=0.1, th_stop=180.0):
        nh, nth = C.c_int64(), C.c_int64()
        check(self.lib.airice_table_dims(self.handle, depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop,
                                         C.byref(nh), C.byref(nth)))
        return nh.value, nth.value

    def table_build(self, depth_m, ice_m, h_top=100000.0, h_step=10.0, th_start=90.1, th_step=0.1, th_stop=180.0,
                    rows=None, columns64=TABLE_COLUMNS64, want32=False, out64=None, out32=None):
        """Build rows [rows[0], rows[1]) of the forward table into SoA device tensors.

        Returns (f64 [len(columns64), cells] or None, f32 [11, cells] or None)."""
        n_h, n_th = self.table_dims(depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop)
        r0, r1 = (0, n_h) if rows is None else rows
        cells = (r1 - r0) * n_th
        p64 = None
        if columns64:
            if out64 is None:
                out64 = torch.empty((len(columns64), cells), dtype=torch.float64, device=self.torch_device)
            ptrs = [None] * _capi.TABLE_COLS64
            for k, name in enumerate(columns64):
                ptrs[TABLE_COLUMNS64.index(name)] = out64[k].data_ptr()
            p64 = ptr_array(ptrs)
        p32 = None
        if want32:
            if out32 is None:
                out32 = torch.empty((_capi.TABLE_COLS32, cells), dtype=torch.float32, device=self.torch_device)
            p32 = ptr_array([out32[k].data_ptr() for k in range(_capi.TABLE_COLS32)])
        check(self.lib.airice_table_build_device(self.handle, depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop,
                                                 r0, r1, p64, p32, _stream_ptr(self.torch_device)))
        return (out64 if columns64 else None), (out32 if want32 else None)

    def forward(self, theta, h, depth_m, ice_m):
        """Batched GetRayTracingSolutions on arbitrary (theta, h) device tensors -> f64 [17, n]."""
        theta = theta.to(self.torch_device, torch.float64).contiguous()
        h = h.to(self.torch_device, torch.float64).contiguous()
        n = theta.numel()
        out = torch.empty((_capi.TABLE_COLS64, n), dtype=torch.float64, device=self.torch_device)
        check(self.lib.airice_forward_device(self.handle, n, theta.data_ptr(), h.data_ptr(), depth_m, ice_m,
                                             ptr_array([out[k].data_ptr() for k in range(_capi.TABLE_COLS64)]),
                                             _stream_ptr(self.torch_device)))
        return out

    def table_create(self, depth_m, ice_m, h_top=100000.0, h_step=10.0, th_start=90.1, th_step=0.1, th_stop=180.0):
        h = C.c_void_p()
        check(self.lib.airice_table_create(self.handle, depth_m, ice_m, h_top, h_step, th_start, th_step, th_stop,
                                           C.byref(h)))
        return Table(self, h)

    def table_create_multi(self, depths_m, ice_m, h_top=100000.0, h_step=10.0, th_start=90.1, th_step=0.1, th_stop=180.0):
        """Tables of several in-ice antennas in one pass (the air walk is shared; airice_table_create_multi)."""
        n = len(depths_m)
        dep = (C.c_double * n)(*[float(x) for x in depths_m])
        hs = (C.c_void_p * n)()
        check(self.lib.airice_table_create_multi(self.handle, n, dep, ice_m, h_top, h_step, th_start, th_step, th_stop,
                                                 C.cast(hs, C.POINTER(C.c_void_p))))
        return [Table(self, C.c_void_p(hs[q])) for q in range(n)]

    def table_load(self, path):
        h = C.c_void_p()
        check(self.lib.airice_table_load(self.handle, str(path).encode(), C.byref(h)))
        return Table(self, h)

    def oldtable_create(self, ice_m, depth_m, start_th=90.05, stop_th=179.95, step_h=25.0, step_th=0.01):
        """MakeTable (old solve-per-cell grid) on the device; reference defaults = 3880 x 8990 solves."""
        h = C.c_void_p()
        check(self.lib.airice_oldtable_create(self.handle, ice_m, depth_m, start_th, stop_th, step_h, step_th, C.byref(h)))
        return OldTable(self, h)

    def oldtable_wrap(self, cols9, ice_m, start_th, stop_th, step_h, step_th):
        cols9 = np.ascontiguousarray(cols9, dtype=np.float64)
        h = C.c_void_p()
        check(self.lib.airice_oldtable_wrap_host(self.handle, ice_m, start_th, stop_th, step_h, step_th, cols9.ctypes.data, C.byref(h)))
        return OldTable(self, h)

    def table_wrap(self, cols32, n_h, n_th, loop_stop_h, h_step):
        """Wrap an [11, n_h*n_th] float32 device tensor (kept alive by the returned Table)."""
        cols32 = cols32.to(self.torch_device, torch.float32).contiguous()
        h = C.c_void_p()
        check(self.lib.airice_table_wrap(self.handle, ptr_array([cols32[k].data_ptr() for k in range(11)]), n_h, n_th,
                                         loop_stop_h, h_step, C.byref(h)))
        return Table(self, h, keepalive=cols32)

    # ------------------------------------------------------------------ kernel 2
    def solve(self, h, d, depth, ice, units=_capi.UNITS_CM_RAD, out=None, ok=None, nevals=False, straight=None):
        """Batched launch-angle solve on device tensors.  Returns (out [ncols, n] f64, ok [n] uint8[, nevals])."""
        h = h.to(self.torch_device, torch.float64).contiguous()
        d = d.to(self.torch_device, torch.float64).contiguous()
        if straight is not None:
            straight = straight.to(self.torch_device, torch.float64).contiguous()
        n = h.numel()
        nc = _capi.SOLVE_COLS_CM_RAD if units == _capi.UNITS_CM_RAD else _capi.SOLVE_COLS
        if out is None:
            out = torch.empty((nc, n), dtype=torch.float64, device=self.torch_device)
        if ok is None:
            ok = torch.empty(n, dtype=torch.uint8, device=self.torch_device)
        nev = torch.empty(n, dtype=torch.int32, device=self.torch_device) if nevals else None
        check(self.lib.airice_solve_device(self.handle, n, h.data_ptr(), d.data_ptr(),
                                           straight.data_ptr() if straight is not None else None, depth, ice, units,
                                           ptr_array([out[k].data_ptr() for k in range(nc)]), ok.data_ptr(),
                                           nev.data_ptr() if nevals else None, _stream_ptr(self.torch_device)))
        return (out, ok, nev) if nevals else (out, ok)

    def solve_multi(self, h, d, depths, ice, units=_capi.UNITS_CM_RAD, out=None, ok=None):
        """n_points Tx heights against len(depths) receivers; d is [n_ant, n_points]; returns (out [ncols, n_ant, n_points], ok)."""
        h = h.to(self.torch_device, torch.float64).contiguous()
        d = d.to(self.torch_device, torch.float64).contiguous()
        n_ant, n = int(d.shape[0]), int(h.numel())
        assert d.shape[1] == n and len(depths) == n_ant
        nc = _capi.SOLVE_COLS_CM_RAD if units == _capi.UNITS_CM_RAD else _capi.SOLVE_COLS
        if out is None:
            out = torch.empty((nc, n_ant, n), dtype=torch.float64, device=self.torch_device)
        if ok is None:
            ok = torch.empty((n_ant, n), dtype=torch.uint8, device=self.torch_device)
        dep = (C.c_double * n_ant)(*[float(x) for x in depths])
        check(self.lib.airice_solve_multi_device(self.handle, n, n_ant, h.data_ptr(), d.data_ptr(), dep, ice, units,
                                                 ptr_array([out[k].data_ptr() for k in range(nc)]), ok.data_ptr(),
                                                 _stream_ptr(self.torch_device)))
        return out, ok

    def solve_host(self, h, d, depth, ice, units=_capi.UNITS_CM_RAD, out=None, ok=None):
        """Same through host buffers (numpy arrays or pinned torch CPU tensors); copies happen inside the call."""
        n = int(h.shape[0])
        nc = _capi.SOLVE_COLS_CM_RAD if units == _capi.UNITS_CM_RAD else _capi.SOLVE_COLS
        if out is None:
            out = np.empty((nc, n), dtype=np.float64)
        if ok is None:
            ok = np.empty(n, dtype=np.uint8)
        check(self.lib.airice_solve_host(self.handle, n, _host_ptr(h), _host_ptr(d), None, depth, ice, units,
                                         _host_ptr(out), _host_ptr(ok)))
        return out, ok

    def solve_host_columns(self, h, d, depth, ice, units=_capi.UNITS_CM_RAD, columns=(), want_ok=True, ok=None):
        """Host-buffer solve that brings back only the listed output columns (and the flag): the host path is PCIe-bound,
        so unused columns are not stored by the kernel and do not cross the link.  Returns ({column index: array}, ok)."""
        n = int(h.shape[0])
        nc = _capi.SOLVE_COLS_CM_RAD if units == _capi.UNITS_CM_RAD else _capi.SOLVE_COLS
        # `columns`: column indices, or {index: preallocated float64 host array / pinned tensor of n elements}
        if isinstance(columns, dict):
            cols = {int(k): v for k, v in columns.items()}
        else:
            cols = {int(k): np.empty(n, dtype=np.float64) for k in columns}
        if any(k < 0 or k >= nc for k in cols):
            raise ValueError("column index out of range for these units")
        if ok is None and want_ok:
            ok = np.empty(n, dtype=np.uint8)
        check(self.lib.airice_solve_host_columns(self.handle, n, _host_ptr(h), _host_ptr(d), None, depth, ice, units,
                                                 ptr_array([_host_ptr(cols[k]) if k in cols else None for k in range(nc)]),
                                                 _host_ptr(ok) if ok is not None else None))
        return cols, ok

    # ------------------------------------------------------------------ kernel 3
    def lookup(self, table, h_cm, d_cm, out=None, ok=None):
        h_cm = h_cm.to(self.torch_device, torch.float64).contiguous()
        d_cm = d_cm.to(self.torch_device, torch.float64).contiguous()
        n = h_cm.numel()
        if out is None:
            out = torch.empty((_capi.LOOKUP_COLS, n), dtype=torch.float64, device=self.torch_device)
        if ok is None:
            ok = torch.empty(n, dtype=torch.uint8, device=self.torch_device)
        check(self.lib.airice_lookup_device(self.handle, table.handle, n, h_cm.data_ptr(), d_cm.data_ptr(),
                                            ptr_array([out[k].data_ptr() for k in range(_capi.LOOKUP_COLS)]),
                                            ok.data_ptr(), _stream_ptr(self.torch_device)))
        return out, ok

    def lookup_host(self, table, h_cm, d_cm, out=None, ok=None):
        n = int(h_cm.shape[0])
        if out is None:
            out = np.empty((_capi.LOOKUP_COLS, n), dtype=np.float64)
        if ok is None:
            ok = np.empty(n, dtype=np.uint8)
        check(self.lib.airice_lookup_host(self.handle, table.handle, n, _host_ptr(h_cm), _host_ptr(d_cm), _host_ptr(out),
                                          _host_ptr(ok)))
        return out, ok

    def lookup_host_columns(self, table, h_cm, d_cm, columns=(), want_ok=True, ok=None):
        """Host-buffer lookup that brings back only the listed columns (see solve_host_columns)."""
        n = int(h_cm.shape[0])
        if isinstance(columns, dict):
            cols = {int(k): v for k, v in columns.items()}
        else:
            cols = {int(k): np.empty(n, dtype=np.float64) for k in columns}
        if any(k < 0 or k >= _capi.LOOKUP_COLS for k in cols):
            raise ValueError("column index out of range")
        if ok is None and want_ok:
            ok = np.empty(n, dtype=np.uint8)
        check(self.lib.airice_lookup_host_columns(self.handle, table.handle, n, _host_ptr(h_cm), _host_ptr(d_cm),
                                                  ptr_array([_host_ptr(cols[k]) if k in cols else None for k in range(_capi.LOOKUP_COLS)]),
                                                  _host_ptr(ok) if ok is not None else None))
        return cols, ok


INICE_COLUMNS = ["launch_d", "launch_r", "launch_ra1", "launch_ra2", "t_d", "t_r", "t_ra1", "t_ra2", "recv_d", "recv_r",
                 "recv_ra1", "recv_ra2", "t_r_1", "t_r_2", "t_ra1_1", "t_ra1_2", "t_ra2_1", "t_ra2_2", "incidence",
                 "L_d", "L_r", "L_ra1", "L_ra2", "zmax_1", "zmax_2", "path_d", "path_r", "path_ra1", "path_ra2"]


def _inice_solve(self, z0, x1, z1):
    """In-ice D/R/Ra solver on device tensors -> (out [29, n] f64, mask [n] uint8)."""
    z0 = z0.to(self.torch_device, torch.float64).contiguous()
    x1 = x1.to(self.torch_device, torch.float64).contiguous()
    z1 = z1.to(self.torch_device, torch.float64).contiguous()
    n = z0.numel()
    out = torch.empty((_capi.INICE_COLS, n), dtype=torch.float64, device=self.torch_device)
    mask = torch.empty(n, dtype=torch.uint8, device=self.torch_device)
    check(self.lib.airice_inice_solve_device(self.handle, n, z0.data_ptr(), x1.data_ptr(), z1.data_ptr(),
                                             ptr_array([out[k].data_ptr() for k in range(_capi.INICE_COLS)]),
                                             mask.data_ptr(), _stream_ptr(self.torch_device)))
    return out, mask


def _inice_solve_host(self, z0, x1, z1, out=None, mask=None):
    """Host buffers in and out (numpy arrays or CPU tensors; pinned memory lets the chunks' copies overlap the kernels)."""
    n = int(z0.shape[0])
    if out is None:
        out = np.empty((_capi.INICE_COLS, n), dtype=np.float64)
    if mask is None:
        mask = np.empty(n, dtype=np.uint8)
    check(self.lib.airice_inice_solve_host(self.handle, n, _host_ptr(z0), _host_ptr(x1), _host_ptr(z1), _host_ptr(out),
                                           _host_ptr(mask)))
    return out, mask


INICE_RAYS_COLUMNS = ["time_0", "time_1", "path_0", "path_1", "launch_0", "launch_1", "recv_0", "recv_1", "incidence_0",
                      "incidence_1"]


def _inice_two_rays(self, rx_depth, distance, tx_depth, want_type=False):
    """IceRayTracing::GetRayTracingSolutions(RxDepth, Distance, TxDepth, ...) on device tensors (attenuation excluded)
    -> (out [10, n] f64, ignore [2, n] int32[, type [2, n] int32])."""
    rx = rx_depth.to(self.torch_device, torch.float64).contiguous()
    ds = distance.to(self.torch_device, torch.float64).contiguous()
    tx = tx_depth.to(self.torch_device, torch.float64).contiguous()
    n = rx.numel()
    out = torch.empty((_capi.INICE_RAYS_COLS, n), dtype=torch.float64, device=self.torch_device)
    ig = torch.empty((2, n), dtype=torch.int32, device=self.torch_device)
    ty = torch.empty((2, n), dtype=torch.int32, device=self.torch_device) if want_type else None
    check(self.lib.airice_inice_two_rays_device(
        self.handle, n, rx.data_ptr(), ds.data_ptr(), tx.data_ptr(),
        ptr_array([out[k].data_ptr() for k in range(_capi.INICE_RAYS_COLS)]), ptr_array([ig[0].data_ptr(), ig[1].data_ptr()]),
        ptr_array([ty[0].data_ptr(), ty[1].data_ptr()]) if want_type else None, _stream_ptr(self.torch_device)))
    return (out, ig, ty) if want_type else (out, ig)


def _inice_two_rays_att(self, rx_depth, distance, tx_depth, A0, frequency_ghz, want_type=False):
    """GetRayTracingSolutions with attenuation -> (out [10, n], att [2, n] = AttRay, ignore [2, n] int32[, type [2, n]])."""
    rx = rx_depth.to(self.torch_device, torch.float64).contiguous()
    ds = distance.to(self.torch_device, torch.float64).contiguous()
    tx = tx_depth.to(self.torch_device, torch.float64).contiguous()
    n = rx.numel()
    out = torch.empty((_capi.INICE_RAYS_COLS, n), dtype=torch.float64, device=self.torch_device)
    att = torch.empty((2, n), dtype=torch.float64, device=self.torch_device)
    ig = torch.empty((2, n), dtype=torch.int32, device=self.torch_device)
    ty = torch.empty((2, n), dtype=torch.int32, device=self.torch_device) if want_type else None
    check(self.lib.airice_inice_two_rays_att_device(
        self.handle, n, rx.data_ptr(), ds.data_ptr(), tx.data_ptr(), float(A0), float(frequency_ghz),
        ptr_array([out[k].data_ptr() for k in range(_capi.INICE_RAYS_COLS)]), ptr_array([att[0].data_ptr(), att[1].data_ptr()]),
        ptr_array([ig[0].data_ptr(), ig[1].data_ptr()]), ptr_array([ty[0].data_ptr(), ty[1].data_ptr()]) if want_type else None,
        _stream_ptr(self.torch_device)))
    return (out, att, ig, ty) if want_type else (out, att, ig)


def _inice_two_rays_att_host(self, rx_depth, distance, tx_depth, A0, frequency_ghz):
    n = int(rx_depth.shape[0])
    out = np.empty((_capi.INICE_RAYS_COLS, n), dtype=np.float64)
    att = np.empty((2, n), dtype=np.float64)
    ig = np.empty((2, n), dtype=np.int32)
    check(self.lib.airice_inice_two_rays_att_host(self.handle, n, _host_ptr(rx_depth), _host_ptr(distance), _host_ptr(tx_depth),
                                                  float(A0), float(frequency_ghz), _host_ptr(out), _host_ptr(att), _host_ptr(ig)))
    return out, att, ig


def _inice_attenuation(self, kind, A0, frequency_ghz, z0, z1, L, zmax=None):
    """GetTotalAttenuationDirect (kind 0) / Reflected (1) / Refracted (2) on device tensors -> [n]."""
    z0 = z0.to(self.torch_device, torch.float64).contiguous()
    z1 = z1.to(self.torch_device, torch.float64).contiguous()
    L = L.to(self.torch_device, torch.float64).contiguous()
    zm = zmax.to(self.torch_device, torch.float64).contiguous() if zmax is not None else None
    out = torch.empty_like(z0)
    check(self.lib.airice_inice_attenuation_device(self.handle, z0.numel(), int(kind), float(A0), float(frequency_ghz), z0.data_ptr(),
                                                   z1.data_ptr(), zm.data_ptr() if zm is not None else None, L.data_ptr(),
                                                   out.data_ptr(), _stream_ptr(self.torch_device)))
    return out


def _inice_focusing(self, zT, xR, zR):
    """GetFocusingFactor(zT, xR, zR) with the initial {1, 1} -> [2, n]."""
    zT = zT.to(self.torch_device, torch.float64).contiguous()
    xR = xR.to(self.torch_device, torch.float64).contiguous()
    zR = zR.to(self.torch_device, torch.float64).contiguous()
    out = torch.empty((2, zT.numel()), dtype=torch.float64, device=self.torch_device)
    check(self.lib.airice_inice_focusing_device(self.handle, zT.numel(), zT.data_ptr(), xR.data_ptr(), zR.data_ptr(),
                                                ptr_array([out[0].data_ptr(), out[1].data_ptr()]), _stream_ptr(self.torch_device)))
    return out


def _inice_ladder_stats(self):
    """(pairs searching two roots, pairs searching one, fRaa evaluations, turning-depth steps) of the last in-ice solve"""
    v = (C.c_int64 * 4)()
    check(self.lib.airice_inice_ladder_stats(self.handle, v))
    return tuple(int(x) for x in v)


def _inice_quadrature_stats(self):
    v = (C.c_int64 * 2)()
    check(self.lib.airice_inice_quadrature_stats(self.handle, v))
    return int(v[0]), int(v[1])


class InIceTable:
    """IceRayTracing::MakeTable's grid (13 columns, GridZValueb[AntNum]) on the device, with batched GetInterpolatedValue."""

    def __init__(self, solver, handle):
        self.solver, self.handle = solver, handle
        info = (C.c_int64 * 3)()
        check(solver.lib.airice_inice_table_info(handle, info))
        self.n_x, self.n_z, self.points = int(info[0]), int(info[1]), int(info[2])

    def columns(self):
        out = np.empty((13, self.points), dtype=np.float64)
        for k in range(13):
            check(self.solver.lib.airice_inice_table_copy_column(self.handle, k, out[k].ctypes.data))
        return out

    def positions(self):
        px, pz = np.empty(self.n_x, dtype=np.float32), np.empty(self.n_z, dtype=np.float32)
        check(self.solver.lib.airice_inice_table_copy_positions(self.handle, px.ctypes.data, pz.ctypes.data))
        return px, pz

    def interp(self, x, z, par):
        s = self.solver
        x = x.to(s.torch_device, torch.float64).contiguous()
        z = z.to(s.torch_device, torch.float64).contiguous()
        out = torch.empty_like(x)
        check(s.lib.airice_inice_table_interp_device(s.handle, self.handle, x.numel(), x.data_ptr(), z.data_ptr(), int(par),
                                                     out.data_ptr(), _stream_ptr(s.torch_device)))
        return out

    def interp_host(self, x, z, par):
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.empty_like(x)
        check(self.solver.lib.airice_inice_table_interp_host(self.solver.handle, self.handle, x.size, x.ctypes.data, z.ctypes.data,
                                                             int(par), out.ctypes.data))
        return out

    def close(self):
        if self.handle:
            self.solver.lib.airice_inice_table_destroy(self.handle)     # holds no pointer into the context
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _inice_table_create(self, shower_hit_distance, shower_depth, zR, step_x=0.1, step_z=0.1, width_x=40.0, width_z=20.0):
    h = C.c_void_p()
    check(self.lib.airice_inice_table_create(self.handle, shower_hit_distance, shower_depth, zR, step_x, step_z, width_x, width_z,
                                             C.byref(h)))
    return InIceTable(self, h)


def _inice_two_rays_host(self, rx_depth, distance, tx_depth):
    n = int(rx_depth.shape[0])
    out = np.empty((_capi.INICE_RAYS_COLS, n), dtype=np.float64)
    ig = np.empty((2, n), dtype=np.int32)
    check(self.lib.airice_inice_two_rays_host(self.handle, n, _host_ptr(rx_depth), _host_ptr(distance), _host_ptr(tx_depth),
                                              _host_ptr(out), _host_ptr(ig)))
    return out, ig


def _ray_path(self, theta, h, depth_m, ice_m, max_points=None):
    """Ray-path polylines (SingleRayAirIceRefraction.C:226-299) of n rays on device tensors.
    max_points=None sizes the rows to the longest path.  -> (x [n, max_points], z [n, max_points], count [n] int32)"""
    theta = theta.to(self.torch_device, torch.float64).contiguous()
    h = h.to(self.torch_device, torch.float64).contiguous()
    n = theta.numel()
    count = torch.empty(n, dtype=torch.int32, device=self.torch_device)
    sp = _stream_ptr(self.torch_device)
    if max_points is None:
        check(self.lib.airice_ray_path_device(self.handle, n, theta.data_ptr(), h.data_ptr(), depth_m, ice_m, 0, None, None,
                                              count.data_ptr(), sp))
        max_points = int(count.max().item()) if n else 0
    x = torch.empty((n, max_points), dtype=torch.float64, device=self.torch_device)
    z = torch.empty((n, max_points), dtype=torch.float64, device=self.torch_device)
    check(self.lib.airice_ray_path_device(self.handle, n, theta.data_ptr(), h.data_ptr(), depth_m, ice_m, max_points,
                                          x.data_ptr() if max_points else None, z.data_ptr() if max_points else None,
                                          count.data_ptr(), sp))
    return x, z, count


def _ray_path_host(self, theta, h, depth_m, ice_m, max_points):
    n = int(theta.shape[0])
    x = np.empty((n, max_points), dtype=np.float64)
    z = np.empty((n, max_points), dtype=np.float64)
    count = np.empty(n, dtype=np.int32)
    check(self.lib.airice_ray_path_host(self.handle, n, _host_ptr(theta), _host_ptr(h), depth_m, ice_m, max_points,
                                        _host_ptr(x) if max_points else None, _host_ptr(z) if max_points else None,
                                        _host_ptr(count)))
    return x, z, count


AirIceSolver.ray_path = _ray_path
AirIceSolver.ray_path_host = _ray_path_host
AirIceSolver.inice_solve = _inice_solve
AirIceSolver.inice_solve_host = _inice_solve_host
AirIceSolver.inice_two_rays = _inice_two_rays
AirIceSolver.inice_two_rays_host = _inice_two_rays_host
AirIceSolver.inice_two_rays_att = _inice_two_rays_att
AirIceSolver.inice_two_rays_att_host = _inice_two_rays_att_host
AirIceSolver.inice_attenuation = _inice_attenuation
AirIceSolver.inice_focusing = _inice_focusing
AirIceSolver.inice_quadrature_stats = _inice_quadrature_stats
AirIceSolver.inice_ladder_stats = _inice_ladder_stats
AirIceSolver.inice_table_create = _inice_table_create


def _host_ptr(a):
    if isinstance(a, torch.Tensor):
        assert a.device.type == "cpu" and a.is_contiguous()
        return a.data_ptr()
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data
