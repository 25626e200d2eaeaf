"""GPU: the drop-in surfaces above the C ABI -- the source-compatible C++ namespace (a driver shaped like the
reference's RunMultiRayCode.C is compiled against include/MultiRayAirIceRefraction.{h,cc}) and the python wrapper's
libAirIceRayTracing.so (loaded by a ctypes stub with the reference's argtypes)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ATMOSPHERE, ATOL_ANGLE_DEG, ROOT, RTOL_DIST, golden

pytestmark = pytest.mark.gpu
PI_M = 3.1415927


@pytest.fixture(scope="module")
def driver_output(tmp_path_factory, solver):
    tmp = tmp_path_factory.mktemp("compat")
    exe = str(tmp / "driver")
    lib = os.path.join(ROOT, "airiceraytracing_b200", "lib")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "compat", "driver.cc"), "-o", exe, "-L" + lib, "-lairice_b200",
                           "-Wl,-rpath," + lib])
    g = golden("old_table.npz")
    qfile = str(tmp / "queries.txt")
    np.savetxt(qfile, np.stack([g["qh"], g["qt"]], axis=1), fmt="%.17g")
    out = subprocess.run([exe, ATMOSPHERE, qfile, str(tmp / "table0.airicetb")], capture_output=True, text=True, check=True).stdout
    rec = {}
    for line in out.splitlines():
        p = line.split()
        if not p:
            continue
        try:
            vals = [float(x) for x in p[1:]]
        except ValueError:
            continue
        rec.setdefault(p[0], []).append(vals)
    rec["_stdout"] = out
    # the same driver with the per-call table builds (MakeRayTracingTable builds in-ice antennas' tables together on first
    # use by default; AIRICE_EAGER_TABLES=1 builds each one in its own call, as the reference does)
    rec["_stdout_eager"] = subprocess.run([exe, ATMOSPHERE, qfile, str(tmp / "table0e.airicetb")], capture_output=True, text=True,
                                          check=True, env=dict(os.environ, AIRICE_EAGER_TABLES="1")).stdout
    return rec


def _close9(got, ref, what):
    got, ref = np.asarray(got), np.asarray(ref)
    for k in (0, 1, 2, 3, 5):
        assert abs(got[k] - ref[k]) <= RTOL_DIST * abs(ref[k]), (what, k, got[k], ref[k])
    for k in (4, 8):
        assert abs(got[k] - ref[k]) * 180 / PI_M <= ATOL_ANGLE_DEG, (what, k)
    for k in (6, 7):
        assert abs(got[k] - ref[k]) <= 1e-9, (what, k)


def test_driver_direct_solve_and_forward(driver_output, oracle):
    ok, ref = oracle.solve_cm(500000.0, 100000.0, -20000.0, 300000.0)
    d = driver_output["direct"][0]
    assert int(d[0]) == int(ok) == 1
    _close9(d[1:], ref, "direct")
    f = np.array(driver_output["forward"][0])
    want = oracle.forward(170.0, 20000.0, 3000.0, -200.0)
    assert np.allclose(f[[2, 3, 4, 5, 6, 7, 8, 9, 10, 16, 17]], want[[2, 3, 4, 5, 6, 7, 8, 9, 10, 16, 17]], rtol=RTOL_DIST, atol=0)
    assert np.abs(f[[11, 12, 13]] - want[[11, 12, 13]]).max() <= ATOL_ANGLE_DEG
    thR = 180 - np.degrees(np.arctan(1000.0 / (5000 - 3000 + 200))) * (np.pi / PI_M)
    a2i, _ = oracle.air2ice(5000.0, 1000.0, 3000.0, -200.0, 180 - (np.arctan(1000.0 / 2200.0) * (180.0 / PI_M)))
    got = np.array(driver_output["air2ice"][0])
    assert np.allclose(got[[1, 2, 3, 4, 5, 6, 7, 8, 9, 14, 15]], a2i[[1, 2, 3, 4, 5, 6, 7, 8, 9, 14, 15]], rtol=RTOL_DIST)
    assert np.abs(got[[10, 11, 16]] - a2i[[10, 11, 16]]).max() <= ATOL_ANGLE_DEG
    assert "main function parameters are 5000 1000 3000 -200" in driver_output["_stdout"]  # M.cc:1466
    # the mutable ice model reaches the kernels: A_ice = 1.775 changes the ice leg
    assert driver_output["direct_A1775"][0][1] != d[1]


def test_driver_per_argument_batch_equals_scalar_call(driver_output):
    """GetHorizontalDistanceToIntersectionPointBatch with NULL for the outputs the caller does not read: the columns that
    were asked for carry the scalar call's bits."""
    c = driver_output["direct_cols"][0]
    d = driver_output["direct"][0]
    assert c[0] == 0 and c[1] == d[0] == 1
    assert c[2] == d[5] and c[3] == d[6]          # launch angle, distance to the intersection point
    assert c[4] == 1 and 1.5 < c[5] < 3.2 and c[6] > 0


def test_driver_per_argument_table_batch_equals_scalar_call(driver_output):
    c = driver_output["table_cols"][0]
    t1 = driver_output["table1"][0]
    assert c[0] == 0 and c[1] == t1[0] and c[2] == t1[5]


def test_driver_medium_accessors_and_fresnel(driver_output, oracle):
    """GetB_air/GetC_air/Getnz_air/Getnz_ice and Refl/Trans_S/P of the source-compatible API against the oracle's medium."""
    med = np.array(driver_output["medium"])
    assert med.shape == (12, 5)
    consts = oracle.constants()
    for z, B, Cc, n_air, n_ice in med:
        assert n_air == oracle.nz_air(z), (z, n_air, oracle.nz_air(z))
        assert n_ice == oracle.nz_ice(-z / 100)
        assert B in consts["B_air"] and Cc in consts["C_air"]
        assert n_air == 1.0 + B * np.exp(-Cc * abs(z))
    # bottom and top layer; 23141.75 / 23141.76 m straddle ATMLAY[3] = 23141.7538 m (half-open layers, ">=" on the lower edge)
    assert med[0, 1] == consts["B_air"][0] and med[10, 1] == consts["B_air"][consts["max_layers"] - 1]
    lay = np.asarray(consts["atmlay_cm"]) / 100
    assert 23141.75 < lay[3] < 23141.76 and med[7, 1] == consts["B_air"][2] and med[8, 1] == consts["B_air"][3]
    n1, n2 = oracle.nz_air(3000.0), oracle.nz_ice(0.0)
    for th, rS, tS, rP, tP in driver_output["fresnel"]:
        sq = np.sqrt(1 - ((n1 / n2) * np.sin(th)) ** 2)
        num, den = n1 * np.cos(th) - n2 * sq, n1 * np.cos(th) + n2 * sq
        assert abs(rS - num / den) < 1e-14 and abs(tS - (1 + num / den)) < 1e-14
        nump, denp = n1 * sq - n2 * np.cos(th), n1 * sq + n2 * np.cos(th)
        assert abs(rP + nump / denp) < 1e-14 and abs(tP - (1 - nump / denp) * (n1 / n2)) < 1e-14


def test_driver_tables_and_lookup(driver_output, oracle):
    assert driver_output["tables"][0] == [2.0]  # antenna 2 shares antenna 0's depth -> two tables (RunMultiRayCode.C:38-52)
    assert driver_output["TotalHeightSteps"][0] == [49.0] and driver_output["TotalAngleSteps"][0] == [177.0]
    for ant, depth_cm in ((0, -20000.0), (1, -15000.0), (2, -20000.0)):
        t = oracle.table_build(depth_cm, 300000.0, 0.5, 92.0, 180.0, 2000.0)
        ok, ref = t.lookup_cm(500000.0, 100000.0, depth_cm, 300000.0)
        got = driver_output["table%d" % ant][0]
        assert int(got[0]) == int(ok) == 1
        g, r = np.array(got[1:]), ref
        # float table: our cells may differ from the oracle's by one float ulp, so compare at float precision
        assert np.allclose(g, r, rtol=3e-7, atol=0)
        t.free()
    assert driver_output["table0"][0] == driver_output["table2"][0]


def test_driver_deferred_table_builds_change_nothing(driver_output):
    """MakeRayTracingTable collects the in-ice antennas and builds their tables in one shared-air pass on first use; every
    number the driver prints (lookups, table columns compared against the batch build, persistence round trip) must be
    the one the per-call builds give."""
    assert driver_output["_stdout"] == driver_output["_stdout_eager"]
    assert "table0 1" in driver_output["_stdout"]
    # tables requested under A_ice = 1.775 and read after the model went back to 1.78, an antenna in air between them
    rc0, rc1, rc2, s0, s1, s2 = driver_output["icemodel_tables"][0]
    assert (rc0, rc1, rc2) == (0, 0, 0) and s0 > 0 and s2 > 0 and s1 == 0      # receiver in air: no optical path in ice


def test_driver_batch_tables_equal_per_antenna_tables(driver_output):
    """MakeRayTracingTables (all antennas in one pass over the grid, air walk shared) against MakeRayTracingTable."""
    rc, differ, cells, rc2, n = driver_output["multitables"][0]
    assert rc == 0 and differ == 0 and cells == 49 * 177 and rc2 == 0 and n == cells


def test_driver_old_table_and_idw(driver_output):
    g = golden("old_table.npz")
    assert driver_output["old_dims"][0] == [float(g["n_h"]), float(g["n_th"]), float(g["cols"].shape[1])]
    for c in range(9):
        got = np.array(driver_output["old_col%d" % c][0])
        ref = g["cols"][c]
        assert np.array_equal(got == -1000, ref == -1000), "sentinel cells differ in column %d" % c
        m = ref != -1000
        tol = ATOL_ANGLE_DEG if c in (4, 8) else None
        if tol:
            assert np.abs(got[m] - ref[m]).max() <= tol
        else:
            assert (np.abs(got[m] - ref[m]) / np.abs(ref[m])).max() <= RTOL_DIST
    q = np.array(driver_output["old_q"])
    ref = g["qv"]
    assert q.shape == ref.shape
    assert np.array_equal(q == -1000, ref == -1000)
    m = ref != -1000
    # IDW over our grid values (each within 1e-9 of the reference's): 1e-9 as well
    assert (np.abs(q[m] - ref[m]) / np.maximum(np.abs(ref[m]), 1e-300)).max() <= RTOL_DIST
    # the batched device lookup returns the scalar host function's bits
    assert driver_output["old_batch_differ"][0] == [0.0, float(ref.shape[0])]


def test_driver_table_persistence(driver_output):
    rcs, idx, ka, kb, same = driver_output["persist"][0]
    assert rcs == 0 and idx >= 0 and ka == 1 and kb == 1 and same == 1


def test_old_table_device_lookup_on_reference_grid_is_bit_exact(solver):
    """GetInterpolatedValue (M.cc:1700-1794) as a device kernel, fed the grid the REFERENCE built (golden): every query
    of the fixture, all nine parameters, bit for bit -- plus rounding edges, out-of-range queries and exact node hits."""
    import torch
    g = golden("old_table.npz")
    T = solver.oldtable_wrap(g["cols"], float(g["ice_cm"]) / 100, float(g["start_th"]), float(g["stop_th"]), float(g["step_h"]),
                             float(g["step_th"]))
    assert (T.n_h, T.n_th) == (int(g["n_h"]), int(g["n_th"]))
    for p in range(9):
        got = T.interp(torch.from_numpy(g["qh"]), torch.from_numpy(g["qt"]), p).cpu().numpy()
        assert np.array_equal(got.view(np.int64), g["qv"][:, p].copy().view(np.int64)), p
        assert np.array_equal(T.interp_host(g["qh"], g["qt"], p).view(np.int64), got.view(np.int64))
    assert T.interp(torch.empty(0, dtype=torch.float64), torch.empty(0, dtype=torch.float64), 1).numel() == 0
    T.close()


def test_old_table_device_build_matches_reference_grid(solver):
    """MakeTable's grid built and kept on the device (airice_oldtable_create) against the reference-built golden grid."""
    g = golden("old_table.npz")
    T = solver.oldtable_create(float(g["ice_cm"]) / 100, float(g["depth_cm"]) / 100, float(g["start_th"]), float(g["stop_th"]),
                               float(g["step_h"]), float(g["step_th"]))
    cols = T.columns()
    ref = g["cols"]
    assert cols.shape == ref.shape
    assert np.array_equal(cols == -1000, ref == -1000)
    m = ref != -1000
    for c in range(9):
        d = np.abs(cols[c][m[c]] - ref[c][m[c]])
        if c in (4, 8):
            assert d.max() <= ATOL_ANGLE_DEG
        elif c in (6, 7):
            assert d.max() <= 1e-9
        else:
            assert (d / np.abs(ref[c][m[c]])).max() <= RTOL_DIST
    ph, pt = T.positions()
    assert ph[0] == 3001.0 and ph[-1] == 100000.0 and pt[0] == 90.05 and pt[-1] == 179.95
    T.close()


def test_table_save_load_round_trip(solver, tmp_path):
    import torch
    T = solver.table_create(-200.0, 3000.0, h_step=500.0, th_start=90.1, th_step=0.25)
    path = tmp_path / "t.airicetb"
    T.save(path)
    assert os.path.getsize(path) == 64 + 4 * 11 * T.cells
    L = solver.table_load(path)
    assert (L.n_h, L.n_th) == (T.n_h, T.n_th)
    assert np.array_equal(L.columns().view(np.int32), T.columns().view(np.int32))
    for a, b in zip(L.row_ranges(), T.row_ranges()):
        assert np.array_equal(a, b)
    rng = np.random.default_rng(5)
    h = torch.from_numpy(rng.uniform(3001, 100000, 50000) * 100)
    d = torch.from_numpy(rng.uniform(1, 60000, 50000) * 100)
    o1, k1 = solver.lookup(T, h, d)
    o2, k2 = solver.lookup(L, h, d)
    assert torch.equal(k1, k2) and torch.equal(o1.view(torch.int64), o2.view(torch.int64))
    # damaged files are refused, not half-loaded
    raw = bytearray(open(path, "rb").read())
    for mutate in (lambda b: b.__setitem__(0, 0x58), lambda b: b.__setitem__(200, b[200] ^ 1), lambda b: b.extend(b"\0\0\0\0")):
        bad = bytearray(raw)
        mutate(bad)
        open(tmp_path / "bad.airicetb", "wb").write(bad)
        with pytest.raises(RuntimeError):
            solver.table_load(tmp_path / "bad.airicetb")
    open(tmp_path / "short.airicetb", "wb").write(raw[:1000])
    with pytest.raises(RuntimeError):
        solver.table_load(tmp_path / "short.airicetb")
    L.close()
    T.close()


def test_python_wrapper_library(solver):
    """Run in a fresh interpreter the way TraceIceToAir.py does: cwd holds Atmosphere.dat, ctypes stub next to the .so."""
    code = r'''
import ctypes, sys, numpy as np
sys.path.insert(0, %r)
from AirIceRayTracing import Py_TraceIceToAir, Py_TraceIceToAirBatch
g = np.load(%r)
ii = ctypes.c_double * 10
worst = [0.0, 0.0, 0.0]
for a, want in zip(g["args"], g["out"]):
    arr = ii(*([1.0] * 10))
    Py_TraceIceToAir(float(a[0]), float(a[1]), float(a[2]), float(a[3]), arr)
    got = np.array(list(arr))
    assert (got[0] == -1000) == (want[0] == -1000), (a, got, want)
    if want[0] != -1000:
        assert got[0] == want[0] and got[1] == want[1] and got[8] == 0 and got[9] == 0
        worst[0] = max(worst[0], (np.abs(got[[2, 3, 6]] - want[[2, 3, 6]]) / np.abs(want[[2, 3, 6]])).max())
        worst[1] = max(worst[1], np.abs(got[[4, 5, 7]] - want[[4, 5, 7]]).max())
    else:
        assert np.all(got == -1000)
sel = g["args"][:, 0] == g["args"][5, 0]
b = Py_TraceIceToAirBatch(float(g["args"][5, 0]), 3000.0, g["args"][sel, 2], g["args"][sel, 3])
assert b.shape == (sel.sum(), 10)
print("WORST", worst[0], worst[1])
''' % (os.path.join(ROOT, "airiceraytracing_b200", "pythonwrapper"), os.path.join(ROOT, "tests", "golden", "pywrap.npz"))
    out = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(ATMOSPHERE), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = [float(x) for x in out.stdout.strip().splitlines()[-1].split()[1:]]
    assert worst[0] <= RTOL_DIST and worst[1] <= ATOL_ANGLE_DEG
    assert "We have a solution!!!" in out.stdout and "We do NOT have a solution!!!" in out.stdout


def test_iceray_driver(tmp_path, solver):
    """SURVEY.md 8c known answers through the source-compatible IceRayTracing::IceRayTracing()."""
    exe = str(tmp_path / "iceray")
    lib = os.path.join(ROOT, "airiceraytracing_b200", "lib")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "compat", "iceray_driver.cc"), "-o", exe, "-L" + lib, "-lairice_b200",
                           "-Wl,-rpath," + lib])
    out = subprocess.run([exe], capture_output=True, text=True, check=True, cwd=str(tmp_path)).stdout   # no Atmosphere.dat here
    rows = [np.array([float(x) for x in line.split()[1:]]) for line in out.splitlines() if line.startswith("case")]
    g = golden("inice.npz")["out"]
    assert len(rows) == 3
    for k, r in enumerate(rows):
        assert np.array_equal(r[8:12] != -1000, g[k, 8:12] != -1000)
    assert abs(rows[0][19] - 0.80023300831165) < 1e-12 and abs(rows[0][20] - 0.759040146127282) < 1e-12
    assert abs(rows[0][0] - 27.380168714) < 1e-7 and abs(rows[0][1] - 25.8628862578) < 1e-7
    assert abs(rows[1][19] - 1.64977154594608) < 1e-12 and abs(rows[1][21] - 1.46757200401104) < 5e-9
    assert (rows[2][8:12] != -1000).sum() == 0
    # GetRayTracingSolutions(RxDepth, Distance, TxDepth, ...) of the same pairs against the plain-C oracle
    from oracle.ref import InIceOracle
    rays = [line.split()[1:] for line in out.splitlines() if line.startswith("rays")]
    assert len(rays) == 3
    cases = np.array([[-180, 100, -5], [-1000, 2000, -200], [-200, 1500, -150]], dtype=float)
    want, ig = InIceOracle().two_rays(cases[:, 2], cases[:, 1], cases[:, 0])
    for k, r in enumerate(rays):
        assert [int(r[0]), int(r[1])] == ig[k].tolist()
        v = np.array([float(x) for x in r[2:]]).reshape(2, 5)        # per ray: T, P, launch, receive, incidence
        for j in range(2):
            if ig[k, j]:
                assert abs(v[j, 0] - want[k, 0 + j]) <= 1e-9 * abs(want[k, 0 + j])
                assert abs(v[j, 1] - want[k, 2 + j]) <= 1e-9 * abs(want[k, 2 + j])
                assert abs(v[j, 2] - want[k, 4 + j]) <= 1e-9 and abs(v[j, 3] - want[k, 6 + j]) <= 1e-9
    # attenuation, focusing, attenuation length and the in-ice table through the same namespace
    from oracle.ref import IceRayReference, reference_available
    if reference_available("libiceray_ref.so"):
        ref = IceRayReference()
        _, att_w, _ = ref.two_rays_att(cases[:, 2], cases[:, 1], cases[:, 0], 1.0, 0.3)
        foc_w = ref.focusing(cases[:, 0], cases[:, 1], cases[:, 2])
        att = np.array([[float(x) for x in line.split()[1:]] for line in out.splitlines() if line.startswith("att ")])
        foc = np.array([[float(x) for x in line.split()[1:]] for line in out.splitlines() if line.startswith("focus")])
        assert np.abs(att - att_w).max() <= 1e-9 and np.abs(foc - foc_w).max() <= 1e-9 * np.abs(foc_w).max()
        ad = float([line.split()[1] for line in out.splitlines() if line.startswith("attdirect")][0])
        assert abs(ad - ref.attenuation(0, 1.0, 0.3, -180.0, -5.0, 0.0, 0.80023300831165)) <= 1e-12
    g = golden("inice_att.npz")
    col0 = np.array([float(x) for x in [line for line in out.splitlines() if line.startswith("tablecol0")][0].split()[1:]])
    ref0 = g["t1_cols"][0]
    assert np.array_equal(col0 == -1000, ref0 == -1000) and (np.abs(col0 - ref0)[ref0 != -1000] <= 1e-9 * np.abs(ref0[ref0 != -1000])).all()
    iv = [float(x) for x in [line for line in out.splitlines() if line.startswith("interp")][0].split()[1:]]
    assert iv[2] == -1000 and iv[0] > 0 and -5 < iv[1] < 1
