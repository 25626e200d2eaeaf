"""Development probe (not a test): parity + timing of the three kernels on one GPU, printed as text."""
import os, sys, time, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD, UNITS_M_DEG
from oracle.ref import ATMOSPHERE, Oracle

def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

S = AirIceSolver(ATMOSPHERE); O = Oracle(ATMOSPHERE)
print("gpu:", torch.cuda.get_device_name(0))
mo, ms = O.constants(), S.medium()
print("medium equal:", all(mo[k] == ms[k] for k in ("max_layers", "B_air", "C_air", "n0", "pi")))
print("fp64 peak TFLOP/s:", S.fp64_peak_tflops(), S.fp64_peak_tflops())

# ---- forward parity
rng = np.random.default_rng(5); n = 100000
th = rng.uniform(90.1, 180, n); h = rng.uniform(3001, 100000, n)
fo = O.forward_batch(th, h, 3000., -200.)
fs = S.forward(torch.from_numpy(th), torch.from_numpy(h), -200., 3000.).cpu().numpy().T
nan_o, nan_s = np.isnan(fo[:, 2]), np.isnan(fs[:, 1])
print("forward NaN-ness equal:", np.array_equal(nan_o, nan_s), int(nan_o.sum()))
okc = ~nan_o
for k, nm in {2: 'X', 3: 'Xair', 4: 'Xice', 5: 'opt', 8: 't', 16: 'geoair', 17: 'geoice'}.items():
    rel = np.abs(fs[okc, k - 1] - fo[okc, k]) / np.abs(fo[okc, k]); print(f"  {nm:7s} max rel {rel.max():.3e} >1e-9: {(rel > 1e-9).sum()}")
for k, nm in {12: 'inc', 13: 'recv'}.items():
    print(f"  {nm:7s} max abs deg {np.abs(fs[okc, k - 1] - fo[okc, k]).max():.3e}")
print(f"  TS/TP max abs {np.abs(fs[okc, 13] - fo[okc, 14]).max():.3e} {np.abs(fs[okc, 14] - fo[okc, 15]).max():.3e}")

# ---- solve parity
def pairs(seed, n, kind):
    rng = np.random.default_rng(seed)
    if kind == "loop":
        h = rng.uniform(3001, 100000, n); ang = rng.uniform(90.2, 179.8, n)
        d = (h - 3000 + 200) * np.tan((180 - ang) * 3.1415927 / 180)
    else:
        h = rng.uniform(3001, 23141, n); d = rng.uniform(1, 20000, n)
    return h * 100, d * 100
for kind, seed in (("loop", 20260418), ("coreas", 20260419)):
    hcm, dcm = pairs(seed, 100000, kind)
    oko, so = O.solve_cm_batch(hcm, dcm, -20000., 300000.)
    out, ok, nev = S.solve(torch.from_numpy(hcm), torch.from_numpy(dcm), -20000., 300000., UNITS_CM_RAD, nevals=True)
    ss, oks, nev = out.cpu().numpy().T, ok.cpu().numpy().astype(bool), nev.cpu().numpy()
    m = oko & oks
    dth = np.abs(ss[m, 4] - so[m, 4]) * 180 / 3.1415927
    print(f"   angle mismatches >1e-9 deg: {(dth > 1e-9).sum()}, >1e-12: {(dth > 1e-12).sum()}")
    print(f"solve[{kind}] flags equal {np.array_equal(oko, oks)} ({(oko != oks).sum()} differ); angle max {dth.max():.3e} deg, >1e-7: {(dth > 1e-7).sum()}, bit-equal {(ss[m, 4] == so[m, 4]).mean():.5f}; evals mean {nev.mean():.3f} max {nev.max()}")
    for k, nm in {0: 'optIce', 1: 'optAir', 2: 'geoIce', 3: 'geoAir', 5: 'Xair'}.items():
        rel = np.abs(ss[m, k] - so[m, k]) / np.abs(so[m, k]); print(f"  {nm:7s} max rel {rel.max():.3e} >1e-9: {(rel > 1e-9).sum()}")
    print(f"  TS/TP max abs {np.abs(ss[m, 6] - so[m, 6]).max():.3e} {np.abs(ss[m, 7] - so[m, 7]).max():.3e}; recv max deg {np.abs(ss[m, 8] - so[m, 8]).max() * 57.3:.3e}")

# ---- table parity (README grid, 200 m step to keep oracle quick) + lookup
ot = O.table_build(-20000., 300000., 0.5, 92.0, 180.0, 200.0)
T = S.table_create(-200., 3000., h_top=100000., h_step=200., th_start=92., th_step=0.5, th_stop=180.)
tc, oc = T.columns(), ot.columns()
print("table dims", (T.n_h, T.n_th), (ot.n_h, ot.n_th), "float cols bit-equal frac", (tc == oc).mean(), "max rel", np.nanmax(np.abs(tc - oc) / np.maximum(np.abs(oc), 1e-30)))
rng = np.random.default_rng(2); n = 100000
hcm = rng.uniform(2900, 101000, n) * 100; ang = rng.uniform(90.2, 179.8, n)
dcm = (hcm - 300000 + 20000) * np.tan((180 - ang) * 3.1415927 / 180)
Tw = S.table_wrap(torch.from_numpy(oc), ot.n_h, ot.n_th, ot.loop_stop_h, ot.height_step)   # oracle-built table -> our lookup
oko, lo = ot.lookup_cm_batch(hcm, dcm, -20000., 300000.)
out, ok = S.lookup(Tw, torch.from_numpy(hcm), torch.from_numpy(dcm)); ls, oks = out.cpu().numpy().T, ok.cpu().numpy().astype(bool)
m = oko & oks
print("lookup flags equal", np.array_equal(oko, oks), (oko != oks).sum(), "solved", oko.mean(), "max rel", (np.abs(ls[m] - lo[m]) / np.maximum(np.abs(lo[m]), 1e-300)).max())
f1, l1 = Tw.row_ranges()
idx_ok = all(ot.find_rows(hh)[0][:2] == [int(f1[ot.n_h - int(np.floor((hh - 3000) / 200.)) - 1]), int(l1[ot.n_h - int(np.floor((hh - 3000) / 200.)) - 1])] for hh in (3000., 3100., 5000., 50000., 99999., 100000.))
print("row ranges match FindClosestAirTxHeight:", idx_ok)

# ---- timing
n = 10_000_000
hcm, dcm = pairs(20260418, n, "loop")
dh, dd = torch.from_numpy(hcm).cuda(), torch.from_numpy(dcm).cuda()
out = torch.empty((9, n), dtype=torch.float64, device="cuda"); okb = torch.empty(n, dtype=torch.uint8, device="cuda")
best, med = ev_time(lambda: S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=okb))
print(f"solve 1e7 loop-shape: best {best:.3f} ms median {med:.3f} ms -> {n / best * 1e3:.4e} solves/s")
perm = torch.argsort(dd / (dh - 280000.0))
dh2, dd2 = dh[perm].contiguous(), dd[perm].contiguous()
best, med = ev_time(lambda: S.solve(dh2, dd2, -20000., 300000., UNITS_CM_RAD, out=out, ok=okb))
print(f"solve 1e7 angle-sorted: best {best:.3f} ms -> {n / best * 1e3:.4e} solves/s")
hcm, dcm = pairs(20260419, n, "coreas")
dh, dd = torch.from_numpy(hcm).cuda(), torch.from_numpy(dcm).cuda()
best, med = ev_time(lambda: S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=okb))
print(f"solve 1e7 coreas-like: best {best:.3f} ms -> {n / best * 1e3:.4e} solves/s")
ph = torch.from_numpy(hcm).pin_memory(); pd = torch.from_numpy(dcm).pin_memory()
po = torch.empty((9, n), dtype=torch.float64).pin_memory(); pk = torch.empty(n, dtype=torch.uint8).pin_memory()
t0 = time.perf_counter(); S.solve_host(ph, pd, -20000., 300000., UNITS_CM_RAD, out=po, ok=pk); t1 = time.perf_counter()
t0 = time.perf_counter(); S.solve_host(ph, pd, -20000., 300000., UNITS_CM_RAD, out=po, ok=pk); t1 = time.perf_counter()
print(f"solve_host 1e7 pinned e2e: {(t1 - t0) * 1e3:.2f} ms -> {n / (t1 - t0):.4e} solves/s; agrees with device: {np.array_equal(po.numpy()[:, :1000], out.cpu().numpy()[:, :1000], equal_nan=True)}")
del out, po
# table: reference grid 9701x900, 17 f64 columns; and float-only
for name, kw in (("reference grid 10m x 0.1deg", dict(h_step=10., th_start=90.1, th_step=0.1)), ("README grid 20m x 0.5deg", dict(h_step=20., th_start=92., th_step=0.5))):
    n_h, n_th = S.table_dims(-200., 3000., **kw); cells = n_h * n_th
    o64 = torch.empty((17, cells), dtype=torch.float64, device="cuda"); o32 = torch.empty((11, cells), dtype=torch.float32, device="cuda")
    best, _ = ev_time(lambda: S.table_build(-200., 3000., out64=o64, **kw))
    print(f"table {name}: {n_h}x{n_th}={cells} cells, 17 f64 cols: {best:.3f} ms -> {cells / best * 1e3:.4e} cells/s, {cells * 136 / best / 1e6:.1f} GB/s stores")
    best, _ = ev_time(lambda: S.table_build(-200., 3000., columns64=None, want32=True, out32=o32, **kw))
    print(f"table {name}: 11 f32 cols only: {best:.3f} ms -> {cells / best * 1e3:.4e} cells/s")
    del o64, o32
# lookup timing on reference grid table
T2 = S.table_create(-200., 3000.)
n = 10_000_000
hcm, dcm = pairs(20260418, n, "loop")
dh, dd = torch.from_numpy(hcm).cuda(), torch.from_numpy(dcm).cuda()
out = torch.empty((9, n), dtype=torch.float64, device="cuda"); okb = torch.empty(n, dtype=torch.uint8, device="cuda")
best, _ = ev_time(lambda: S.lookup(T2, dh, dd, out=out, ok=okb))
print(f"lookup 1e7 on 9701x900 table: {best:.3f} ms -> {n / best * 1e3:.4e} lookups/s; solved {okb.float().mean().item():.4f}")
