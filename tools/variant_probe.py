"""Development probe: time the solve kernel of alternative builds of the library (lib/variant_*.so)."""
import glob, os, sys, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
import airiceraytracing_b200._capi as capi
capi.LIB_PATH = sys.argv[1]
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD
S = AirIceSolver(os.path.join(%r, "tests", "golden", "Atmosphere.dat"))
n = 10_000_000
rng = np.random.default_rng(20260418)
h = rng.uniform(3001, 100000, n); ang = rng.uniform(90.2, 179.8, n)
d = (h - 3000 + 200) * np.tan((180 - ang) * 3.1415927 / 180)
dh, dd = torch.from_numpy(h * 100).cuda(), torch.from_numpy(d * 100).cuda()
out = torch.empty((9, n), dtype=torch.float64, device="cuda"); ok = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3): S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=ok)
torch.cuda.synchronize(); ts = []
for _ in range(7):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=ok); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
T = S.table_create(-200., 3000.)
o2 = torch.empty((9, n), dtype=torch.float64, device="cuda")
for _ in range(3): S.lookup(T, dh, dd, out=o2, ok=ok)
torch.cuda.synchronize(); tl = []
for _ in range(7):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); S.lookup(T, dh, dd, out=o2, ok=ok); b.record(); torch.cuda.synchronize(); tl.append(a.elapsed_time(b))
print(os.path.basename(sys.argv[1]), "lookup 1e7: best %%.3f ms -> %%.3e /s, ok %%.4f checksum %%.10e" %% (min(tl), n / min(tl) * 1e3, ok.float().mean().item(), float(o2[5][ok.bool()].sum())))
n_h, n_th = S.table_dims(-200.0, 3000.0)
o32 = torch.empty((11, n_h * n_th), dtype=torch.float32, device="cuda")
for _ in range(3): S.table_build(-200.0, 3000.0, columns64=None, want32=True, out32=o32)
torch.cuda.synchronize(); tt = []
for _ in range(7):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); S.table_build(-200.0, 3000.0, columns64=None, want32=True, out32=o32); b.record(); torch.cuda.synchronize(); tt.append(a.elapsed_time(b))
print(os.path.basename(sys.argv[1]), "table 9701x900 f32: best %%.4f ms" %% min(tt))
S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=ok)
print(os.path.basename(sys.argv[1]), "solve 1e7: best %%.3f ms  -> %%.3e solves/s, ok %%.4f, checksum %%.10e" %% (min(ts), n / min(ts) * 1e3, ok.float().mean().item(), float(out[5][ok.bool()].sum())))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "airiceraytracing_b200", "lib", "variant_*.so"))) + [os.path.join(ROOT, "airiceraytracing_b200", "lib", "libairice_b200.so")]:
    subprocess.run([sys.executable, "-c", code, lib])
