"""Development probe: shared-air multi-antenna table pass of alternative builds of the library (lib/variant_*.so)."""
import glob, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, time, numpy as np, torch
sys.path.insert(0, %r)
import airiceraytracing_b200._capi as capi
capi.LIB_PATH = sys.argv[1]
from airiceraytracing_b200 import AirIceSolver
S = AirIceSolver(os.path.join(%r, "tests", "golden", "Atmosphere.dat"))
n_ant = 64
depths = [-(200.0 * (k + 1) / n_ant) for k in range(n_ant)]
def wall(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3, r
best = 1e9
for rep in range(6):
    ms, Ts = wall(lambda: S.table_create_multi(depths, 3000.0))
    if rep == 5:
        x = Ts[17].columns(); chk = float(np.nansum(x.astype(np.float64)))
    for T in Ts: T.close()
    torch.cuda.synchronize()
    if rep: best = min(best, ms)
print(os.path.basename(sys.argv[1]), "64 reference-grid tables: best %%.2f ms (wall, incl. row kernels), checksum %%.12e" %% (best, chk))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "airiceraytracing_b200", "lib", "variant_*.so"))) + [os.path.join(ROOT, "airiceraytracing_b200", "lib", "libairice_b200.so")]:
    subprocess.run([sys.executable, "-c", code, lib])
