"""Development probe: time the in-ice solver of alternative builds of the library (lib/variant_*.so)."""
import glob, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
import airiceraytracing_b200._capi as capi
capi.LIB_PATH = sys.argv[1]
from airiceraytracing_b200 import AirIceSolver
S = AirIceSolver(os.path.join(%r, "tests", "golden", "Atmosphere.dat"))
n = 2_000_000
rng = np.random.default_rng(7)
dz0 = torch.from_numpy(rng.uniform(-1501, -1, n)).cuda(); dz1 = torch.from_numpy(rng.uniform(-201, -1, n)).cuda(); dx1 = torch.from_numpy(rng.uniform(1, 3001, n)).cuda()
for _ in range(2): S.inice_solve(dz0, dx1, dz1)
torch.cuda.synchronize(); ts = []
for _ in range(4):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); o, mk = S.inice_solve(dz0, dx1, dz1); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(os.path.basename(sys.argv[1]), "inice 2e6: best %%.2f ms -> %%.3e /s, checksum %%.10e mask %%d" %% (min(ts), n / min(ts) * 1e3, float(o[19].nan_to_num().sum()), int(mk.long().sum())))
''' % (ROOT, ROOT)
for lib in sorted(glob.glob(os.path.join(ROOT, "airiceraytracing_b200", "lib", "variant_*.so"))) + [os.path.join(ROOT, "airiceraytracing_b200", "lib", "libairice_b200.so")]:
    subprocess.run([sys.executable, "-c", code, lib])
