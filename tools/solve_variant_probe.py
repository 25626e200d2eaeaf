"""Development probe: rebuild libairice_b200.so with each set of -D knobs (on the GPU box) and time the C4 solve."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = sys.argv[1:] or ["", "-DAIRICE_PEEL_TOP=1"]
code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, %r)
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD
S = AirIceSolver(os.path.join(%r, "tests", "golden", "Atmosphere.dat"))
n = 10_000_000
rng = np.random.default_rng(20260418)
h = rng.uniform(3001, 100000, n); ang = rng.uniform(90.2, 179.8, n)
d = (h - 3000 + 200) * np.tan((180 - ang) * 3.1415927 / 180)
dh, dd = torch.from_numpy(h * 100).cuda(), torch.from_numpy(d * 100).cuda()
out = torch.empty((9, n), dtype=torch.float64, device="cuda"); ok = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(5): S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=ok)
torch.cuda.synchronize(); ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=ok); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("solve 1e7: best %%.4f ms median %%.4f ms  checksum %%r" %% (min(ts), float(np.median(ts)), float(out[5].nan_to_num().sum())))
''' % (ROOT, ROOT)
for v in variants:
    env = dict(os.environ, AIRICE_EXTRA_NVCC=v)
    subprocess.check_call([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from airiceraytracing_b200.build import build; build(force=True)" % ROOT],
                          env=env, stdout=subprocess.DEVNULL)
    print("variant [%s]" % v, flush=True)
    subprocess.check_call([sys.executable, "-c", code])
