"""Short single-GPU workload for ncu captures: a few launches of each kernel on resident inputs."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD
ATM = os.path.join(ROOT, "tests", "golden", "Atmosphere.dat")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
S = AirIceSolver(ATM)
ONCE = os.environ.get("AIRICE_NCU_ONCE") == "1"     # one launch per kernel: keeps a --set full report small
rng = np.random.default_rng(20260418)
h = rng.uniform(3001, 100000, n); ang = rng.uniform(90.2, 179.8, n)
d = (h - 3000 + 200) * np.tan((180 - ang) * 3.1415927 / 180)
dh, dd = torch.from_numpy(h * 100).cuda(), torch.from_numpy(d * 100).cuda()
out = torch.empty((9, n), dtype=torch.float64, device="cuda"); ok = torch.empty(n, dtype=torch.uint8, device="cuda")
if which in ("all", "solve"):
    for _ in range(1 if ONCE else 3):
        S.solve(dh, dd, -20000., 300000., UNITS_CM_RAD, out=out, ok=ok)
if which in ("all", "table"):
    for _ in range(1 if ONCE else 3):
        S.table_build(-200., 3000., h_step=20., th_start=92., th_step=0.5)
    for _ in range(1 if ONCE else 2):
        S.table_build(-200., 3000.)
if which in ("all", "multi"):
    Ts = S.table_create_multi([-200.0 * (k + 1) / 8 for k in range(8)], 3000.)
    for T in Ts:
        T.close()
if which in ("all", "lookup"):
    T = S.table_create(-200., 3000.)
    o2 = torch.empty((9, n), dtype=torch.float64, device="cuda")
    for _ in range(1 if ONCE else 3):
        S.lookup(T, dh, dd, out=o2, ok=ok)
if which in ("all", "inice"):
    ni = n
    rng2 = np.random.default_rng(7)
    z0 = torch.from_numpy(rng2.uniform(-1501, -1, ni)).cuda(); z1 = torch.from_numpy(rng2.uniform(-201, -1, ni)).cuda()
    x1 = torch.from_numpy(rng2.uniform(1, 3001, ni)).cuda()
    for _ in range(0 if ONCE else 2):
        S.inice_solve(z0, x1, z1)
    S.inice_two_rays(z1, x1, z0)
if which in ("all", "path"):
    nr = 1024
    tr = torch.full((nr,), 170.0, dtype=torch.float64, device="cuda") - 40.0 * torch.rand(nr, device="cuda", dtype=torch.float64)
    hr = torch.full((nr,), 20000.0, dtype=torch.float64, device="cuda")
    mp = int(S.ray_path(tr, hr, -200.0, 3000.0, max_points=0)[2].max().item())
    for _ in range(1 if ONCE else 2):
        S.ray_path(tr, hr, -200.0, 3000.0, max_points=mp)
torch.cuda.synchronize()
print("done", which, n)
