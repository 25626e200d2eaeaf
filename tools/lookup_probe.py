"""Development probe: lookup kernel time on the reference grid, 1e7 random queries."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver
S = AirIceSolver(os.path.join(ROOT, "tests", "golden", "Atmosphere.dat"))
n = 10_000_000
rng = np.random.default_rng(20260418)
h = rng.uniform(3001, 100000, n); ang = rng.uniform(90.2, 179.8, n)
d = (h - 3000 + 200) * np.tan((180 - ang) * 3.1415927 / 180)
dh, dd = torch.from_numpy(h * 100).cuda(), torch.from_numpy(d * 100).cuda()
out = torch.empty((9, n), dtype=torch.float64, device="cuda"); ok = torch.empty(n, dtype=torch.uint8, device="cuda")
T = S.table_create(-200., 3000.)
for mode in ("", "1"):
    os.environ["AIRICE_LOOKUP_LITERAL"] = mode
    for _ in range(3): S.lookup(T, dh, dd, out=out, ok=ok)
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); S.lookup(T, dh, dd, out=out, ok=ok); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("lookup 1e7 (%s): best %.3f ms median %.3f ms, solved %.4f checksum %.12e" % (
        "literal search" if mode else "position table", min(ts), float(np.median(ts)), float(ok.float().mean()), float(out[5][ok.bool()].sum())))
