"""Development probe: tie / flag census of the batched solve against the plain-C oracle over many seeds (each seed: 1e5
random pairs of the C4 distribution + 2e4 in the clamped-bracket zone).  Prints one line per seed and the totals."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD
from oracle.ref import Oracle, ATMOSPHERE
S, O = AirIceSolver(ATMOSPHERE), Oracle(ATMOSPHERE)
PI_M = 3.1415927
seeds = range(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
tot = dict(n=0, ties=0, flags=0, worst=0.0, relmax=0.0)
for seed in seeds:
    rng = np.random.default_rng(1000 + seed)
    n = 100000
    h = np.concatenate([rng.uniform(3001, 100000, n), rng.uniform(3001, 100000, n // 5)])
    ang = np.concatenate([rng.uniform(90.2, 179.8, n), rng.uniform(90.05, 106.0, n // 5)])
    d = (h - 3000 + 200) * np.tan((180 - ang) * PI_M / 180)
    out, ok = S.solve(torch.from_numpy(h * 100).cuda(), torch.from_numpy(d * 100).cuda(), -20000.0, 300000.0, UNITS_CM_RAD)
    out, ok = out.cpu().numpy().T, ok.cpu().numpy().astype(bool)
    ok_r, ref = O.solve_cm_batch(h * 100, d * 100, -20000.0, 300000.0)
    flags = int((ok != ok_r).sum())
    m = ok_r & ok
    dang = np.abs(out[m, 4] - ref[m, 4]) * 180 / PI_M
    ties = int((dang > 1e-7).sum())
    g = dang <= 1e-7
    rel = np.abs(out[m][g][:, [0, 1, 2, 3, 5]] - ref[m][g][:, [0, 1, 2, 3, 5]]) / np.maximum(np.abs(ref[m][g][:, [0, 1, 2, 3, 5]]), 1e-300)
    print("seed %d: %d solves, %d flag differences, %d ties, max angle difference (non-tie) %.2e deg, max rel err %.2e" % (
        seed, int(m.sum()), flags, ties, dang[g].max(), rel.max()), flush=True)
    tot["n"] += int(m.sum()); tot["ties"] += ties; tot["flags"] += flags
    tot["worst"] = max(tot["worst"], float(dang[g].max())); tot["relmax"] = max(tot["relmax"], float(rel.max()))
print("total:", tot)
