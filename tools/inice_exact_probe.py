"""Development probe: how exactly the in-ice kernels reproduce the reference build (bits), on the golden pairs and on
random pairs against the oracle running on this box's host."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver
from oracle.ref import ATMOSPHERE, InIceOracle
S = AirIceSolver(ATMOSPHERE)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
g = np.load(os.path.join(ROOT, "tests", "golden", "inice.npz"))
rng = np.random.default_rng(4242)
sets = {"golden": (g["z0"], g["x1"], g["z1"], g["out"])}
z0, z1, x1 = rng.uniform(-1501, -1, n), rng.uniform(-201, -1, n), rng.uniform(1, 3001, n)
sets["random"] = (z0, x1, z1, InIceOracle().solve_batch(z0, x1, z1))
names = {0: 'LangD', 1: 'LangR', 2: 'LangRa0', 3: 'LangRa1', 4: 'tD', 5: 'tR', 6: 'tRa0', 7: 'tRa1', 8: 'RangD', 9: 'RangR', 10: 'RangRa0',
         11: 'RangRa1', 18: 'inc', 19: 'LD', 20: 'LR', 21: 'LRa0', 22: 'LRa1', 23: 'zmax0', 24: 'zmax1', 25: 'pD', 26: 'pR', 27: 'pRa0', 28: 'pRa1'}
br = {0:0,4:0,8:0,19:0,25:0,1:1,5:1,9:1,18:1,20:1,26:1,2:2,6:2,10:2,21:2,23:2,27:2,3:3,7:3,11:3,22:3,24:3,28:3}
for nm, (a0, ax, a1, ref) in sets.items():
    out, mask = S.inice_solve(torch.from_numpy(a0), torch.from_numpy(ax), torch.from_numpy(a1)); got = out.cpu().numpy().T
    fr, fg = ref[:, 8:12] != -1000, got[:, 8:12] != -1000
    same = (got.view(np.int64) == ref.view(np.int64)) | (np.isnan(got) & np.isnan(ref))
    cols = [k for k in range(29) if k not in (12, 13, 14, 15, 16, 17)]   # sub-times: unset in the reference when the branch is absent
    print(f"[{nm}] n={len(a0)} flag mismatches {(fr != fg).any(1).sum()}; pairs with all {len(cols)} compared slots bit-equal: {same[:, cols].all(1).sum()}")
    for k, label in names.items():
        m = fr[:, br[k]] & fg[:, br[k]]
        if not m.any(): continue
        a, r = got[m, k], ref[m, k]
        print(f"  {label:8s} n={m.sum():6d} bit-equal {same[m, k].sum():6d} max abs {np.abs(a-r).max():.3e} max rel {(np.abs(a-r)/np.maximum(np.abs(r),1e-300)).max():.3e}")
