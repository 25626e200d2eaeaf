import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver
from oracle.ref import ATMOSPHERE, IceRayReference
S = AirIceSolver(ATMOSPHERE)
rng = np.random.default_rng(2024); n = 20000
z0, z1, x1 = rng.uniform(-1501, -1, n), rng.uniform(-201, -1, n), rng.uniform(1, 3001, n)
ref = IceRayReference().solve_batch(z0, x1, z1)
for rep in range(3):
    out, mask = S.inice_solve(torch.from_numpy(z0), torch.from_numpy(x1), torch.from_numpy(z1)); torch.cuda.synchronize()
    got = out.cpu().numpy().T
    fr, fg = ref[:, 8:12] != -1000, got[:, 8:12] != -1000
    bad = np.where((fr != fg).any(1))[0]
    print("rep", rep, "mismatch", len(bad), bad[:10])
for b in bad[:6]:
    print(b, z0[b], x1[b], z1[b], "ref flags", fr[b].astype(int), "got", fg[b].astype(int), "mask", int(mask[b]), "ref L", ref[b,19:23], "got L", got[b,19:23])
# single-pair re-run of the bad ones
if len(bad):
    o2, m2 = S.inice_solve(torch.from_numpy(z0[bad]), torch.from_numpy(x1[bad]), torch.from_numpy(z1[bad])); torch.cuda.synchronize()
    g2 = o2.cpu().numpy().T
    print("re-run alone flags:", (g2[:, 8:12] != -1000).astype(int)[:6].tolist())
