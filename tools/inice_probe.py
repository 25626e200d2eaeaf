"""Development probe: parity + timing of the in-ice kernel on one GPU."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver
from oracle.ref import ATMOSPHERE, IceRayReference
S = AirIceSolver(ATMOSPHERE)
rng = np.random.default_rng(2024); n = 20000
z0, z1, x1 = rng.uniform(-1501, -1, n), rng.uniform(-201, -1, n), rng.uniform(1, 3001, n)
ref = IceRayReference().solve_batch(z0, x1, z1)
out, mask = S.inice_solve(torch.from_numpy(z0), torch.from_numpy(x1), torch.from_numpy(z1)); got = out.cpu().numpy().T
fr, fg = ref[:, 8:12] != -1000, got[:, 8:12] != -1000
print("flag mismatches:", (fr != fg).any(1).sum(), "of", n)
for k, nm in {0: 'LangD', 1: 'LangR', 2: 'LangRa0', 3: 'LangRa1', 4: 'tD', 5: 'tR', 6: 'tRa0', 8: 'RangD', 9: 'RangR', 10: 'RangRa0', 18: 'inc', 19: 'LD', 20: 'LR', 21: 'LRa0', 23: 'zmax0', 25: 'pD', 27: 'pRa0'}.items():
    b = {0:0,4:0,8:0,19:0,25:0,1:1,5:1,9:1,18:1,20:1,2:2,6:2,10:2,21:2,23:2,27:2,3:3}[k]
    m = fr[:, b] & fg[:, b]
    a, r = got[m, k], ref[m, k]
    print(f"{nm:8s} n={m.sum():6d} max abs {np.abs(a-r).max():.3e} max rel {(np.abs(a-r)/np.maximum(np.abs(r),1e-300)).max():.3e}")
n = 2_000_000
rng = np.random.default_rng(7)
dz0 = torch.from_numpy(rng.uniform(-1501, -1, n)).cuda(); dz1 = torch.from_numpy(rng.uniform(-201, -1, n)).cuda(); dx1 = torch.from_numpy(rng.uniform(1, 3001, n)).cuda()
for _ in range(2): S.inice_solve(dz0, dx1, dz1)
torch.cuda.synchronize(); ts = []
for _ in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); o, mk = S.inice_solve(dz0, dx1, dz1); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f"in-ice solve {n}: best {min(ts):.2f} ms -> {n/min(ts)*1e3:.3e} solves/s; branch hist", torch.bincount(torch.tensor([bin(i).count('1') for i in range(16)], device='cuda')[mk.long()]).tolist())
import time
hz0, hz1, hx1 = dz0.cpu().numpy(), dz1.cpu().numpy(), dx1.cpu().numpy()
S.inice_solve_host(hz0, hx1, hz1)
t = time.perf_counter(); oh, mh = S.inice_solve_host(hz0, hx1, hz1); t = time.perf_counter() - t
print(f"in-ice solve through host buffers {n}: {t*1e3:.1f} ms -> {n/t:.3e} solves/s; equal to device path:",
      bool(np.array_equal(oh, o.cpu().numpy(), equal_nan=True) and np.array_equal(mh, mk.cpu().numpy())))
pz0, pz1, px1 = dz0.cpu().pin_memory(), dz1.cpu().pin_memory(), dx1.cpu().pin_memory()
po, pm = torch.empty((29, n), dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
S.inice_solve_host(pz0, px1, pz1, out=po, mask=pm)
t = time.perf_counter(); S.inice_solve_host(pz0, px1, pz1, out=po, mask=pm); t = time.perf_counter() - t
print(f"same with pinned buffers: {t*1e3:.1f} ms -> {n/t:.3e} solves/s; equal:", bool(torch.equal(po.nan_to_num(), o.cpu().nan_to_num())))
