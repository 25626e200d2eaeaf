"""Development probe: where the time of the shared-air 64-antenna tables goes (build + pack vs lookups)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver
S = AirIceSolver(os.path.join(ROOT, "tests", "golden", "Atmosphere.dat"))
n_ant = int(sys.argv[1]) if len(sys.argv) > 1 else 64
depths = [-(200.0 * (k + 1) / n_ant) for k in range(n_ant)]
def wall(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3, r
for rep in range(3):
    ms, Ts = wall(lambda: S.table_create_multi(depths, 3000.0))
    ms2, _ = wall(lambda: [T.close() for T in Ts])
    print("multi create %.2f ms, close %.2f ms" % (ms, ms2))
for rep in range(2):
    ms, T = wall(lambda: S.table_create(-200.0, 3000.0)); ms2, _ = wall(T.close)
    print("single create %.3f ms close %.3f" % (ms, ms2))
