"""ncu target: one shared-air pass over 64 antennas' reference-grid tables (library path in argv[1], optional)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import airiceraytracing_b200._capi as capi
if len(sys.argv) > 1:
    capi.LIB_PATH = sys.argv[1]
import torch
from airiceraytracing_b200 import AirIceSolver
S = AirIceSolver(os.path.join(ROOT, "tests", "golden", "Atmosphere.dat"))
n_ant = 64
for rep in range(2):
    Ts = S.table_create_multi([-(200.0 * (k + 1) / n_ant) for k in range(n_ant)], 3000.0)
    for T in Ts:
        T.close()
torch.cuda.synchronize()
print("done")
