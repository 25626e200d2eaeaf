"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD, UNITS_M_DEG
S = AirIceSolver(os.path.join(ROOT, "tests", "golden", "Atmosphere.dat"))
rng = np.random.default_rng(1)
n = 5001
h = rng.uniform(3001, 100000, n); ang = rng.uniform(90.2, 179.8, n)
d = (h - 3000 + 200) * np.tan((180 - ang) * 3.1415927 / 180)
th, td = torch.from_numpy(h * 100).cuda(), torch.from_numpy(d * 100).cuda()
S.solve(th, td, -20000., 300000., UNITS_CM_RAD)
S.solve(torch.from_numpy(h).cuda(), torch.from_numpy(d).cuda(), 50., 3000., UNITS_M_DEG)
S.solve_host(h * 100, d * 100, -20000., 300000., UNITS_CM_RAD)
S.forward(torch.from_numpy(ang).cuda(), torch.from_numpy(h).cuda(), -200., 3000.)
S.table_build(-200., 3000., h_step=2000., th_step=1.0, th_start=92.)
T = S.table_create(-200., 3000., h_step=500., th_step=0.5)
S.lookup(T, th, td)
S.lookup_host(T, h * 100, d * 100)
T.close()
T = S.table_create(-150., 3000., h_step=500., th_step=0.5)      # reuses the spare buffers of the table above
T.close()
Ts = S.table_create_multi([-200., -1., -37.5], 3000., h_step=500., th_step=0.5)   # shared-air pass, 3 antennas
S.lookup(Ts[1], th, td)
for T in Ts:
    T.close()
T = S.table_create(50., 3000., h_step=500., th_step=0.5)        # receiver in air: separate table + pack kernels
S.lookup(T, th, td)
T.close()
ni = 3001
z0, z1, x1 = rng.uniform(-1501, -1, ni), rng.uniform(-201, -1, ni), rng.uniform(1, 3001, ni)
S.inice_solve(torch.from_numpy(z0), torch.from_numpy(x1), torch.from_numpy(z1))
S.inice_solve_host(z0, x1, z1)
S.inice_two_rays(torch.from_numpy(z1), torch.from_numpy(x1), torch.from_numpy(z0), want_type=True)
S.inice_two_rays_host(z1, x1, z0)
S.ray_path(torch.from_numpy(ang[:33]), torch.from_numpy(h[:33]), -200., 3000.)
S.ray_path_host(ang[:5], h[:5], -57.5, 2800., 700)
S.solve_multi(th[:1000], torch.stack([td[:1000], td[:1000] * 0.9]), [-20000., -5000.], 300000., UNITS_CM_RAD)
torch.cuda.synchronize()
print("sanitize target done")
