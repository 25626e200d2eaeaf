"""Development probe: latency of the reference-shaped scalar calls (batch-of-1 through the host C ABI), with the mapped
block for small calls and with AIRICE_NO_MAPPED=1 (copy calls around the launch)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time, numpy as np
sys.path.insert(0, %r)
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD
S = AirIceSolver(os.path.join(%r, "tests", "golden", "Atmosphere.dat"))
h = np.array([500000.0]); d = np.array([100000.0])
out = np.empty((9, 1)); ok = np.empty(1, dtype=np.uint8)
T = S.table_create(-200.0, 3000.0, h_step=1000.0, th_start=92.0, th_step=0.5)
for name, fn in (("solve", lambda: S.solve_host(h, d, -20000.0, 300000.0, UNITS_CM_RAD, out=out, ok=ok)),
                 ("lookup", lambda: S.lookup_host(T, h, d, out=out, ok=ok))):
    for _ in range(200): fn()
    t0 = time.perf_counter()
    for _ in range(3000): fn()
    print("%%s scalar call: %%.1f us (mapped=%%s) launch angle %%r ok %%d" %% (name, (time.perf_counter() - t0) / 3000 * 1e6,
          os.environ.get("AIRICE_NO_MAPPED") is None, float(out[4, 0]), int(ok[0])))
''' % (ROOT, ROOT)
for env in ({}, {"AIRICE_NO_MAPPED": "1"}):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env))
