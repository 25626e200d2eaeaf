"""Generates the golden vectors in this directory by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref, built by
oracle/Makefile from /root/reference + oracle/gsl_standin).  Run once in the build container:

    python tests/golden/make_golden.py

The reference ships no tests and no expected outputs, so these files ARE the pin: the oracle
restatement, the host build of the device math and the CUDA kernels are all checked against them.
Everything is seeded; the inputs are stored next to the outputs so the fixtures are self-contained.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref import IceRayReference, PyWrapReference, Reference, old_interp, old_make_table  # noqa: E402

PI_M = 3.1415927
ICE_CM, DEPTH_CM = 300000.0, -20000.0


def raypath_from_cli(dst, argv4):
    import re
    import shutil
    import subprocess
    import tempfile
    exe = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "SingleRayAirIceRefraction")
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copy(os.path.join(HERE, "Atmosphere.dat"), tmp)
        out = subprocess.run([exe] + ["%g" % v for v in argv4], cwd=tmp, capture_output=True, text=True, check=True).stdout
        pts = np.loadtxt(os.path.join(tmp, "RayPathinAirnIce.txt"))
    total = float(re.search(r"Multiple Layer fitting is ([-0-9.eE+]+)", out).group(1))
    # the file holds "index x z" printed with 6 significant digits
    np.savez_compressed(dst, argv=np.array(argv4, dtype=float), printed_total_x_air=total, x=pts[:, 1].astype(np.float32),
                        z=pts[:, 2].astype(np.float32), stdout=out)


def main():
    ref = Reference()
    c = ref.constants()
    np.savez(os.path.join(HERE, "constants.npz"), max_layers=c["max_layers"], atmlay_cm=c["atmlay_cm"], B_air=c["B_air"],
             C_air=c["C_air"], pi=c["pi"], n0=c["n0"], n_air_3000=ref.nz_air(3000.0), n_ice_0=ref.nz_ice(0.0),
             n_ice_200=ref.nz_ice(200.0))

    # ---- forward cells (GetRayTracingSolutions): BASELINE config 1 first, edge cells, then random ones
    rng = np.random.default_rng(101)
    th = [170.0, 90.1, 90.2, 91.0, 91.5, 95.0, 135.0, 179.9, 180.0, 90.1, 180.0, 100.0, 100.0, 100.0, 100.0]
    hh = [20000.0, 100000.0, 3010.0, 3010.0, 50000.0, 3000.5, 3217.48275, 8363.53902, 23141.7538, 23141.03, 3001.0,
          3217.0, 3218.0, 8363.0, 8364.0]
    th = np.concatenate([th, rng.uniform(90.1, 180.0, 3000), rng.uniform(90.1, 92.0, 500)])
    hh = np.concatenate([hh, rng.uniform(3001.0, 100000.0, 3000), rng.uniform(3001.0, 100000.0, 500)])
    fwd = ref.forward_batch(th, hh, 3000.0, -200.0, True)
    fwd_air = ref.forward_batch(th[:400], hh[:400] + 100.0, 3000.0, 50.0, False)  # receiver 50 m above the surface
    np.savez_compressed(os.path.join(HERE, "forward.npz"), theta=th, h=hh, ice=3000.0, depth=-200.0, out=fwd,
                        out_air=fwd_air, depth_air=50.0)

    # ---- direct solves (GetHorizontalDistanceToIntersectionPoint, cm/rad API)
    rng = np.random.default_rng(20260418)
    n = 3000
    h = rng.uniform(3001, 100000, n)
    ang = rng.uniform(90.2, 179.8, n)
    d = (h - 3000 + 200) * np.tan((180 - ang) * PI_M / 180)  # RunMultiRayCode_loop.C:88-96
    rng = np.random.default_rng(20260419)
    h2 = rng.uniform(3001, 23141, n)
    d2 = rng.uniform(1, 20000, n)
    # hand-picked: README example, config-1 inverse, near-vertical, short range, near-horizon (unsolvable), tiny d
    h3 = np.array([5000.0, 20000.0, 3001.0, 3100.0, 3050.0, 99999.0, 10000.0, 4000.0, 3500.0, 60000.0])
    d3 = np.array([1000.0, 3018.9072284385093, 0.5, 50.0, 99.0, 1.0, 400000.0, 120000.0, 5.0, 900000.0])
    hcm = np.concatenate([h, h2, h3]) * 100
    dcm = np.concatenate([d, d2, d3]) * 100
    ok, out = ref.solve_cm_batch(hcm, dcm, DEPTH_CM, ICE_CM)
    np.savez_compressed(os.path.join(HERE, "solve.npz"), h_cm=hcm, d_cm=dcm, depth_cm=DEPTH_CM, ice_cm=ICE_CM, ok=ok,
                        out=out)
    # receiver in air (depth >= 0 branch, M.cc:1472-1476)
    ok_a, out_a = ref.solve_cm_batch(hcm[:500], dcm[:500], 5000.0, ICE_CM)
    np.savez_compressed(os.path.join(HERE, "solve_air.npz"), h_cm=hcm[:500], d_cm=dcm[:500], depth_cm=5000.0,
                        ice_cm=ICE_CM, ok=ok_a, out=out_a)

    # ---- forward table on a coarse grid (README angles, 2 km height step) + lookups into it
    grid = dict(angle_step=0.5, angle_start=92.0, angle_stop=180.0, height_step=2000.0)
    ref.clear_tables()
    ref.set_grid(**grid)
    ref.make_table(DEPTH_CM, ICE_CM)
    info = ref.table_info()
    cols = ref.table_columns()
    rng = np.random.default_rng(77)
    nq = 3000
    hq = rng.uniform(2900, 101000, nq)
    aq = rng.uniform(90.2, 179.8, nq)
    dq = (hq - 3000 + 200) * np.tan((180 - aq) * PI_M / 180)
    hq[:8] = [3000.0, 5000.0, 7000.0, 100000.0, 99000.0, 2999.0, 100001.0, 51000.0]  # exact rows, out of range
    okq, outq = ref.lookup_cm_batch(hq * 100, dq * 100, DEPTH_CM, ICE_CM)
    rows_idx = np.array([ref.find_rows(x)[0] for x in hq[(hq >= 3000) & (hq <= 100000)][:500]])
    rows_h = hq[(hq >= 3000) & (hq <= 100000)][:500]
    thd_idx = []
    for hx, dx in zip(rows_h[:300], dq[(hq >= 3000) & (hq <= 100000)][:300]):
        idx, _ = ref.find_rows(hx)
        if dx <= cols[1][idx[0]]:
            (i1, i2), cv = ref.find_thd(dx, idx[0], idx[1])
            thd_idx.append([hx, dx, idx[0], idx[1], i1, i2])
    np.savez_compressed(os.path.join(HERE, "table_lookup.npz"), n_h=info["n_h"], n_th=info["n_th"], cols=cols,
                        h_cm=hq * 100, d_cm=dq * 100, ok=okq, out=outq, rows_h=rows_h, rows_idx=rows_idx,
                        thd_idx=np.array(thd_idx), depth_cm=DEPTH_CM, ice_cm=ICE_CM, **grid)
    ref.set_grid()  # back to the shipped defaults

    # ---- old solve-per-cell table (MakeTable / GetInterpolatedValue) on a coarse grid
    og = dict(start_th=90.05, stop_th=179.95, step_h=4000.0, step_th=1.5)
    info, ocols = old_make_table(ref, ICE_CM, DEPTH_CM, **og)
    rng = np.random.default_rng(55)
    qh = rng.uniform(3001, 100000, 400)
    qt = rng.uniform(90.05, 179.95, 400)
    qh[:3] = [3001.0, 7001.0, 100000.0]
    qt[:3] = [90.05, 91.55, 179.95]
    qv = np.array([[old_interp(ref, float(a), float(b), p) for p in range(9)] for a, b in zip(qh, qt)])
    np.savez_compressed(os.path.join(HERE, "old_table.npz"), cols=ocols, n_h=info["n_h"], n_th=info["n_th"], qh=qh, qt=qt, qv=qv,
                        ice_cm=ICE_CM, depth_cm=DEPTH_CM, **og)

    # ---- in-ice solver IceRayTracing::IceRayTracing(0, z0, x1, z1): 29 outputs per pair
    rng = np.random.default_rng(3)
    n = 4000
    z0 = rng.uniform(-1501, -1, n)
    z1 = rng.uniform(-201, -1, n)
    x1 = rng.uniform(1, 3001, n)
    sw = rng.random(n) < 0.3                       # Tx shallower than Rx: exercises the flip (IceRayTracing.cc:631)
    z0, z1 = np.where(sw, z1, z0), np.where(sw, z0, z1)
    z0[:6] = [-180.0, -1000.0, -200.0, -100.0, -50.0, -5.0]     # SURVEY.md 8c known answers + same-depth + near surface
    x1[:6] = [100.0, 2000.0, 1500.0, 100.0, 20.0, 300.0]
    z1[:6] = [-5.0, -200.0, -150.0, -100.0, -50.0, -3.0]
    np.savez_compressed(os.path.join(HERE, "inice.npz"), z0=z0, x1=x1, z1=z1, out=IceRayReference().solve_batch(z0, x1, z1))

    # ---- python wrapper C ABI (Py_TraceIceToAir), metres/degrees, pi = 4 atan(1)
    pw = PyWrapReference()
    rng = np.random.default_rng(9)
    args = [(-10.0, 3000.0, 8050.0, 10000.0), (-200.0, 3000.0, 5000.0, 1000.0), (-100.0, 3000.0, 3200.0, 100.0),
            (-10.0, 3000.0, 3500.0, 90000.0), (-150.0, 2800.0, 20000.0, 3000.0)]
    for _ in range(60):
        args.append((-float(rng.uniform(1, 200)), 3000.0, float(rng.uniform(3010, 60000)), float(rng.uniform(10, 40000))))
    res = np.array([pw.py_trace(*a) for a in args])
    np.savez(os.path.join(HERE, "pywrap.npz"), args=np.array(args), out=res)
    # ---- BASELINE config 1: the reference CLI itself, `./SingleRayAirIceRefraction 200 170 20000 3000`, run in a scratch
    # directory next to a copy of Atmosphere.dat: the total horizontal distance it prints and the ray-path file it writes
    raypath_from_cli(os.path.join(HERE, "raypath_c1.npz"), (200, 170, 20000, 3000))
    raypath_from_cli(os.path.join(HERE, "raypath_low.npz"), (57.5, 135, 5000.25, 2800))
    print("golden vectors written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
