"""Generates tests/golden/inice_att.npz by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/libiceray_ref.so, built by
oracle/Makefile from /root/reference/IceRayTracing.cc + the GSL stand-in whose QAGS is checked against QUADPACK in
tests/test_oracle.py): SURVEY.md 8f-4 -- attenuation, focusing factor, in-ice interpolation table.

    python tests/golden/make_golden_inice_att.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref import IceRayReference  # noqa: E402


def main():
    ref = IceRayReference()
    rng = np.random.default_rng(404)
    n = 3000
    tx = -rng.uniform(1, 1500, n)
    rx = -rng.uniform(1, 200, n)
    dist = rng.uniform(1, 3000, n)
    tx[:4] = [-180.0, -1000.0, -200.0, -100.0]
    dist[:4] = [100.0, 2000.0, 1500.0, 100.0]
    rx[:4] = [-5.0, -200.0, -150.0, -100.0]
    sets = {}
    for tag, (A0, f) in {"a": (1.0, 0.3), "b": (1.0, 0.1), "c": (2.5, 1.7)}.items():
        out, att, ig = ref.two_rays_att(rx, dist, tx, A0, f)
        sets["out_" + tag], sets["att_" + tag], sets["ig_" + tag] = out, att, ig
        sets["A0f_" + tag] = np.array([A0, f])
    # the three attenuation entry points on the rays of the 29-slot solution
    sol = ref.solve_batch(tx, dist, rx)
    kinds = []
    for i in range(600):
        o = sol[i]
        for kind, (recv, L, zm) in enumerate(((8, 19, None), (9, 20, None), (10, 21, 23))):
            if o[recv] != -1000:
                zmax = o[zm] if zm is not None else 0.0
                kinds.append([kind, tx[i], rx[i], zmax, o[L], ref.attenuation(kind, 1.0, 0.3, tx[i], rx[i], zmax, o[L])])
    m = 500
    foc = ref.focusing(tx[:m], dist[:m], rx[:m])
    # a small table: 1 m x 1 m steps over 12 m x 8 m around a shower at (60 m, -30 m), receiver at -20 m; and one that
    # touches the surface clamp (shower depth -6 m)
    tabs = {}
    for tag, (hit, dep, zR) in {"t1": (60.0, -30.0, -20.0), "t2": (3.0, -6.0, -50.0)}.items():
        cols, px, pz = ref.make_table(hit, dep, zR, step_x=1.0, step_z=1.0, width_x=12.0, width_z=8.0)
        qx = rng.uniform(px[0] - 1, px[-1] + 1, 300)
        qz = rng.uniform(pz[0] - 1, pz[-1] + 1, 300)
        qx[:4] = [px[0], px[-1], px[3], 0.5 * (px[2] + px[3])]
        qz[:4] = [pz[0], pz[-1], pz[2], pz[1]]
        qv = np.stack([ref.interp(qx, qz, p) for p in range(13)], axis=1)
        tabs.update({tag + "_cols": cols, tag + "_px": px, tag + "_pz": pz, tag + "_qx": qx, tag + "_qz": qz, tag + "_qv": qv,
                     tag + "_args": np.array([hit, dep, zR])})
    np.savez_compressed(os.path.join(HERE, "inice_att.npz"), tx=tx, rx=rx, dist=dist, att_calls=np.array(kinds), foc=foc, **sets, **tabs)
    print("written inice_att.npz:", {k: v.shape for k, v in tabs.items() if k.endswith("cols")}, len(kinds), "attenuation calls")


if __name__ == "__main__":
    main()
