"""Development probe: attenuation / focusing / in-ice table kernels against the golden fixture, by ray type."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver
S = AirIceSolver(None)
g = np.load(os.path.join(ROOT, "tests", "golden", "inice_att.npz"))
t = torch.from_numpy
for tag in "abc":
    A0, f = g["A0f_" + tag]
    out, att, ig, ty = S.inice_two_rays_att(t(g["rx"]), t(g["dist"]), t(g["tx"]), A0, f, want_type=True)
    att, ig, ty = att.cpu().numpy().T, ig.cpu().numpy().T, ty.cpu().numpy().T
    want = g["att_" + tag]
    d = np.abs(att - want)
    for k in (1, 2, 3, 4):
        m = (ty == k) & (ig == 1)
        if m.any():
            print(tag, "type", k, "n", m.sum(), "bit-equal", (att[m] == want[m]).sum(), ">1e-9:", (d[m] > 1e-9).sum(), ">1e-7:", (d[m] > 1e-7).sum(), "max", d[m].max())
m = g["foc"].shape[0]
got = S.inice_focusing(t(g["tx"][:m].copy()), t(g["dist"][:m].copy()), t(g["rx"][:m].copy())).cpu().numpy().T
want = g["foc"]
fin = ~np.isnan(want)
rel = np.abs(got - want) / np.abs(want)
o, ig, ty = S.inice_two_rays(t(g["rx"][:m].copy()), t(g["dist"][:m].copy()), t(g["tx"][:m].copy()), want_type=True)
ty = ty.cpu().numpy().T
print("focusing: nan equal", np.array_equal(np.isnan(got), np.isnan(want)), "ones equal", np.array_equal(got == 1, want == 1))
for k in (1, 2, 3, 4):
    mm = (ty == k) & fin & (want != 1)
    if mm.any(): print(" type", k, "n", mm.sum(), "max rel", rel[mm].max(), ">1e-9:", (rel[mm] > 1e-9).sum())
flip = (g["tx"][:m] > g["rx"][:m])
print(" flipped pairs:", flip.sum(), "max rel among flipped", rel[flip][fin[flip]].max() if flip.any() else None, "unflipped", rel[~flip][fin[~flip]].max())
print("quad stats", S.inice_quadrature_stats())
