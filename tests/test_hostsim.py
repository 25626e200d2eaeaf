"""CPU: the product's device math (airice_core.cuh / airice_solve.cuh compiled for the host by tests/hostsim)
against the golden vectors and the oracle.  This is the GPU-less check that the closed forms, the Newton solve and
the bisection replay reproduce the reference; the same comparisons run against the real kernels in
test_gpu_parity.py."""
import numpy as np

from conftest import assert_forward_close, assert_solve_close, golden

PI_M = 3.1415927


def test_medium_is_bit_identical(hostsim, oracle):
    m, c = hostsim.medium(), oracle.constants()
    assert m["max_layers"] == c["max_layers"]
    assert m["B_air"] == c["B_air"] and m["C_air"] == c["C_air"] and m["pi"] == c["pi"]
    assert m["atmlay_cm"][:4] == c["atmlay_cm"][:4]


def test_forward_cells_match_golden(hostsim):
    g = golden("forward.npz")
    got = hostsim.forward(g["theta"], g["h"], float(g["ice"]), float(g["depth"]), True)
    assert_forward_close(got, g["out"], "forward golden")
    n = g["out_air"].shape[0]
    got = hostsim.forward(g["theta"][:n], g["h"][:n] + 100.0, float(g["ice"]), float(g["depth_air"]), False)
    assert_forward_close(got, g["out_air"], "forward golden (receiver in air)")


def test_solves_match_golden(hostsim):
    g = golden("solve.npz")
    ok, out, st = hostsim.solve_cm(g["h_cm"], g["d_cm"], float(g["depth_cm"]), float(g["ice_cm"]))
    assert_solve_close(ok, out, g["ok"], g["out"], PI_M, "solve golden")
    # the point of the design: ~2-3 distance evaluations instead of the reference's ~30
    assert st[:, 0].mean() < 3.0 and st[:, 0].max() <= 8
    assert st[:, 1].mean() < 0.05
    g = golden("solve_air.npz")
    ok, out, st = hostsim.solve_cm(g["h_cm"], g["d_cm"], float(g["depth_cm"]), float(g["ice_cm"]))
    assert_solve_close(ok, out, g["ok"], g["out"], PI_M, "solve golden (receiver in air)")


def test_solves_match_oracle_on_fresh_pairs(hostsim, oracle):
    rng = np.random.default_rng(4242)
    n = 20000
    h = rng.uniform(3001, 100000, n)
    ang = rng.uniform(90.2, 179.8, n)
    d = (h - 3000 + 200) * np.tan((180 - ang) * PI_M / 180)
    ok_r, ref = oracle.solve_cm_batch(h * 100, d * 100, -20000.0, 300000.0)
    ok, out, _ = hostsim.solve_cm(h * 100, d * 100, -20000.0, 300000.0)
    assert_solve_close(ok, out, ok_r, ref, PI_M, "fresh pairs")


def test_replayed_angle_is_the_bisection_midpoint_not_the_true_root(hostsim, oracle):
    """The reference's answer is ~5e-8 deg off the true root (bisection artefact): forward 170 deg,
    invert, expect the reference's 169.99999995110377 rather than 170."""
    f = oracle.forward(170.0, 20000.0, 3000.0, -200.0)
    ok, out, st = hostsim.solve_cm(np.array([2000000.0]), np.array([f[2] * 100]), -20000.0, 300000.0)
    assert ok[0]
    launch_deg = out[0, 4] * 180 / PI_M
    assert abs(launch_deg - 169.99999995110377) < 1e-9
    assert abs(st[0, 2] - 170.0) < 1e-9  # theta* of the Newton phase is the true root


def test_closed_form_replay_in_the_special_zones(hostsim, oracle):
    """The bisection replay is a closed form (cell index from theta*, stop level from the bracket width): check it where
    the bracket is unusual -- straight-line angles within 16 deg of horizontal (clamped lower end, 0.05-deg scan, brackets
    down to a fraction of a degree), distances no ray reaches (GSL walks to the upper end), near-vertical rays."""
    rng = np.random.default_rng(20261018)
    n = 12000
    h = rng.uniform(3001, 100000, n)
    ang = np.concatenate([rng.uniform(90.02, 106.2, n // 2), rng.uniform(90.0005, 90.2, n // 4), rng.uniform(179.0, 179.9999, n // 4)])
    d = (h - 3000 + 200) * np.tan((180 - ang) * PI_M / 180)
    d[::7] *= rng.uniform(1.5, 30.0, d[::7].size)          # beyond the reach of any ray from that height
    ok_r, ref = oracle.solve_cm_batch(h * 100, d * 100, -20000.0, 300000.0)
    ok, out, st = hostsim.solve_cm(h * 100, d * 100, -20000.0, 300000.0)
    assert np.array_equal(ok, ok_r)
    # launch angles are compared for EVERY pair here, solved or not: an unsolved pair still returns the end of GSL's walk
    fin = np.isfinite(ref[:, 4]) & np.isfinite(out[:, 4])
    assert np.array_equal(np.isfinite(ref[:, 4]), np.isfinite(out[:, 4]))
    dang = np.abs(out[fin, 4] - ref[fin, 4]) * 180 / PI_M
    assert (dang > 1e-7).mean() <= 1e-3 and dang.max() <= 2.5e-7
    assert_solve_close(ok, out, ok_r, ref, PI_M, "special zones")


def test_first_pass_of_the_two_pass_launch(hostsim):
    """airice_solve_theta_t<DEFER>: the first pass either returns the very angle of the complete solve or flags the pair
    for the second pass -- and flags few (the slow paths are rare, that is the premise of the two-pass launch)."""
    rng = np.random.default_rng(7)
    n = 60000
    h = rng.uniform(3001, 100000, n)
    ang = np.concatenate([rng.uniform(90.2, 179.8, n // 2), rng.uniform(90.001, 96.0, n // 4), rng.uniform(170.0, 179.999, n // 4)])
    d = (h - 3000 + 200) * np.tan((180 - ang) * PI_M / 180)
    d[::11] *= rng.uniform(1.2, 20.0, d[::11].size)        # unreachable distances: the walk to the upper end
    td, tf, hard = hostsim.solve_defer(h * 100, d * 100, -20000.0, 300000.0)
    easy = ~hard
    assert np.array_equal(td[easy].view(np.int64), tf[easy].view(np.int64))     # NaN results included, bit for bit
    assert np.isnan(td[hard]).all()
    assert hard[: n // 2].mean() < 0.015 and hard.mean() < 0.05
    assert hard.any()                                                          # the test does exercise the flag
