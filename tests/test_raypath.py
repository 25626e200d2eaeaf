"""Ray-path emission (BASELINE config 1: `./SingleRayAirIceRefraction 200 170 20000 3000` writes RayPathinAirnIce.txt).
Chain of evidence: reference CLI output (tests/golden/raypath_*.npz, 6 printed digits) -> plain-C oracle (double) ->
host build of the device code / CUDA kernels."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden

DP = C.POINTER(C.c_double)


def oracle_path(oracle, theta, h, ice, depth_pos, max_points=40000):
    f = oracle.lib.oracle_ray_path
    f.restype = C.c_long
    f.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_long, DP, DP]
    x, z = np.zeros(max_points), np.zeros(max_points)
    n = f(C.cast(oracle._a, C.c_void_p), theta, h, ice, depth_pos, max_points, x.ctypes.data_as(DP), z.ctypes.data_as(DP))
    return x[:min(n, max_points)], z[:min(n, max_points)], n


def printed6(a, b):
    """b was printed with 6 significant digits"""
    return np.abs(a - b) <= 5.1e-6 * np.maximum(np.abs(b), 1e-30) + 1e-30


@pytest.mark.parametrize("name", ["raypath_c1.npz", "raypath_low.npz"])
def test_oracle_path_matches_reference_cli(oracle, name):
    g = golden(name)
    depth, theta, h, ice = g["argv"]
    x, z, n = oracle_path(oracle, theta, h, ice, depth)
    assert n == g["x"].size
    assert printed6(x[1:], g["x"][1:].astype(np.float64)).all() and x[0] == 0.0
    assert printed6(z, g["z"].astype(np.float64)).all()
    # the total the CLI prints is the x of the last air point
    n_ice = int(np.ceil(depth + 1))
    assert printed6(np.array([x[n - n_ice]]), np.array([float(g["printed_total_x_air"])])).all()
    if name == "raypath_c1.npz":
        assert n == 17206
        fwd = golden("forward.npz")["out"][0]                  # same ray through the forward tracer (unmodified M.cc)
        assert abs(x[-1] - fwd[2]) < 1e-6 and abs(x[n - n_ice] - fwd[3]) < 1e-6


def sim_path(hostsim, theta, h, ice, depth, max_points=40000):
    f = hostsim.lib.sim_ray_path
    f.restype = C.c_long
    f.argtypes = [C.c_double] * 4 + [C.c_long, DP, DP]
    x, z = np.zeros(max_points), np.zeros(max_points)
    n = f(theta, h, ice, depth, max_points, x.ctypes.data_as(DP), z.ctypes.data_as(DP))
    return x[:min(n, max_points)], z[:min(n, max_points)], n


CASES = [(170.0, 20000.0, 3000.0, 200.0), (135.0, 5000.25, 2800.0, 57.5), (95.0, 3100.0, 3000.0, 10.0),
         (179.9, 99999.0, 3000.0, 200.0), (120.0, 3000.0, 3000.0, 5.0), (150.0, 8363.5425, 3000.0, 100.0),
         (100.0, 23141.7538, 3217.48275, 1.0)]


@pytest.mark.parametrize("case", CASES)
def test_host_build_path_matches_oracle(hostsim, oracle, case):
    theta, h, ice, depth = case
    xo, zo, no = oracle_path(oracle, theta, h, ice, depth)
    xs, zs, ns = sim_path(hostsim, theta, h, ice, -depth)
    assert ns == no
    assert np.array_equal(zs, zo)
    assert np.abs(xs - xo).max() <= 1e-8 + 1e-10 * np.abs(xo).max()


def test_host_build_path_edge_cases(hostsim):
    # transmitter below the surface, outside every layer, or L >= 1: no ray
    assert sim_path(hostsim, 170.0, 2000.0, 3000.0, -200.0)[2] == 0
    assert sim_path(hostsim, 170.0, 2.0e5, 3000.0, -200.0)[2] == 0
    assert sim_path(hostsim, 90.0, 3100.0, 3000.0, -200.0)[2] == 0
    # depth >= 0 (no receiver in the ice): the path ends on the surface, no ice points
    x, z, n = sim_path(hostsim, 150.0, 5000.0, 3000.0, 50.0)
    assert z[-1] == 3000.0 and z[0] == 5000.0 and n == 1784 + 219      # two layers: 5000 -> 3217.48 -> 3000
    # truncation keeps the count
    x, z, n = sim_path(hostsim, 170.0, 20000.0, 3000.0, -200.0, max_points=100)
    assert n == 17206 and x.size == 100


@pytest.mark.gpu
def test_kernel_path_matches_oracle_and_cli(solver, oracle):
    import torch
    g = golden("raypath_c1.npz")
    rng = np.random.default_rng(12)
    theta = np.concatenate([[170.0], rng.uniform(91.0, 179.9, 63)])
    h = np.concatenate([[20000.0], rng.uniform(3001.0, 60000.0, 63)])
    h[5], theta[5] = 2000.0, 150.0                                   # below the surface: no ray
    x, z, count = solver.ray_path(torch.from_numpy(theta), torch.from_numpy(h), -200.0, 3000.0)
    x, z, count = x.cpu().numpy(), z.cpu().numpy(), count.cpu().numpy()
    assert count[0] == 17206 and count[5] == 0 and x.shape[1] == count.max()
    assert printed6(x[0, 1:17206], g["x"][1:].astype(np.float64)).all() and printed6(z[0, :17206], g["z"].astype(np.float64)).all()
    for r in range(theta.size):
        xo, zo, no = oracle_path(oracle, theta[r], h[r], 3000.0, 200.0, max_points=x.shape[1] + 8) if count[r] else (None, None, 0)
        if r == 5:
            assert np.isnan(x[r]).all()
            continue
        assert count[r] == no
        assert np.array_equal(z[r, :no], zo)
        assert np.abs(x[r, :no] - xo).max() <= 1e-8 + 1e-10 * np.abs(xo).max()
        assert np.isnan(x[r, no:]).all() and np.isnan(z[r, no:]).all()
    # host-buffer entry point, truncated rows, counts only
    xh, zh, ch = solver.ray_path_host(theta[:4], h[:4], -200.0, 3000.0, 1000)
    assert np.array_equal(ch, count[:4]) and np.array_equal(xh, x[:4, :1000], equal_nan=True) and np.array_equal(zh, z[:4, :1000], equal_nan=True)
    x0, z0, c0 = solver.ray_path(torch.from_numpy(theta), torch.from_numpy(h), -200.0, 3000.0, max_points=0)
    assert x0.shape == (64, 0) and np.array_equal(c0.cpu().numpy(), count)


@pytest.mark.gpu
def test_ray_path_host_multi_chunk_equals_device_path(solver):
    """airice_ray_path_host cuts a batch into chunks of 16M points that alternate between two streams; every chunk has its
    own plan scratch (a shared one was overwritten by the next chunk's plan kernel: round-1 advisor finding).  2 600 rays x
    17 300 points = 45M points = 3 chunks."""
    import torch
    rng = np.random.default_rng(12)
    n, mp_ = 2600, 17300
    theta = rng.uniform(120.0, 179.0, n)
    h = rng.uniform(15000.0, 20000.0, n)
    xd, zd, cd = solver.ray_path(torch.from_numpy(theta), torch.from_numpy(h), -200.0, 3000.0, max_points=mp_)
    xh, zh, ch = solver.ray_path_host(theta, h, -200.0, 3000.0, mp_)
    assert np.array_equal(ch, cd.cpu().numpy())
    assert np.array_equal(xh, xd.cpu().numpy(), equal_nan=True) and np.array_equal(zh, zd.cpu().numpy(), equal_nan=True)
    assert (ch > 12000).all()
