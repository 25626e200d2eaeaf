"""CPU: the N>1 path (index sharding + the single gather) with world_size 2 and 3 over gloo.  The per-shard compute
is the oracle here; on the GPU box the same functions wrap AirIceSolver (see bench.py and test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ATMOSPHERE, ROOT

PI_M = 3.1415927


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from airiceraytracing_b200.dist import gather_columns, pairs_sharded, shard_range, solve_sharded, table_sharded
    from oracle.ref import InIceOracle, Oracle
    o = Oracle(ATMOSPHERE)
    rng = np.random.default_rng(99)
    h = rng.uniform(3001, 100000, n)
    ang = rng.uniform(95, 179.8, n)
    d = (h - 3000 + 200) * np.tan((180 - ang) * PI_M / 180)

    def solve_fn(hs, ds):
        ok, out = o.solve_cm_batch(hs.numpy() * 100, ds.numpy() * 100, -20000.0, 300000.0)
        return torch.from_numpy(out.T.copy()), torch.from_numpy(ok.astype(np.uint8))

    full, ok = solve_sharded(solve_fn, torch.from_numpy(h), torch.from_numpy(d))
    root_only, _ = solve_sharded(solve_fn, torch.from_numpy(h), torch.from_numpy(d), dst=0)

    t = o.table_build(-20000.0, 300000.0, 2.0, 92.0, 180.0, 9000.0)   # 11 rows x 45 angles: ragged over 2 and 3 ranks
    cols = torch.from_numpy(t.columns())

    def build_fn(r0, r1):
        return cols[:, r0 * t.n_th:r1 * t.n_th].clone()

    table = table_sharded(build_fn, t.n_h, t.n_th)

    # the in-ice two-ray selection through the generic per-pair sharding (three inputs, two flag rows)
    io = InIceOracle()
    m = min(n, 60)
    tx, rx, dist_m = -rng.uniform(1, 1500, m), -rng.uniform(1, 200, m), rng.uniform(1, 3000, m)

    def rays_fn(rxs, dss, txs):
        o10, ig = io.two_rays(rxs.numpy(), dss.numpy(), txs.numpy())
        return torch.from_numpy(o10.T.copy()), torch.from_numpy(ig.T.copy())

    rays, ign = pairs_sharded(rays_fn, torch.from_numpy(rx), torch.from_numpy(dist_m), torch.from_numpy(tx))
    if rank == 0:
        ok_ref, ref = o.solve_cm_batch(h * 100, d * 100, -20000.0, 300000.0)
        q.put(dict(solve_equal=bool(np.array_equal(full.numpy().T, ref, equal_nan=True)),
                   flags_equal=bool(np.array_equal(ok.numpy().astype(bool), ok_ref)),
                   root_equal=bool(root_only is not None and torch.equal(root_only.nan_to_num(), full.nan_to_num())),
                   table_equal=bool(torch.equal(table.nan_to_num(), cols.nan_to_num())),
                   rays_equal=bool(np.array_equal(rays.numpy().T, io.two_rays(rx, dist_m, tx)[0], equal_nan=True) and
                                   np.array_equal(ign.numpy().T, io.two_rays(rx, dist_m, tx)[1])),
                   ranges=[shard_range(n, r, world) for r in range(world)]))
    else:
        assert root_only is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 500), (2, 1)])
def test_sharded_solve_and_table_reassemble_in_caller_order(oracle_built, world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    res = None
    for _ in range(120):
        try:
            res = q.get(timeout=1)
            break
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res is not None
    assert res["solve_equal"] and res["flags_equal"] and res["root_equal"] and res["table_equal"] and res["rays_equal"]
    rng = res["ranges"]
    assert rng[0][0] == 0 and rng[-1][1] == n and all(a[1] == b[0] for a, b in zip(rng[:-1], rng[1:]))


def test_shard_range_properties():
    from airiceraytracing_b200.dist import shard_range, shard_sizes
    for n in (0, 1, 7, 8, 9, 10_000_000, 9701):
        for world in (1, 2, 4, 8):
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            assert [shard_range(n, r, world)[0] for r in range(world)] == list(np.cumsum([0] + sizes[:-1]))
