"""GPU, N > 1: the sharded entry points of airiceraytracing_b200.dist with the CUDA solver under NCCL -- results must be
the single-GPU call's bits.  Needs >= 2 visible GPUs (skipped otherwise; `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_dist.py -m gpu` runs it, the log of that run is kept under profiles/)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ATMOSPHERE, ROOT

pytestmark = pytest.mark.gpu
PI_M = 3.1415927


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD
    from airiceraytracing_b200.dist import PeerGather, pairs_sharded, solve_sharded, table_sharded
    S = AirIceSolver(ATMOSPHERE, device=rank)
    dev = torch.device("cuda", rank)
    rng = np.random.default_rng(99)
    h = rng.uniform(3001, 100000, n) * 100
    ang = rng.uniform(90.2, 179.8, n)
    d = (h - 300000 + 20000) * np.tan((180 - ang) * PI_M / 180)
    ht, dt = torch.from_numpy(h).to(dev), torch.from_numpy(d).to(dev)

    def solve_fn(hs, ds):
        return S.solve(hs, ds, -20000.0, 300000.0, UNITS_CM_RAD)

    res = {}
    full, ok = solve_sharded(solve_fn, ht, dt)                       # NCCL all-gather
    root, rok = solve_sharded(solve_fn, ht, dt, dst=0)               # NCCL gather to rank 0
    # peer-memory gather: every rank's kernel stores its shard straight into rank 0's buffer over NVLink
    pg = PeerGather(S, ncols=9, n_total=n, dst=0)
    pfull, pok = pg.solve(S, ht, dt, -20000.0, 300000.0, UNITS_CM_RAD)
    # replicated form: every rank ends with the whole block (own shard computed in place, pieces copied to the peers)
    pg2 = PeerGather(S, ncols=9, n_total=n, dst=None, chunks=3)
    rfull, rok = pg2.solve(S, ht, dt, -20000.0, 300000.0, UNITS_CM_RAD)
    torch.cuda.synchronize()
    rep_ok = bool(torch.equal(rfull.view(torch.int64), full.view(torch.int64)) and torch.equal(rok, ok))
    flags = [None] * world
    dist.all_gather_object(flags, rep_ok)
    kw = dict(h_step=500.0, th_start=90.1, th_step=0.1, th_stop=180.0)
    n_h, n_th = S.table_dims(-200.0, 3000.0, **kw)
    table = table_sharded(lambda r0, r1: S.table_build(-200.0, 3000.0, rows=(r0, r1), **kw)[0], n_h, n_th)
    # lookups through the generic per-pair sharding, table replicated
    T = S.table_create(-200.0, 3000.0, **kw)
    lo, lk = pairs_sharded(lambda a, b: S.lookup(T, a, b), ht, dt)
    if rank == 0:
        one, ok1 = solve_fn(ht, dt)
        whole, _ = S.table_build(-200.0, 3000.0, **kw)
        lo1, lk1 = S.lookup(T, ht, dt)
        torch.cuda.synchronize()
        eq = lambda a, b: bool(torch.equal(a.view(torch.int64), b.view(torch.int64)))
        res = dict(all_gather=eq(full, one) and bool(torch.equal(ok, ok1)),
                   gather_root=eq(root, one) and bool(torch.equal(rok, ok1)),
                   peer_store=eq(pfull, one) and bool(torch.equal(pok, ok1)), peer_replicated=all(flags),
                   table=eq(table, whole), lookup=eq(lo, lo1) and bool(torch.equal(lk, lk1)),
                   solved=float(ok1.float().mean()))
        q.put(res)
    else:
        assert root is None and pfull is None
    dist.barrier()
    pg.close()
    pg2.close()
    T.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1_000_003, 7])
def test_sharded_cuda_paths_equal_single_gpu(n):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    res = None
    for _ in range(300):
        try:
            res = q.get(timeout=1)
            break
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res is not None
    assert res["all_gather"] and res["gather_root"] and res["peer_store"] and res["peer_replicated"] and res["table"] and res["lookup"], res
    assert res["solved"] > 0.9
