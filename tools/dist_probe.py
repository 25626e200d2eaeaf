"""Development probe (torchrun, N ranks): what the result reassembly of a sharded C4 batch costs, per mechanism.
weak = 1e7 pairs per GPU, strong = 1e7 pairs in total.  Times: CUDA events on every rank, max over ranks."""
import os, sys, json, numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from airiceraytracing_b200 import AirIceSolver, UNITS_CM_RAD
from airiceraytracing_b200.dist import PeerGather, shard_range, solve_sharded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
S = AirIceSolver(os.path.join(ROOT, "tests", "golden", "Atmosphere.dat"), device=rank)
dev = torch.device("cuda", rank)

def batch(n, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    h = 3001 + (100000 - 3001) * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    a = 90.2 + (179.8 - 90.2) * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    d = (h - 3000 + 200) * torch.tan((180 - a) * (3.1415927 / 180))
    return h * 100, d * 100

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(float(t))
    return min(ts), float(np.median(ts))

res = {"world": world}
for mode, n in (("weak", 10_000_000 * world), ("strong", 10_000_000)):
    h, d = batch(n, 20260418)
    b, e = shard_range(n, rank, world)
    out = torch.empty((9, e - b), dtype=torch.float64, device=dev); ok = torch.empty(e - b, dtype=torch.uint8, device=dev)
    fn = lambda hs, ds: S.solve(hs, ds, -20000.0, 300000.0, UNITS_CM_RAD, out=out, ok=ok)
    r = {}
    r["compute_only"] = timed(lambda: fn(h[b:e], d[b:e]))
    r["nccl_all_gather"] = timed(lambda: solve_sharded(fn, h, d))
    r["nccl_gather_root"] = timed(lambda: solve_sharded(fn, h, d, dst=0))
    pg = PeerGather(S, 9, n, dst=0)
    # sync=False + an NCCL barrier-sized all-reduce on the stream: device-side ordering only, no host sync inside the timing
    tok = torch.zeros(1, device=dev)
    def peer():
        pg.solve(S, h, d, -20000.0, 300000.0, UNITS_CM_RAD, sync=False); dist.all_reduce(tok)
    r["peer_store_root"] = timed(peer)
    pg.close()
    for ch in (1, 4, 8):
        pg2 = PeerGather(S, 9, n, dst=None, chunks=ch)
        def rep():
            pg2.solve(S, h, d, -20000.0, 300000.0, UNITS_CM_RAD, sync=False); dist.all_reduce(tok)
        r["peer_replicated_chunks%d" % ch] = timed(rep)
        pg2.close()
    res[mode] = {k: {"best_ms": v[0], "median_ms": v[1], "solves_per_s": n / v[0] * 1e3} for k, v in r.items()}
    res[mode]["pairs"] = n
    del h, d, out, ok
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(res, indent=1))
dist.barrier(); dist.destroy_process_group()
