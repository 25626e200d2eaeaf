"""Multi-GPU plumbing: the hot path shards by index with no exchange during compute (every cell and every pair is
independent); the only collective is ONE gather that reassembles results in caller order.

One process per GPU (torchrun), torch.distributed for the gather: NCCL over NVLink/NVSwitch on the GPU box, gloo in
the CPU tests.  The compute callbacks run on the local shard only, so the same functions drive the CUDA solver
(`AirIceSolver.solve`, `.table_build`) and -- in tests -- the CPU oracle."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [begin, end) of item indices for `rank`; shards differ in size by at most one."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_columns(local, n_total, group=None, dst=None):
    """Reassemble a [ncols, n_local] shard into [ncols, n_total] in index order with a single collective.

    dst=None -> all ranks get the result (all_gather); dst=r -> only rank r does (gather), others return None.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    assert local.shape[-1] == sizes[rank], (local.shape, sizes, rank)
    width = max(sizes)
    ncols = local.shape[0]
    padded = local
    if local.shape[-1] != width:
        padded = torch.zeros((ncols, width), dtype=local.dtype, device=local.device)
        padded[:, :local.shape[-1]] = local
    padded = padded.contiguous()
    if dst is None:
        buf = torch.empty((world * ncols, width), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        buf = buf.view(world, ncols, width)
    else:
        bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, bufs, dst=dst, group=group)
        if rank != dst:
            return None
        buf = torch.stack(bufs)
    return torch.cat([buf[r, :, :sizes[r]] for r in range(world)], dim=1)


def solve_sharded(solve_fn, h, d, group=None, dst=None):
    """Index-shard a pair batch: every rank holds the full (h, d) description (or generates it), solves its own
    contiguous slice with `solve_fn(h_slice, d_slice) -> (out [ncols, m], ok [m])`, and one gather returns the
    results in caller order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = h.shape[0]
    b, e = shard_range(n, rank, world)
    out, ok = solve_fn(h[b:e], d[b:e])
    packed = torch.cat([out, ok.to(out.dtype).unsqueeze(0)], dim=0)  # flags ride along: still one collective
    full = gather_columns(packed, n, group, dst)
    if full is None:
        return None, None
    return full[:-1], full[-1].to(torch.uint8)


def pairs_sharded(fn, *arrays, group=None, dst=None):
    """The same for any per-pair entry point (in-ice solve, two-ray selection, table lookup): `arrays` are equally long
    1-D inputs, `fn(*slices) -> (out [ncols, m], flags [m] or [nflag, m])`; returns (out, flags) in caller order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = arrays[0].shape[0]
    b, e = shard_range(n, rank, world)
    out, flags = fn(*[a[b:e] for a in arrays])
    fl = flags if flags.dim() == 2 else flags.unsqueeze(0)
    packed = torch.cat([out, fl.to(out.dtype)], dim=0)
    full = gather_columns(packed, n, group, dst)
    if full is None:
        return None, None
    nf = fl.shape[0]
    got = full[-nf:].to(flags.dtype)
    return full[:-nf], (got if flags.dim() == 2 else got[0])


def table_sharded(build_fn, n_h, n_th, group=None, dst=None):
    """Partition table ROWS contiguously (angles stay whole, so stores stay coalesced and the layer count is uniform
    along a row); `build_fn(row_begin, row_end) -> [ncols, (row_end-row_begin)*n_th]`.  One gather reassembles the
    table in the reference's cell order ihei*n_th + iang."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    r0, r1 = shard_range(n_h, rank, world)
    local = build_fn(r0, r1)
    sizes = [s * n_th for s in shard_sizes(n_h, world)]
    width = max(sizes)
    ncols = local.shape[0]
    padded = torch.zeros((ncols, width), dtype=local.dtype, device=local.device)
    padded[:, :local.shape[1]] = local
    if dst is None:
        buf = torch.empty((world * ncols, width), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(buf, padded.contiguous(), group=group)
        buf = buf.view(world, ncols, width)
    else:
        bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded.contiguous(), bufs, dst=dst, group=group)
        if rank != dst:
            return None
        buf = torch.stack(bufs)
    return torch.cat([buf[r, :, :sizes[r]] for r in range(world)], dim=1)
