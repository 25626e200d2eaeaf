"""Multi-GPU plumbing: the hot path shards by index with no exchange during compute (every cell and every pair is
independent); the only collective is ONE gather that reassembles results in caller order.

One process per GPU (torchrun), torch.distributed for the gather: NCCL over NVLink/NVSwitch on the GPU box, gloo in
the CPU tests.  The compute callbacks run on the local shard only, so the same functions drive the CUDA solver
(`AirIceSolver.solve`, `.table_build`) and -- in tests -- the CPU oracle.

`PeerGather` is the B200 form of that one gather: the consumer's result block is mapped into every producer's address
space (CUDA IPC over NVLink peer access) and the producers' solve kernels get their output-column pointers INSIDE it, so
the stores of a shard travel to their final place while the kernel computes -- no staging buffer, no collective launch, no
concatenation.  The result rate of one B200 (6.8e9 solves/s x 73 B = 0.5 TB/s) is of the order of a GPU's NVLink
ingress (0.9 TB/s), so for a single consumer the link, not the kernels, bounds the job beyond two producers; the peer
stores make the job run AT that bound."""
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


class PeerGather:
    """Result block [ncols f64 columns + one uint8 flag row] x n_total in ONE place, written by every rank's kernel.

    dst = r: the block lives on rank r; every other rank maps it and stores its shard into it over NVLink.
    dst = None: every rank holds a block; a rank computes its shard into its own block in `chunks` pieces and copies each
    finished piece to the other ranks' blocks on a side stream while the next piece computes."""

    def __init__(self, solver, ncols, n_total, dst=0, group=None, chunks=4):
        self.solver, self.ncols, self.n, self.dst, self.group, self.chunks = solver, int(ncols), int(n_total), dst, group, int(chunks)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.col_bytes = self.n * 8
        self.flag_off = self.ncols * self.col_bytes
        nbytes = self.flag_off + self.n + 256
        owner = dst is None or self.rank == dst
        self.own, handle = solver.peer_alloc(nbytes) if owner else (None, None)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.maps = {}                       # rank -> base pointer of that rank's block as seen from here
        for r, hd in enumerate(handles):
            if hd is None:
                continue
            self.maps[r] = self.own if r == self.rank else solver.peer_open(hd)
        self.out = self.ok = None
        if owner:
            self.out = solver.wrap_device_memory(self.own, (self.ncols, self.n), torch.float64)
            self.ok = solver.wrap_device_memory(self.own + self.flag_off, (self.n,), torch.uint8)
        self.side = torch.cuda.Stream(device=solver.torch_device) if dst is None else None

    def _ptrs(self, base, begin):
        return [base + k * self.col_bytes + begin * 8 for k in range(self.ncols)], base + self.flag_off + begin

    def solve(self, solver, h, d, depth, ice, units, sync=True):
        """Every rank holds the full (h, d); solves its contiguous slice; returns (out [ncols, n], ok [n]) where the block
        lives (None, None elsewhere).  sync=False leaves the cross-rank ordering to the caller (bench timing)."""
        b, e = shard_range(self.n, self.rank, self.world)
        if self.dst is not None:
            outp, okp = self._ptrs(self.maps[self.dst], b)
            if e > b:
                solver.solve_into(h[b:e], d[b:e], depth, ice, units, outp, okp)
        else:
            cur = torch.cuda.current_stream(solver.torch_device)
            step = max(1, -(-(e - b) // self.chunks))
            for c0 in range(b, e, step):
                c1 = min(e, c0 + step)
                outp, okp = self._ptrs(self.own, c0)
                solver.solve_into(h[c0:c1], d[c0:c1], depth, ice, units, outp, okp)
                ev = torch.cuda.Event()
                ev.record(cur)
                self.side.wait_event(ev)
                for r, base in self.maps.items():
                    if r == self.rank:
                        continue
                    po, pk = self._ptrs(base, c0)
                    for k in range(self.ncols):
                        solver.peer_copy(po[k], outp[k], (c1 - c0) * 8, self.side.cuda_stream)
                    solver.peer_copy(pk, okp, c1 - c0, self.side.cuda_stream)
            cur.wait_stream(self.side)
        if sync:
            dist.barrier(group=self.group)     # every producer's stores have landed before any consumer reads
        return (self.out, self.ok) if self.out is not None else (None, None)

    def close(self):
        for r, base in self.maps.items():
            if r != self.rank:
                self.solver.peer_close(base)
        self.maps = {}
        self.out = self.ok = None
        if self.own is not None:
            self.solver.peer_free(self.own)
            self.own = None
