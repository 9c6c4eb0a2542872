"""Multi-GPU sharding of the Bloch / SLR paths: one process per GPU, contiguous ranges of the
flattened spin index s = p + npos*f (blochC.c:468-473) or position index ix + iy*nx (abrx.c:73),
no collective during compute, ONE gather of the result planes to the calling rank at the end.

torch.distributed is plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Balanced contiguous ranges: [(start, count)] * world, counts differ by at most one."""
    if world < 1 or n < 0:
        raise ValueError("world >= 1 and n >= 0 required")
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        c = base + (1 if r < extra else 0)
        out.append((start, c))
        start += c
    return out


_FULL = {}


def _full_buffer(planes, total, like):
    """The gathered result [planes, total] on the destination rank: allocated ONCE per shape and device and reused by every
    later call (the caller consumes it before the next gather), so that no allocation sits inside a timed step."""
    import torch
    key = (planes, total, like.dtype, str(like.device))
    buf = _FULL.get(key)
    if buf is None:
        if len(_FULL) > 8:
            _FULL.clear()
        buf = _FULL[key] = torch.empty((planes, total), dtype=like.dtype, device=like.device)
    return buf


def _peer(group, r):
    import torch.distributed as dist
    return r if group is None else dist.get_global_rank(group, r)


def _exchange(ops):
    """One batched point-to-point call (a single NCCL group launch on the GPU box)."""
    import torch.distributed as dist
    return dist.batch_isend_irecv(ops) if ops else []


def gather_planes(local, counts, dst=0, group=None, full=None):
    """Gather per-rank result planes to `dst`.

    local  : tensor [planes, >= counts[rank]] on this rank (only the first counts[rank] columns are valid)
    counts : per-rank column counts
    returns tensor [planes, sum(counts)] on rank `dst` (the reused buffer of _full_buffer unless `full` is given), None
    elsewhere.  Every rank sends exactly its valid columns and `dst` receives them straight into their place of the final
    layout: no padding, no list of receive buffers, no concatenation.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    planes = local.shape[0]
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(int)
    if local.shape[1] < counts[rank]:
        raise ValueError(f"local has {local.shape[1]} columns, this rank's shard has {counts[rank]}")
    ops = []
    if rank == dst:
        if full is None:
            full = _full_buffer(planes, int(starts[-1]), local)
        full[:, starts[rank]:starts[rank] + counts[rank]] = local[:, :counts[rank]]
        for r in range(world):
            if r != dst and counts[r]:
                ops += [dist.P2POp(dist.irecv, full[p, starts[r]:starts[r] + counts[r]], _peer(group, r), group) for p in range(planes)]
    elif counts[rank]:
        ops = [dist.P2POp(dist.isend, local[p, :counts[rank]], _peer(group, dst), group) for p in range(planes)]
    for w in _exchange(ops):
        w.wait()
    return full if rank == dst else None


class PeerResult:
    """The final [planes, total] float64 result of a sharded run, living on rank `dst`'s GPU and mapped into every other
    rank's address space (CUDA IPC, include/mbrf.h mbrf_peer_*).  Every rank's kernel stores its slice straight into it over
    NVLink; nothing is transferred afterwards.  Collective constructor (all ranks of `group`); `exchange` is a callable
    that all-gathers a picklable object (default: torch.distributed.all_gather_object on `group`)."""

    def __init__(self, lib, planes, total, group=None, dst=0, exchange=None):
        import ctypes as C
        import torch.distributed as dist
        from ._lib import check
        self.lib, self.planes, self.total, self.dst = lib, planes, total, dst
        self.rank = dist.get_rank(group)
        world = dist.get_world_size(group)
        self.owner = self.rank == dst
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        if self.owner:
            check(lib.mbrf_peer_alloc(planes * total * 8, C.byref(ptr), handle))
        if exchange is None:
            def exchange(o):
                outl = [None] * world
                dist.all_gather_object(outl, o, group=group)
                return outl
        handles = exchange(handle.raw if self.owner else None)
        if not self.owner:
            check(lib.mbrf_peer_open(handles[dst], C.byref(ptr)))
        self.base = ptr.value

    def plane_ptr(self, p, col0=0):
        return self.base + (p * self.total + col0) * 8

    def tensor(self):
        """[planes, total] torch view of the buffer (owner rank only)."""
        import torch
        if not self.owner:
            return None
        holder = type("_Buf", (), {})()
        holder.__cuda_array_interface__ = {"shape": (self.planes, self.total), "typestr": "<f8", "data": (self.base, False),
                                           "version": 2, "strides": None}
        self._holder = holder
        return torch.as_tensor(holder, device="cuda")

    def close(self):
        if self.base:
            (self.lib.mbrf_peer_free if self.owner else self.lib.mbrf_peer_close)(self.base)
            self.base = 0


def piece_widths(width: int, chunks) -> list[int]:
    """Padded widths of the pieces a shard of `width` columns is cut into for the pipelined gather: `chunks` is a number
    of equal pieces or a list of fractions (e.g. [0.85, 0.15]: a large piece whose transfer hides behind the simulation of
    the rest, and a small last piece whose transfer is the only exposed one).  The same on every rank."""
    if isinstance(chunks, int):
        chunks = [1.0] * max(1, chunks)
    fr = [float(f) for f in chunks]
    if not fr or min(fr) <= 0:
        raise ValueError("chunk fractions must be positive")
    tot = sum(fr)
    out, used = [], 0
    for f in fr[:-1]:
        w = int(width * f / tot)
        out.append(w)
        used += w
    out.append(width - used)
    return out


def bloch_sharded(lib, dev_args: dict, nspins: int, out_local, workspace, stream, mode=0, gamma=6726.1,
                  group=None, chunks=1, full=None, peer=None):
    """Run this rank's spin range through mbrf_bloch_device and gather mx,my,mz on rank 0.

    dev_args holds device pointers / sizes: b1r,b1i,gx,gy,gz,dt,ntime,t1,t2,df,nf,dx,dy,dz,npos.
    out_local is a [3, max shard] float64 device tensor.  Mode 0/1 only (one value per spin).
    stream: a torch.cuda.Stream (a raw CUDA stream pointer is accepted on one GPU only: the gather must be ordered
    behind the kernel, which needs the stream object).
    full: optional preallocated [3, nspins] result on rank 0 (default: one reused buffer per shape).
    peer: a PeerResult([3, nspins]) -- the ranks' kernels then store straight into rank 0's buffer (no gather step at all);
    returns None, the result is peer.tensor() on rank 0 once the caller has ordered completion (barrier / all-reduce).
    chunks > 1: the shard is simulated in that many pieces and piece k is gathered (NCCL, on a side stream) while piece
    k + 1 is being simulated, so that only the last piece's transfer is exposed: rank 0 receives (world - 1) x 24 bytes
    per spin over its NVLink ingress, 0.19 ms for 8 x 10^6 spins against 1.1 ms of simulation.
    """
    import torch.distributed as dist
    from ._lib import check
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = shard_bounds(nspins, world)
    s0, cnt = bounds[rank]
    a = dev_args
    raw = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream

    def simulate(first, count, o0, o1, o2):
        check(lib.mbrf_bloch_device(a["b1r"], a["b1i"], a["gx"], a["gy"], a["gz"], a["dt"], a["ntime"], a["t1"], a["t2"],
                                    a["df"], a["nf"], a["dx"], a["dy"], a["dz"], a["npos"], first, count, None, None, None, 1,
                                    o0, o1, o2, mode, gamma, workspace, raw))

    if world == 1:
        simulate(s0, cnt, out_local[0].data_ptr(), out_local[1].data_ptr(), out_local[2].data_ptr())
        return out_local[:, :cnt]
    if peer is not None:
        # one kernel per rank, storing straight into rank 0's result over NVLink (PeerResult); the caller orders completion
        simulate(s0, cnt, peer.plane_ptr(0, s0), peer.plane_ptr(1, s0), peer.plane_ptr(2, s0))
        return None
    if not hasattr(stream, "cuda_stream"):
        raise TypeError("bloch_sharded on several GPUs needs a torch.cuda.Stream: the gather is ordered behind the kernel by an event")

    return pipelined_gather(lambda first, size, piece: simulate(s0 + first, size, piece[0].data_ptr(), piece[1].data_ptr(),
                                                                piece[2].data_ptr()),
                            out_local, bounds, chunks, stream=stream, group=group, full=full)


def abr_sharded(lib, dev_args: dict, npos: int, out_local, workspace, stream, convention=0, group=None, chunks=1, full=None,
                peer=None):
    """Forward SLR over this rank's contiguous range of the position index ix + iy*nx (abrx.c:67-78) through
    mbrf_abr_device, alpha/beta planes (re, im, re, im) gathered on rank 0 exactly like the Bloch planes.
    dev_args: device pointers / sizes rfr, rfi, gx, gy, ns, x, nx, y, ny.  out_local: [4, widest shard] float64."""
    import torch.distributed as dist
    from ._lib import check
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = shard_bounds(npos, world)
    p0, cnt = bounds[rank]
    a = dev_args
    raw = stream.cuda_stream if hasattr(stream, "cuda_stream") else stream

    def run(first, count, piece):
        check(lib.mbrf_abr_device(a["rfr"], a["rfi"], a["gx"], a["gy"], a["ns"], a["x"], a["nx"], a["y"], a["ny"], convention,
                                  first, count, piece[0].data_ptr(), piece[1].data_ptr(), piece[2].data_ptr(),
                                  piece[3].data_ptr(), workspace, raw))

    if world == 1:
        run(p0, cnt, out_local)
        return out_local[:, :cnt]
    if peer is not None:
        check(lib.mbrf_abr_device(a["rfr"], a["rfi"], a["gx"], a["gy"], a["ns"], a["x"], a["nx"], a["y"], a["ny"], convention,
                                  p0, cnt, peer.plane_ptr(0, p0), peer.plane_ptr(1, p0), peer.plane_ptr(2, p0),
                                  peer.plane_ptr(3, p0), workspace, raw))
        return None
    if not hasattr(stream, "cuda_stream"):
        raise TypeError("abr_sharded on several GPUs needs a torch.cuda.Stream")
    return pipelined_gather(lambda first, size, piece: run(p0 + first, size, piece), out_local, bounds, chunks,
                            stream=stream, group=group, full=full)


def pipelined_gather(fill_piece, out_local, bounds, chunks, stream=None, group=None, dst=0, full=None):
    """Produce this rank's shard piece by piece (fill_piece(first, size, piece) fills columns [first, first + size) of the
    shard into the [planes, >= size] tensor view `piece`) and move piece k to rank `dst` while piece k + 1 is produced.

    Rank `dst` produces its own pieces IN PLACE in the final [planes, total] buffer and posts, per piece, one batched
    receive that lands every other rank's columns in their final place; the other ranks stage a piece in `out_local` and
    send its valid columns.  No receive lists, no concatenation, no allocation per call (`full`, or the reused buffer of
    _full_buffer).  On CUDA tensors the transfers run on a side stream behind an event of `stream` (a torch.cuda.Stream);
    on CPU tensors (gloo, tests) the same sequence runs synchronously.  Returns [planes, total] on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    planes = out_local.shape[0]
    s0, cnt = bounds[rank]
    total = sum(c for _, c in bounds)
    width = max(c for _, c in bounds)
    widths = piece_widths(width, chunks)
    starts = [sum(widths[:k]) for k in range(len(widths))]
    on_gpu = out_local.is_cuda
    if rank == dst:
        if full is None:
            full = _full_buffer(planes, total, out_local)
        elif tuple(full.shape) != (planes, total):
            raise ValueError(f"full must be [{planes}, {total}]")
    else:
        if out_local.numel() < planes * width:
            raise ValueError("out_local too small: [planes, widest shard] needed")
        flat = out_local.reshape(-1)
    comm = _comm_stream(out_local.device) if on_gpu else None
    works = []
    for k, (o, w) in enumerate(zip(starts, widths)):
        size = max(0, min(w, cnt - o))
        ops = []
        if rank == dst:
            if size:
                fill_piece(o, size, full[:, s0 + o:s0 + o + size])
            for r, (b0, c) in enumerate(bounds):
                sz = max(0, min(w, c - o))
                if r != dst and sz:
                    ops += [dist.P2POp(dist.irecv, full[p, b0 + o:b0 + o + sz], _peer(group, r), group) for p in range(planes)]
        elif size:
            piece = flat[planes * o:planes * (o + w)].view(planes, w)          # piece-major staging, rows contiguous
            fill_piece(o, size, piece)
            ops = [dist.P2POp(dist.isend, piece[p, :size], _peer(group, dst), group) for p in range(planes)]
        if on_gpu:
            ev = torch.cuda.Event()
            ev.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                works += _exchange(ops)
        else:
            for wk in _exchange(ops):
                wk.wait()
    if on_gpu:
        with torch.cuda.stream(stream):
            for wk in works:
                wk.wait()
    return full if rank == dst else None


def gather_sweep(res, n, ndesigns, group=None, dst=0, device=None):
    """Final gather of a design sweep (x, objective, status of every instance to the calling rank).
    res: this rank's dict(index, x [k, 2n-1], ripple_stop [k], info [k, 8]) from fir.fir_ap_cvx_sweep.
    Returns the same dict for ALL `ndesigns` instances, ordered by instance index, on rank `dst`; None elsewhere.
    One collective: every rank contributes a [ceil(ndesigns / world), 2n - 1 + 1 + 8 + 1] block (NCCL on `device`, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(res)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    rows, width = -(-ndesigns // world), 2 * n - 1 + 10
    blk = np.full((rows, width), -1.0)
    k = len(res["index"])
    blk[:k, 0] = res["index"]
    blk[:k, 1:2 * n] = res["x"]
    blk[:k, 2 * n] = res["ripple_stop"]
    blk[:k, 2 * n + 1:2 * n + 9] = res["info"]
    t = torch.from_numpy(blk)
    if device is not None:
        t = t.to(device)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=_peer(group, dst), group=group)
    if rank != dst:
        return None
    allb = torch.cat(bufs).cpu().numpy()
    allb = allb[allb[:, 0] >= 0]
    allb = allb[np.argsort(allb[:, 0])]
    if allb.shape[0] != ndesigns or not np.array_equal(allb[:, 0], np.arange(ndesigns)):
        raise RuntimeError("sweep gather: the ranks' instances do not partition the sweep")
    return dict(index=allb[:, 0].astype(int), x=allb[:, 1:2 * n], ripple_stop=allb[:, 2 * n], info=allb[:, 2 * n + 1:2 * n + 9])


_COMM = {}


def _comm_stream(device):
    import torch
    key = str(device)
    if key not in _COMM:
        _COMM[key] = torch.cuda.Stream(device=device)
    return _COMM[key]
