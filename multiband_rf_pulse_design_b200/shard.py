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


def gather_planes(local, counts, dst=0, group=None):
    """Gather per-rank result planes to `dst`.

    local  : tensor [planes, max(counts)] on this rank (only the first counts[rank] columns are valid)
    counts : per-rank column counts
    returns tensor [planes, sum(counts)] on rank `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    width = max(counts)
    if local.shape[1] != width:
        raise ValueError(f"local must be padded to the widest shard ({width} columns)")
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:, :c] for b, c in zip(bufs, counts)], dim=1)


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
                  group=None, chunks=1):
    """Run this rank's spin range through mbrf_bloch_device and gather mx,my,mz on rank 0.

    dev_args holds device pointers / sizes: b1r,b1i,gx,gy,gz,dt,ntime,t1,t2,df,nf,dx,dy,dz,npos.
    out_local is a [3, max shard] float64 device tensor.  Mode 0/1 only (one value per spin).
    stream: raw CUDA stream pointer, or a torch.cuda.Stream (needed for chunks > 1).
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

    if world == 1 or chunks == 1 or not hasattr(stream, "cuda_stream"):
        simulate(s0, cnt, out_local[0].data_ptr(), out_local[1].data_ptr(), out_local[2].data_ptr())
        if world == 1:
            return out_local[:, :cnt]
        return gather_planes(out_local, [c for _, c in bounds], dst=0, group=group)

    return pipelined_gather(lambda first, size, piece: simulate(s0 + first, size, piece[0].data_ptr(), piece[1].data_ptr(),
                                                                piece[2].data_ptr()),
                            out_local, bounds, chunks, stream=stream, group=group)


def pipelined_gather(fill_piece, out_local, bounds, chunks, stream=None, group=None, dst=0):
    """Produce this rank's shard piece by piece (fill_piece(first, size, piece[planes, width]) fills columns
    [first, first + size) of the shard into `piece`) and gather piece k while piece k + 1 is produced.
    On CUDA tensors the gathers run on a side stream behind an event of `stream` (a torch.cuda.Stream); on CPU tensors
    (gloo, tests) the same sequence runs synchronously.  Returns [planes, sum(counts)] on rank `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    planes = out_local.shape[0]
    cnt = bounds[rank][1]
    width = max(c for _, c in bounds)
    widths = piece_widths(width, chunks)
    starts = [sum(widths[:k]) for k in range(len(widths))]
    if out_local.numel() < planes * width:
        raise ValueError("out_local too small: [planes, widest shard] needed")
    flat = out_local.reshape(-1)
    pieces = [flat[planes * o:planes * (o + w)].view(planes, w) for o, w in zip(starts, widths)]   # piece-major staging
    on_gpu = out_local.is_cuda
    comm = _comm_stream(out_local.device) if on_gpu else None
    recv, works = [], []
    for k, (o, w) in enumerate(zip(starts, widths)):
        size = max(0, min(w, cnt - o))
        if size:
            fill_piece(o, size, pieces[k])
        bufs = None
        if on_gpu:
            ev = torch.cuda.Event()
            ev.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                bufs = [torch.empty_like(pieces[k]) for _ in range(world)] if rank == dst else None
                works.append(dist.gather(pieces[k], bufs, dst=dst, group=group, async_op=True))
        else:
            bufs = [torch.empty_like(pieces[k]) for _ in range(world)] if rank == dst else None
            dist.gather(pieces[k], bufs, dst=dst, group=group)
        recv.append(bufs)

    def assemble():
        # (receiving every plane and piece straight into its place -- one gather per plane and piece, no concatenation --
        # was measured at 2 GPUs and is 0.03 ms slower: six collectives instead of two)
        if rank != dst:
            return None
        cols = []
        for r, (_, c) in enumerate(bounds):
            for k, (o, w) in enumerate(zip(starts, widths)):
                size = max(0, min(w, c - o))
                if size:
                    cols.append(recv[k][r][:, :size])
        return torch.cat(cols, dim=1)

    if not on_gpu:
        return assemble()
    with torch.cuda.stream(stream):
        for wk in works:
            wk.wait()
        return assemble()


_COMM = {}


def _comm_stream(device):
    import torch
    key = str(device)
    if key not in _COMM:
        _COMM[key] = torch.cuda.Stream(device=device)
    return _COMM[key]
