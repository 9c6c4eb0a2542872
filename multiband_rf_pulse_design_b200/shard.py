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


def bloch_sharded(lib, dev_args: dict, nspins: int, out_local, workspace, stream, mode=0, gamma=6726.1,
                  group=None):
    """Run this rank's spin range through mbrf_bloch_device and gather mx,my,mz on rank 0.

    dev_args holds device pointers / sizes: b1r,b1i,gx,gy,gz,dt,ntime,t1,t2,df,nf,dx,dy,dz,npos.
    out_local is a [3, max shard] float64 device tensor.  Mode 0/1 only (one value per spin).
    """
    import torch.distributed as dist
    from ._lib import check
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = shard_bounds(nspins, world)
    s0, cnt = bounds[rank]
    a = dev_args
    check(lib.mbrf_bloch_device(a["b1r"], a["b1i"], a["gx"], a["gy"], a["gz"], a["dt"], a["ntime"], a["t1"], a["t2"],
                                a["df"], a["nf"], a["dx"], a["dy"], a["dz"], a["npos"], s0, cnt, None, None, None, 1,
                                out_local[0].data_ptr(), out_local[1].data_ptr(), out_local[2].data_ptr(), mode,
                                gamma, workspace, stream))
    if world == 1:
        return out_local[:, :cnt]
    return gather_planes(out_local, [c for _, c in bounds], dst=0, group=group)
