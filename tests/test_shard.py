"""N>1 host logic on CPU: world_size-2 gloo processes exercise the spin-range partition and the
final gather (the only collective of the path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_shard_bounds_cover_exactly():
    from multiband_rf_pulse_design_b200.shard import shard_bounds
    for n in (0, 1, 7, 1000, 10 ** 6, 10 ** 6 + 3):
        for w in (1, 2, 3, 4, 8):
            b = shard_bounds(n, w)
            assert len(b) == w and b[0][0] == 0
            assert sum(c for _, c in b) == n
            assert all(b[i][0] + b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in b) - min(c for _, c in b) <= 1
    with pytest.raises(ValueError):
        shard_bounds(5, 0)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiband_rf_pulse_design_b200.shard import gather_planes, shard_bounds
    bounds = shard_bounds(n, world)
    s0, cnt = bounds[rank]
    width = max(c for _, c in bounds)
    local = torch.full((3, width), float("nan"), dtype=torch.float64)
    s = torch.arange(s0, s0 + cnt, dtype=torch.float64)
    # stand-in for the kernel's output: a known function of the GLOBAL spin index per plane
    local[0, :cnt], local[1, :cnt], local[2, :cnt] = s, -2 * s, s * s
    full = gather_planes(local, [c for _, c in bounds], dst=0)
    if rank == 0:
        ref = torch.arange(n, dtype=torch.float64)
        ok = (full.shape == (3, n) and torch.equal(full[0], ref) and torch.equal(full[1], -2 * ref)
              and torch.equal(full[2], ref * ref))
        q.put(bool(ok))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def _worker_pieces(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiband_rf_pulse_design_b200.shard import pipelined_gather, shard_bounds
    bounds = shard_bounds(n, world)
    s0, cnt = bounds[rank]
    width = max(c for _, c in bounds)
    ok = True
    for chunks in (1, 3, [0.85, 0.15], [0.5, 0.3, 0.2]):
        local = torch.full((3, width), float("nan"), dtype=torch.float64)

        def fill(first, size, piece):       # stand-in for the kernel: a known function of the GLOBAL spin index
            s = torch.arange(s0 + first, s0 + first + size, dtype=torch.float64)
            piece[0, :size], piece[1, :size], piece[2, :size] = s, -2 * s, s * s
        full = pipelined_gather(fill, local, bounds, chunks)
        if rank == 0:
            ref = torch.arange(n, dtype=torch.float64)
            ok = ok and (full.shape == (3, n) and torch.equal(full[0], ref) and torch.equal(full[1], -2 * ref)
                         and torch.equal(full[2], ref * ref))
        else:
            assert full is None
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def _worker_sweep(rank, world, port, ndesigns, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiband_rf_pulse_design_b200.shard import gather_sweep
    n = 5
    mine = np.arange(rank, ndesigns, world)                 # fir_ap_cvx_sweep: instance i -> rank i mod world
    res = dict(index=mine, x=np.outer(mine + 1.0, np.arange(1, 2 * n)), ripple_stop=mine * 0.5,
               info=np.tile(mine[:, None].astype(float), (1, 8)))
    full = gather_sweep(res, n, ndesigns)
    if rank == 0:
        ids = np.arange(ndesigns)
        ok = (np.array_equal(full["index"], ids) and np.array_equal(full["x"], np.outer(ids + 1.0, np.arange(1, 2 * n)))
              and np.array_equal(full["ripple_stop"], ids * 0.5) and np.array_equal(full["info"][:, 3], ids.astype(float)))
        q.put(bool(ok))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def _worker_sweep_grouped(rank, world, port, q):
    """fir_ap_cvx_sweep itself (solve stubbed out: no GPU here) with grouped batches and host threads, then the real gather"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multiband_rf_pulse_design_b200 import fir
    from multiband_rf_pulse_design_b200.shard import gather_sweep

    def fake(n, f_list, a_list, d_list, obj_list, peak_list, ipm_max_iter=None, want_h=True, oversamp=15):
        B = len(f_list)
        x = np.zeros((B, 2 * n - 1))
        x[:, 0], x[:, 1], x[:, 2] = obj_list, peak_list, [f[0] for f in f_list]
        info = np.ones((B, 8))
        info[:, 7] = np.asarray(obj_list) * 2
        return x, None, info, (10, 2)
    fir._solve_batch_ap_device = fake
    n = 6
    f = np.array([-0.5, -0.2, 0.1, 0.4])
    objs, peaks, fadds = np.logspace(-2, 2, 5), np.logspace(-4, -2, 3), np.linspace(0, 0.02, 3)
    fl, ol, pl = fir.sweep_grid(f, objs, peaks, fadds)
    r = fir.fir_ap_cvx_sweep(n, f, [1, 1, 0, 0], [0.1, 0.1], objs, peaks, fadds, rank=rank, world=world, batch=8, concurrent_batches=3)
    full = gather_sweep(r, n, len(fl))
    if rank == 0:
        ok = np.array_equal(full["index"], np.arange(len(fl)))
        for i in range(len(fl)):
            ok = ok and full["x"][i, 0] == ol[i] and full["x"][i, 1] == pl[i] and full["x"][i, 2] == fl[i][0] and full["ripple_stop"][i] == 2 * ol[i]
        q.put(bool(ok))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_grouped_sweep_and_gather_world2_gloo():
    """batch composition (designs grouped by Peak / band edges, several batches in flight) must not disturb the partition the
    gather checks: every instance once, in instance order, each row with its own design's numbers"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_sweep_grouped, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


@pytest.mark.parametrize("ndesigns", [7, 16])
def test_sweep_gather_world2_gloo(ndesigns):
    """the final gather of a design sweep (x, ripple_stop, status rows of every instance to rank 0), ragged shares"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_sweep, args=(r, 2, port, ndesigns, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


def test_piece_widths():
    from multiband_rf_pulse_design_b200.shard import piece_widths
    for width in (1, 7, 1000, 10 ** 6):
        for chunks in (1, 2, 4, [0.85, 0.15], [1, 1, 1]):
            w = piece_widths(width, chunks)
            assert sum(w) == width and all(v >= 0 for v in w)
    with pytest.raises(ValueError):
        piece_widths(10, [0.5, -0.1])


@pytest.mark.parametrize("n", [1001, 4096])
def test_pipelined_gather_world2_gloo(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_pieces, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


@pytest.mark.parametrize("n", [1001, 4096])
def test_gather_world2_gloo(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


def test_sweep_partition_is_a_partition():
    """fir_ap_cvx_sweep gives design i to rank i mod world: every instance exactly once."""
    import numpy as np
    from multiband_rf_pulse_design_b200.fir import sweep_grid
    fl, ol, pl = sweep_grid([-0.5, -0.2, 0.2, 0.5], [0.1, 1.0, 10.0], [1e-3, 2e-3], [0.0, 0.01, 0.02, 0.03])
    assert len(fl) == len(ol) == len(pl) == 24
    assert fl[1][0] == -0.51 and fl[1][1] == -0.19          # f_add fastest, widens every band on both sides
    for world in (1, 2, 3, 8):
        seen = np.concatenate([np.arange(r, len(fl), world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(24))
