"""Batch-width history of one rank's share of the cfg4 sweep: python tools/sweep_trace.py [world] [rank]
(MBRF_PDHG_TRACE=1 prints one line per convergence check: iteration, live designs, batch width)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n, per_gpu = 256, 512
total = per_gpu * world
objs = np.logspace(-2, 1, max(1, total // 8))
peaks = np.logspace(-3.2, -2, 8)
fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs[:8], peaks, [0.0], max_iter=512)
if os.environ.get("PYPROF"):
    import cProfile, pstats
    cProfile.run('fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, [0.0], rank=rank, world=world, batch=per_gpu, max_iter=2000)', "/tmp/prof")
    pstats.Stats("/tmp/prof").sort_stats("cumtime").print_stats(14)
t0 = time.perf_counter()
r = fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, [0.0], rank=rank, world=world,
                         batch=per_gpu, max_iter=60000)
sec = time.perf_counter() - t0
info = r["info"]
print(f"world {world} rank {rank}: {sec:.2f} s, solved {(info[:,0]==1).sum()}/{info.shape[0]}, iters mean {info[:,1].mean():.0f} max {info[:,1].max():.0f}")
bad = np.nonzero(info[:, 0] != 1)[0]
for b in bad:
    print("unsolved local", b, "status", info[b, 0], "obj", info[b, 2], "dual", info[b, 3], "pr", info[b, 4], "dr", info[b, 5])
