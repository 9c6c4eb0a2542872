"""Per-rank solve time of the cfg4 sweep under torchrun (one process per GPU) — host contention check.
torchrun --nproc-per-node N tools/sweep_mp.py ; or SOLO=1 WORLD=N RANK=r python tools/sweep_mp.py on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from multiband_rf_pulse_design_b200._lib import check
from bench import H1_DUALBAND
world = int(os.environ.get("WORLD_SIZE", os.environ.get("WORLD", "1")))
rank = int(os.environ.get("RANK", "0"))
local = 0 if os.environ.get("SOLO") else int(os.environ.get("LOCAL_RANK", "0"))
check(m.lib().mbrf_set_device(local))
n, per_gpu = 256, 512
objs = np.logspace(-2, 1, max(1, per_gpu * world // 8))
peaks = np.logspace(-3.2, -2, 8)
fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs[:8], peaks, [0.0], max_iter=512)
for rep in range(int(os.environ.get("REPS", "1")) - 1):
    t0 = time.perf_counter()
    fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, [0.0], rank=rank, world=world,
                         batch=per_gpu, max_iter=60000, seed_stride=(lambda v: v if v == 'auto' else int(v))(os.environ.get('SEED', '0')))
    print(f"world {world} rank {rank} rep {rep}: {time.perf_counter() - t0:.2f} s", flush=True)
t0 = time.perf_counter()
r = fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, [0.0], rank=rank, world=world,
                         batch=per_gpu, max_iter=60000, seed_stride=(lambda v: v if v == 'auto' else int(v))(os.environ.get('SEED', '0')))
sec = time.perf_counter() - t0
info = r["info"]
print(f"world {world} rank {rank}: {sec:.2f} s, solved {(info[:,0]==1).sum()}/{info.shape[0]}, iters mean {info[:,1].mean():.0f} max {info[:,1].max():.0f}", flush=True)
