"""Trace of one slow design of the cfg4 sweep (world 8, rank 0, local index 489 by default) solved alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND
world, rank, loc = 8, 0, int(sys.argv[1]) if len(sys.argv) > 1 else 489
objs = np.logspace(-2, 1, 512); peaks = np.logspace(-3.2, -2, 8)
fl, ol, pl = fir.sweep_grid(H1_DUALBAND["f"], objs, peaks, [0.0])
i = np.arange(rank, len(fl), world)[loc]
print("design", i, "obj", ol[i], "peak", pl[i], flush=True)
d = fir.assemble_fir_ap(256, fl[i], H1_DUALBAND["a"], H1_DUALBAND["d"], ol[i], pl[i])
x, t, info = fir._solve_batch_ap(256, [d], max_iter=int(os.environ.get("MAXIT", "100000")))
print("info", info[0])
