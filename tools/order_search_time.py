"""Time-to-answer of the arbitrary-phase order search (BASELINE config 5 shape: fir_ap(..., min_order=1), orders searched
downwards from n) on the dual-band H-1 spec."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multiband_rf_pulse_design_b200 as m
from bench import H1_DUALBAND
n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m.fir_ap_cvx(64, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 0.1, 1e-3, max_iter=256)
t = time.perf_counter()
h, st, n_op, f_op = m.fir_ap(n0, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 1e-3, 1, 0, 0, max_iter=60000)
print(f"fir_ap min_order from n={n0}: status {st}, minimal order {n_op}, {time.perf_counter() - t:.1f} s", flush=True)
