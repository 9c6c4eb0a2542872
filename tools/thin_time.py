"""Per-iteration cost of the thin path (1, 2, 4, 8 designs) on the cfg4 LP: fixed iteration count, no convergence."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND
for B in (1, 2, 4, 8):
    designs = [fir.assemble_fir_ap(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 5.0 + 0.1 * b, 10 ** -2.5) for b in range(B)]
    fir._solve_batch_ap(256, designs, max_iter=256)
    t = time.perf_counter()
    fir._solve_batch_ap(256, designs, max_iter=6400, eps_pr=1e-30)
    dt = time.perf_counter() - t
    print(f"B={B}: {dt / 6400 * 1e6:.1f} us per iteration (incl. checks every 64)", flush=True)
