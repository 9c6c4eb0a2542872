import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from oracle.fir_problems import H1_DUALBAND as S
for os_, mi in ((16, 2000000),):
    t0 = time.time()
    h, st, ex = m.fir_qp_cvx(256, S["f"], S["a"], S["d"], 120, 1e6, return_info=True, oversamp=os_, max_iter=mi)
    print(f"oversamp {os_}: {st} iters {ex['info'][1]:.0f} {time.time()-t0:.1f}s obj {ex['info'][2]:.8g} dual {ex['info'][3]:.8g} viol {ex['info'][4]:.2e}", flush=True)
