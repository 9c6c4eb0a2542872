import sys, warnings
import numpy as np
sys.path.insert(0, ".")
import bench
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
f, a, d, dt = bench.c13_bssfp_spec()
lib = m.lib()
for split in (-1, 0):
    lib.mbrf_ipm_set_option(5, float(split))
    for n in (50, 56, 57, 58, 59, 60, 64):
        hs, st, ex = fir.fir_ap_cvx_batch(n, [f], a, d, [0.1], [1e-3], return_info=True, method="ipm")
        i = ex["info"][0]
        print(f"split={split} n={n}: status {int(i[0])} iters {int(i[1])} obj {i[2]:.8f} dual {i[3]:.8f} viol {i[4]:.2e} dres {i[5]:.2e}")
