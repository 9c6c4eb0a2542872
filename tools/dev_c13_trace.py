import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
f, a, d, dt = bench.c13_bssfp_spec()
lib = m.lib()
n = int(sys.argv[1]); split = float(sys.argv[2])
lib.mbrf_ipm_set_option(5, split)
lib.mbrf_ipm_set_option(3, 1.0)
hs, st, ex = fir.fir_ap_cvx_batch(n, [f], a, d, [0.1], [1e-3], return_info=True, method="ipm")
print(ex["info"][0])
