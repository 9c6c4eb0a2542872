import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND
lib = m.lib()
for mode in (2, 1, 2):
    lib.mbrf_pdhg_set_gemm(mode)
    t = time.perf_counter()
    _, st, ex = fir.fir_qp_cvx(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 120, 1.0, return_info=True)
    print(mode, st, time.perf_counter() - t, ex["info"][1], flush=True)
