"""Latency of ONE fir_ap_cvx design at N = 256 (what a MATLAB user's call sees) with the factorisation split over many CTAs
(default) and with the one-CTA-per-design kernel (mbrf_ipm_set_option(5, 0))."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import H1_DUALBAND as S
lib = m.lib()
for split in (-1, 0):
    lib.mbrf_ipm_set_option(5, float(split))
    for obj in (0.1, 1e4):
        fir.fir_ap_cvx_batch(256, [S["f"]], S["a"], S["d"], [obj], [1.0])
        t = time.perf_counter()
        hs, st, ex = fir.fir_ap_cvx_batch(256, [S["f"]], S["a"], S["d"], [obj], [1.0], return_info=True)
        dt = time.perf_counter() - t
        print(f"split_max={split} obj={obj:g}: {st[0]} in {dt*1e3:.0f} ms, {int(ex['info'][0,1])} iterations, objective {ex['info'][0,2]:.10f}")
