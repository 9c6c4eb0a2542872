import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import build_fir_ap, violation_fir_ap
import bench
f, a, d, dt = bench.c13_bssfp_spec()
for n in [int(v) for v in sys.argv[1].split(",")]:
    hs, st, ex = fir.fir_ap_cvx_batch(n, [f], a, d, [0.1], [1e-3], return_info=True, method="ipm", ipm_max_iter=300)
    i = ex["info"][0]
    viol = violation_fir_ap(build_fir_ap(n, f, a, d, 0.1, 1e-3), np.concatenate([ex["x"][0], [ex["ripple_stop"][0]]])) if i[0] == 1 else float("nan")
    hs2, st2, ex2 = fir.fir_ap_cvx_batch(n, [f], a, d, [0.1], [1e-3], return_info=True, method="pdhg", max_iter=200000)
    j = ex2["info"][0]
    print(f"n={n}: ipm {st[0]} code {int(i[0])} it {int(i[1])} obj {i[2]:.8g} viol(cpu) {viol:.2e} | pdhg {st2[0]} code {int(j[0])} it {int(j[1])} obj {j[2]:.8g} viol {j[4]:.1e}", flush=True)
