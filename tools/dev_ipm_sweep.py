#!/usr/bin/env python
"""time the cfg4 slice (obj x Peak x f_add grid of fir_ap_cvx designs at N = 256) on the interior-point solver"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import H1_DUALBAND as S
nobj, npk, nfa = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 8, 4)))
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 2
m.lib().mbrf_ipm_set_option(0, float(prec))
f = np.array(S["f"])
df_min = (f[2:-1:2] - f[1:-2:2]).min()
objs = np.logspace(-2, 4, nobj); peaks = np.logspace(-4, -2, npk); fadds = np.linspace(0, 0.9 * df_min / 2, nfa)
for rep in range(2):
    t0 = time.time()
    r = fir.fir_ap_cvx_sweep(256, f, S["a"], S["d"], objs, peaks, fadds, batch=512, method="ipm")
    dt = time.time() - t0
    info = r["info"]
    n = info.shape[0]
    st = info[:, 0]
    print(f"{n} designs in {dt:.2f}s = {n/dt:.1f} designs/s; status counts {dict(zip(*np.unique(st, return_counts=True)))}; "
          f"iters mean {info[:,1].mean():.1f} max {info[:,1].max():.0f}; max viol {info[st==1,4].max() if (st==1).any() else float('nan'):.1e}", flush=True)
