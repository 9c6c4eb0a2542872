#!/usr/bin/env python
"""cfg3 trial: fir_qp_cvx(256, f, a, d, 120, obj) on the first-order solver for obj = 1 .. 1e6"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from oracle.fir_problems import H1_DUALBAND as S, build_fir_qp, objective_fir_qp, violation_fir_qp
for obj in [float(v) for v in sys.argv[1:]] or [1.0, 1e3, 1e6]:
    t0 = time.time()
    h, st, ex = m.fir_qp_cvx(256, S["f"], S["a"], S["d"], 120, obj, return_info=True, max_iter=300000)
    dt = time.time() - t0
    p = build_fir_qp(256, S["f"], S["a"], S["d"], 120, obj)
    x = ex["x"]
    print(f"obj {obj:g}: {st} iters {ex['info'][1]:.0f} {dt:.2f}s objective {ex['info'][2]:.8g} dual {ex['info'][3]:.8g} "
          f"recomputed {objective_fir_qp(p, x):.8g} viol {violation_fir_qp(p, x):.2e} E {np.linalg.norm(x):.6g} Peak {np.hypot(x[:256], x[256:]).max():.6g}", flush=True)
