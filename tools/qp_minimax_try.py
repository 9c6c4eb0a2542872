import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multiband_rf_pulse_design_b200 as m
from oracle.fir_problems import build_fir_qp, objective_fir_qp_minimax, violation_fir_qp_minimax
SPEC = dict(f=[-0.6, -0.25, 0.1, 0.45], a=[1, 1, 0.5, 0.5], d=[0.05, 0.05])
for n, k, obj2 in [(16, 2.0, [0.5, 1.0]), (16, 2.0, [0.1, 0.0]), (12, 1.5, [1.0, 2.0])]:
    h, st, ex = m.fir_qp_cvx(n, SPEC["f"], SPEC["a"], SPEC["d"], k, obj2, return_info=True)
    p = build_fir_qp(n, SPEC["f"], SPEC["a"], SPEC["d"], k, 0.0)
    print(n, k, obj2, st, "iters", ex["info"][1], "solver obj", ex["info"][2], "dual", ex["info"][3], "recomputed", objective_fir_qp_minimax(p, ex["x"], obj2),
          "viol", violation_fir_qp_minimax(p, ex["x"]), flush=True)
