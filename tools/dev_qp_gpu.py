import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import multiband_rf_pulse_design_b200 as m
from oracle.fir_problems import build_fir_qp, objective_fir_qp, violation_fir_qp
spec = dict(f=[-0.6, -0.25, 0.1, 0.45], a=[1, 1, 0.5, 0.5], d=[0.05, 0.05])
for n, k, obj in ((16, 2.0, 1.0), (16, 2.0, 0.0), (64, 8.0, 1.0)):
    t = time.time()
    h, st, ex = m.fir_qp_cvx(n, spec["f"], spec["a"], spec["d"], k, obj, return_info=True)
    p = build_fir_qp(n, spec["f"], spec["a"], spec["d"], k, obj)
    x = ex["x"]
    print(n, k, obj, st, "info", np.array2string(ex["info"], precision=6), "obj(cpu) %.7f viol(cpu) %.2e" % (objective_fir_qp(p, x), violation_fir_qp(p, x)), "%.2fs" % (time.time() - t))
H1 = dict(f=[-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006], a=[0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886], d=[0.014436, 0.022361, 0.017683])
t = time.time()
h, st, ex = m.fir_qp_cvx(256, H1["f"], H1["a"], H1["d"], 120, 1.0, return_info=True, max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 100000)
p = build_fir_qp(256, H1["f"], H1["a"], H1["d"], 120, 1.0)
print("cfg3 n=256", st, np.array2string(ex["info"], precision=6), "obj(cpu) %.7f viol(cpu) %.2e" % (objective_fir_qp(p, ex["x"]), violation_fir_qp(p, ex["x"])), "%.2fs" % (time.time() - t))
