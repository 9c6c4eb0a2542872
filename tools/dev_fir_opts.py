import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
H1 = dict(f=[-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006], a=[0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886], d=[0.014436, 0.022361, 0.017683])
objs = np.logspace(-2, 1, 8); peaks = np.logspace(-3.2, -2, 8)
lib = m.lib()
base = [0.9, 0.2, 0.8, 0.36, 0.5]
trials = [("e99+a.2", {0: 0.99, 3: 0.2}), ("a.1", {3: 0.1}), ("a.15", {3: 0.15}), ("e99+a.15", {0: 0.99, 3: 0.15}),
          ("e99+a.1", {0: 0.99, 3: 0.1}), ("e99+a.05", {0: 0.99, 3: 0.05}), ("e99+a.2+s.3", {0: 0.99, 3: 0.2, 1: 0.3})]
ce = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for name, ch in trials:
    for i, v in enumerate(base):
        lib.mbrf_pdhg_set_option(i, ch.get(i, v))
    t = time.time()
    r = fir.fir_ap_cvx_sweep(256, H1["f"], H1["a"], H1["d"], objs, peaks, [0.0], max_iter=60000, check_every=ce)
    info = r["info"]
    print(f"{name:10s} solved {int((info[:,0]==1).sum())}/64 mean it {info[:,1].mean():8.0f} max {info[:,1].max():6.0f}  {time.time()-t:.2f}s", flush=True)
