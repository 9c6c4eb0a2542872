import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multiband_rf_pulse_design_b200 import fir
H1 = dict(f=[-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006], a=[0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886], d=[0.014436, 0.022361, 0.017683])
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
fir.fir_ap_cvx_batch(256, [H1["f"]], H1["a"], H1["d"], [0.1], [1e-3], max_iter=256)     # warm-up: context, buffers
for B in (1, 2, 4, 8, 64, 128, 512):
    objs = list(np.logspace(-2, 1, B)) if B > 1 else [0.1]
    ts = []
    for rep in range(2):
        t = time.time()
        hs, st, ex = fir.fir_ap_cvx_batch(256, [H1["f"]] * B, H1["a"], H1["d"], objs, [1e-3] * B, return_info=True, max_iter=iters,
                                          eps_pr=1e-30)     # never converge: fixed iteration count
        ts.append(time.time() - t)
    print(f"B={B:4d}: {min(ts):.3f} s for {iters} iterations -> {min(ts) / iters * 1e6:7.1f} us/iteration, {min(ts) / iters / B * 1e6:7.2f} us per design-iteration", flush=True)
