import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multiband_rf_pulse_design_b200 import fir
H1 = dict(f=[-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006], a=[0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886], d=[0.014436, 0.022361, 0.017683])
objs = np.logspace(-2, 1, 8); peaks = np.logspace(-3.2, -2, 8)
t = time.time()
r = fir.fir_ap_cvx_sweep(256, H1["f"], H1["a"], H1["d"], objs, peaks, [0.0], max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 60000)
print("seconds", time.time() - t)
info = r["info"].reshape(8, 8, 8)
np.set_printoptions(linewidth=200, precision=3, suppress=True)
print("status (rows obj, cols Peak)\n", info[:, :, 0])
print("iters/1000\n", info[:, :, 1] / 1000)
print("obj\n", info[:, :, 2] * 1000)
print("pr*1e6\n", info[:, :, 4] * 1e6)
print("dr*1e6\n", info[:, :, 5] * 1e6)
print("relgap*1e5\n", np.abs(info[:, :, 2] - info[:, :, 3]) / np.abs(info[:, :, 2]) * 1e5)
