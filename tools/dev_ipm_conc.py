import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import H1_DUALBAND as S
f = np.array(S["f"]); df_min = (f[2:-1:2] - f[1:-2:2]).min()
objs = np.logspace(-2, 4, 16); peaks = np.logspace(-4, -2, 16); fadds = np.linspace(0, 0.9 * df_min / 2, 16)[:4]
fir.fir_ap_cvx_sweep(256, f, S["a"], S["d"], objs[:2], peaks[-2:], fadds[:2], batch=8, method="ipm")
for batch, conc in ((512, 1), (512, 2), (256, 2), (256, 4), (128, 4)):
    t0 = time.time()
    r = fir.fir_ap_cvx_sweep(256, f, S["a"], S["d"], objs, peaks, fadds, batch=batch, method="ipm", concurrent_batches=conc)
    dt = time.time() - t0
    st = r["info"][:, 0]
    print(f"batch {batch} concurrent {conc}: {len(st)} designs {dt:.2f}s = {len(st)/dt:.0f}/s  solved {int((st==1).sum())} infeasible {int((st==2).sum())} limit {int((st==3).sum())}", flush=True)
