import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND
objs = np.logspace(-2, 1, 64); peaks = np.logspace(-3.2, -2, 8)
r = fir.fir_ap_cvx_sweep(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, [0.0], batch=512, max_iter=60000)
fl, ol, pl = fir.sweep_grid(H1_DUALBAND["f"], objs, peaks, [0.0])
it = r["info"][:, 1]
ol = np.array(ol); pl = np.array(pl)
print("iterations by obj (rows: every 4th obj; columns: the 8 peaks)")
for o in objs[::4]:
    sel = np.isclose(ol, o)
    print(f"obj {o:8.3f}: " + " ".join(f"{int(v):6d}" for v in it[sel][np.argsort(pl[sel])]))
