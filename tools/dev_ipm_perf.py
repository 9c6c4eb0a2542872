#!/usr/bin/env python
"""IPM timing / iteration study: the cfg4 slice under different switch thresholds, plus the golden weights as a check."""
import os, sys, time, json, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import H1_DUALBAND as S
lib = m.lib()
W = json.load(open("tests/golden/fir_ap_weights_known.json"))
cases = [k for k, v in W.items() if v["status"] == 0]
f = np.array(S["f"]); df_min = (f[2:-1:2] - f[1:-2:2]).min()
objs = np.logspace(-2, 4, 16); peaks = np.logspace(-4, -2, 8); fadds = np.linspace(0, 0.9 * df_min / 2, 4)
for sw in [float(v) for v in sys.argv[1:]] or [1e-3]:
    lib.mbrf_ipm_set_option(1, sw)
    hs, st, ex = fir.fir_ap_cvx_batch(256, [W[c]["f"] for c in cases], S["a"], S["d"], [W[c]["obj"] for c in cases], [1.0] * len(cases), return_info=True, method="ipm")
    rel = [abs(ex["info"][i, 2] - W[c]["cone_free_obj"]) / W[c]["cone_free_obj"] for i, c in enumerate(cases)]
    print(f"switch {sw:g}: goldens max rel {max(rel):.1e} status {set(st)} iters {ex['info'][:,1].astype(int).tolist()}")
    for rep in range(2):
        t0 = time.time()
        r = fir.fir_ap_cvx_sweep(256, f, S["a"], S["d"], objs, peaks, fadds, batch=512, method="ipm")
        dt = time.time() - t0
    info = r["info"]; stt = info[:, 0]
    print(f"   sweep 512: {dt:.2f}s = {512/dt:.0f}/s; solved {int((stt==1).sum())} infeasible {int((stt==2).sum())} limit {int((stt==3).sum())}; iters mean {info[:,1].mean():.1f} solved-mean {info[stt==1,1].mean():.1f}", flush=True)
ms = C.c_float()
lib.mbrf_ipm_cholesky_bench(512, 512, 1, 3, C.byref(ms)); print("cholesky dd 512x512:", ms.value, "ms")
lib.mbrf_ipm_cholesky_bench(512, 512, 0, 3, C.byref(ms)); print("cholesky fp64 512x512:", ms.value, "ms")
