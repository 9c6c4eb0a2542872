#!/usr/bin/env python
"""Known answers of the fir_ap_cvx LP at N = 256 on the dual-band H-1 spec for the stop-band weights the reference's own
callers use (obj = 1e4: dzrf_mb.m:167-170; 1e5: fir_qp.m:47) and for widened band edges (fir_ap.m:70-83), from HiGHS
(scipy.optimize.linprog, dual simplex) on the problem restated by oracle/fir_problems.py.  The peak cones are dropped (the
GPU tests use Peak = 1, which leaves them inactive).  Each case takes 1-13 minutes on one core:

    python tests/golden/make_golden_fir_weights.py [name-substring]      -> tests/golden/fir_ap_weights_known.json
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.fir_problems import H1_DUALBAND, build_fir_ap, solve_fir_ap_highs  # noqa: E402

f0 = np.array(H1_DUALBAND["f"])
DF_MIN = float((f0[2:-1:2] - f0[1:-2:2]).min())            # fir_ap.m:60-61


def widen(fa):
    f = f0.copy()
    f[0::2] -= fa
    f[1::2] += fa
    return f


# (name, obj, f_add as a fraction of df_min / 2)
CASES = [("w4", 4.0, 0.0), ("w10", 10.0, 0.0), ("w100", 100.0, 0.0), ("w1e4", 1e4, 0.0), ("w1e5", 1e5, 0.0),
         # widened band edges: N = 256 stays feasible only up to f_add ~ 0.15 * df_min / 2 (HiGHS returns status 4 at 0.3 and 0.6:
         # recorded, the GPU solvers must fail there too)
         ("w0.01_fa0.12", 0.01, 0.12), ("w30_fa0.06", 30.0, 0.06), ("w1_fa0.06", 1.0, 0.06), ("w1e3_fa0.12", 1e3, 0.12),
         ("w0.1_fa0.12", 0.1, 0.12), ("w1e4_fa0.06", 1e4, 0.06),
         ("w0.1_fa0.3", 0.1, 0.3), ("w10_fa0.6", 10.0, 0.6), ("w1_fa0.6", 1.0, 0.6)]
path = os.path.join(HERE, "fir_ap_weights_known.json")
for name, obj, frac in CASES:
    if len(sys.argv) > 1 and sys.argv[1] != name:
        continue
    t0 = time.time()
    fa = frac * DF_MIN / 2
    f = widen(fa)
    p = build_fir_ap(256, f, H1_DUALBAND["a"], H1_DUALBAND["d"], obj, 1.0)
    res, _ = solve_fir_ap_highs(p, 0)
    rec = dict(n=256, f=[float(v) for v in f], a=list(map(float, H1_DUALBAND["a"])), d=list(map(float, H1_DUALBAND["d"])), obj=obj,
               f_add=fa, peak=1.0, rows=int(p["w"].size), status=int(res.status), cone_free_obj=float(res.fun) if res.status == 0 else None,
               ripple_stop=float(res.x[-1]) if res.status == 0 else None, x1=float(res.x[0]) if res.status == 0 else None,
               seconds=round(time.time() - t0, 1))
    print(name, rec, flush=True)
    old = json.load(open(path)) if os.path.exists(path) else {}
    old[name] = rec
    json.dump(old, open(path + ".tmp", "w"), indent=1)
    os.replace(path + ".tmp", path)
