#!/usr/bin/env python
"""Known answers for the convex FIR design step, from an INDEPENDENT CPU solver (HiGHS via SciPy) on the
problems restated from fir_ap_cvx.m (oracle/fir_problems.py).  The reference's own solver (CVX) is not
available offline, so these are the pins for the solver path.

For each case: cone-free LP optimum, and the optimum with every 2-D peak cone replaced by its circumscribed
(outer) and inscribed (inner) 32-gon -> lower / upper bounds on the true SOCP optimum.

    python tests/golden/make_golden_fir.py        (n = 256 cases take ~1 min each)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.fir_problems import H1_DUALBAND, build_fir_ap, solve_fir_ap_highs  # noqa: E402

LOWPASS = dict(f=[-1, -0.5, -0.2, 0.2, 0.5, 1], a=[0, 0, 1, 1, 0, 0], d=[0.01, 0.02, 0.01])
TWOBAND = dict(f=[-0.8, -0.45, -0.3, -0.05, 0.1, 0.35, 0.5, 0.9], a=[0, 0, 1, 1, 0, 0, 0.5, 0.5], d=[0.02, 0.02, 0.02, 0.03])
CASES = [
    ("lowpass_n24", 24, LOWPASS, 0.1, 0.02),
    ("lowpass_n24_obj10", 24, LOWPASS, 10.0, 0.02),
    ("lowpass_n24_tightpeak", 24, LOWPASS, 0.1, 0.0105),
    ("lowpass_n24_peak_infeasible", 24, LOWPASS, 0.1, 1e-3),
    ("lowpass_n10_infeasible", 10, LOWPASS, 0.1, 0.05),
    ("twoband_n40", 40, TWOBAND, 1.0, 0.01),
    ("h1_dualband_n256", 256, H1_DUALBAND, 0.1, 1e-3),
    ("h1_dualband_n128_infeasible", 128, H1_DUALBAND, 0.1, 1e-3),
]
out = {}
for name, n, spec, obj, peak in CASES:
    if len(sys.argv) > 1 and sys.argv[1] not in name:
        continue
    t0 = time.time()
    p = build_fir_ap(n, spec["f"], spec["a"], spec["d"], obj, peak)
    free, _ = solve_fir_ap_highs(p, 0)
    rec = dict(n=n, f=list(map(float, spec["f"])), a=list(map(float, spec["a"])), d=list(map(float, spec["d"])),
               obj=obj, peak=peak, rows=int(p["w"].size), stop_rows=int(p["stop"].size),
               cone_free_status=int(free.status), cone_free_obj=float(free.fun) if free.status == 0 else None)
    if free.status == 0 and n <= 64:
        outer, inner = solve_fir_ap_highs(p, 32)
        rec.update(outer_status=int(outer.status), outer_obj=float(outer.fun) if outer.status == 0 else None,
                   inner_status=int(inner.status), inner_obj=float(inner.fun) if inner.status == 0 else None)
    rec["seconds"] = round(time.time() - t0, 1)
    out[name] = rec
    print(name, rec, flush=True)
path = os.path.join(HERE, "fir_ap_known.json")
if os.path.exists(path) and len(sys.argv) > 1:
    old = json.load(open(path))
    old.update(out)
    out = old
json.dump(out, open(path, "w"), indent=1)
