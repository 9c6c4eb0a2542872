#!/usr/bin/env python
"""Developer check of the GPU FIR solver (not a test): solve a few fir_ap_cvx designs, report objective,
violation (recomputed on the CPU from the returned x with the problem restated by the oracle) and time."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import multiband_rf_pulse_design_b200 as m  # noqa: E402
from multiband_rf_pulse_design_b200 import fir  # noqa: E402
from oracle.fir_problems import H1_DUALBAND, build_fir_ap, violation_fir_ap  # noqa: E402

LOWPASS = dict(f=[-1, -0.5, -0.2, 0.2, 0.5, 1], a=[0, 0, 1, 1, 0, 0], d=[0.01, 0.02, 0.01])


def run(n, spec, objs, peaks, **kw):
    B = len(objs)
    t0 = time.time()
    hs, st, ex = fir.fir_ap_cvx_batch(n, [spec["f"]] * B, spec["a"], spec["d"], objs, peaks, return_info=True, **kw)
    dt = time.time() - t0
    for b in range(min(B, 6)):
        p = build_fir_ap(n, spec["f"], spec["a"], spec["d"], objs[b], peaks[b])
        z = np.concatenate([ex["x"][b], [ex["ripple_stop"][b]]])
        print(f"  n={n} obj={objs[b]} Peak={peaks[b]}: {st[b]} info={np.array2string(ex['info'][b], precision=7)} "
              f"c'z={p['c'] @ z:.8f} viol={violation_fir_ap(p, z):.2e}")
    print(f"  batch of {B}: {dt:.2f} s wall  ({B / dt:.2f} designs/s), launches so far {m.lib().mbrf_launch_count()}")


import os as _os
m.lib().mbrf_pdhg_set_gemm(int(_os.environ.get("MBRF_DMMA", "1")))
which = sys.argv[1] if len(sys.argv) > 1 else "small"
if which in ("small", "all"):
    run(24, LOWPASS, [0.1, 10.0, 0.1, 0.1], [0.02, 0.02, 0.0105, 1e-3])
    run(10, LOWPASS, [0.1], [1e-3], max_iter=20000)
if which in ("n256", "all"):
    run(256, H1_DUALBAND, [0.1], [1e-3])
if which in ("batch", "all"):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    objs = list(np.logspace(-2, 1, B))
    run(256, H1_DUALBAND, objs, [1e-3] * B, max_iter=int(sys.argv[3]) if len(sys.argv) > 3 else 30000)
