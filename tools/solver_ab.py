"""A/B of the PDHG product kernels on the cfg4 slice: fp64 DMMA (mode 1) vs tcgen05 int8 split-integer (mode 2, ND digits).
usage: python tools/solver_ab.py [designs] [modes like 1,2:6,2:5]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND

lib = m.lib()
designs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2:6", "2:5"]
n = 256
objs = np.logspace(-2, 1, max(1, designs // 8))
peaks = np.logspace(-3.2, -2, 8)
fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs[:8], peaks, [0.0], max_iter=512)
ref = None
for md in modes:
    mode, nd = (md.split(":") + ["6"])[:2]
    assert lib.mbrf_pdhg_set_gemm(int(mode)) == 0
    assert lib.mbrf_pdhg_set_tc_digits(int(nd)) == 0
    t0 = time.perf_counter()
    r = fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, [0.0], batch=designs, max_iter=int(os.environ.get('MAXIT', '60000')))
    sec = time.perf_counter() - t0
    info = r["info"]
    if ref is None:
        ref = info
    dobj = np.abs(info[:, 2] - ref[:, 2]) / np.maximum(np.abs(ref[:, 2]), 1e-12)
    print(f"mode {md}: {sec:.2f} s, {designs / sec:.1f} designs/s, solved {(info[:, 0] == 1).sum()}/{designs}, iters mean {info[:, 1].mean():.0f} max {info[:, 1].max():.0f}, "
          f"max viol {info[:, 4].max():.2e}, max rel obj diff vs first mode {dobj.max():.2e}", flush=True)
