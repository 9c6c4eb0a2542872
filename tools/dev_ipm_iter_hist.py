"""Distribution of Newton iterations by final status over the cfg4 sweep (4096 designs): where do the stragglers sit?"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import H1_DUALBAND as S
f = np.array(S["f"]); df = float((f[2:-1:2] - f[1:-2:2]).min())
no = int(sys.argv[1]) if len(sys.argv) > 1 else 16
objs, peaks, fadds = np.logspace(-2, 4, no), np.logspace(-4, -2, no), np.linspace(0, 0.9 * df / 2, no)
fir.fir_ap_cvx_sweep(256, f, S["a"], S["d"], objs[:2], peaks[-2:], fadds[:2], batch=8)
t = time.perf_counter()
conc = int(sys.argv[2]) if len(sys.argv) > 2 else 2
order = sys.argv[3] if len(sys.argv) > 3 else "grouped"
r = fir.fir_ap_cvx_sweep(256, f, S["a"], S["d"], objs, peaks, fadds, batch=int(sys.argv[4]) if len(sys.argv) > 4 else 512, concurrent_batches=conc, order=order)
print("concurrent", conc, "order", order)
dt = time.perf_counter() - t
info = r["info"]
st, it = info[:, 0], info[:, 1]
print(f"{len(st)} designs in {dt:.2f} s = {len(st)/dt:.1f} designs/s")
for code, name in ((1, "solved"), (2, "infeasible"), (3, "limit")):
    v = it[st == code]
    if v.size:
        print(name, v.size, "iterations: min %d p50 %d p90 %d p99 %d max %d" % (v.min(), np.percentile(v, 50), np.percentile(v, 90), np.percentile(v, 99), v.max()),
              "hist(10s):", np.histogram(v, bins=np.arange(0, 111, 10))[0].tolist())
