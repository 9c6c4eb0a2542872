#!/usr/bin/env python
"""e2e timing of the drop-in Bloch call with PAGEABLE result arrays (what a MEX gateway has) vs page-locked ones, and over
several devices in one call (mbrf_set_fanout)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
import torch
lib = m.lib()
ndev = lib.mbrf_device_count()
b1 = np.load("tests/golden/pulses.npz")["b1_cfg2_gauss"].astype(np.complex128)
nt = b1.size
for fan in sorted({1, ndev}):
    lib.mbrf_set_fanout(fan)
    nf = 1000 * fan
    df = np.linspace(-5000, 5000, nf).reshape(-1, 1); dp = np.linspace(-5, 5, 1000).reshape(-1, 1)
    gr = np.full((nt, 1), 0.05)
    n = nf * 1000
    pinned = tuple(torch.empty(n, dtype=torch.float64).pin_memory().numpy() for _ in range(3))
    pageable = tuple(np.empty(n) for _ in range(3))
    ref = None
    for name, out in (("pinned", pinned), ("pageable", pageable), ("fresh pageable (MATLAB-like: new arrays per call)", None)):
        for _ in range(3):
            r = m.blochC(b1, gr, 8e-3 / nt, 1e3, 1e3, df, dp, 0, out=out)
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            r = m.blochC(b1, gr, 8e-3 / nt, 1e3, 1e3, df, dp, 0, out=out)
            ts.append(time.perf_counter() - t0)
        if ref is None:
            ref = [a.copy() for a in r]
        err = max(float(np.abs(a - b).max()) for a, b in zip(r, ref))
        print(f"fanout {fan} {name:50s}: min {min(ts)*1e3:.3f} ms median {sorted(ts)[5]*1e3:.3f} ms -> {n*nt/min(ts):.3e} spin-steps/s  (diff vs pinned {err:.1e})", flush=True)
    if fan > 1:   # the same job on one device must give the same numbers
        lib.mbrf_set_fanout(1)
        r1 = m.blochC(b1, gr, 8e-3 / nt, 1e3, 1e3, df, dp, 0)
        print("fanout vs single device max diff", max(float(np.abs(a - b).max()) for a, b in zip(r1, ref)))
# mode 2 + M0 + SLR through the pipeline
lib.mbrf_set_fanout(ndev)
rng = np.random.default_rng(0)
nts = 64
b1s = rng.normal(0, 0.05, nts) + 1j * rng.normal(0, 0.05, nts)
df = np.linspace(-3000, 3000, 700); dp = np.linspace(-2, 2, 300).reshape(-1, 1)
m0 = [rng.normal(0, 0.3, (300, 700)) for _ in range(3)]
a = m.blochC(b1s, np.full((nts, 1), 0.1), 1e-5, 0.5, 0.05, df, dp, 2, *m0)
lib.mbrf_set_fanout(1)
b = m.blochC(b1s, np.full((nts, 1), 0.1), 1e-5, 0.5, 0.05, df, dp, 2, *m0)
print("mode 2 + M0, fanout vs single:", max(float(np.abs(x - y).max()) for x, y in zip(a, b)), a[0].shape)
from oracle import ref as R
R.build()
w = R.blochsimfz_oracle(b1s, np.full(nts, 0.1), None, None, 1e-5, 0.5, 0.05, df[:5], dp[:, 0], mode=0)
g = m.blochC(b1s, np.full((nts, 1), 0.1), 1e-5, 0.5, 0.05, df[:5], dp, 0)
print("vs oracle:", max(float(np.abs(x.ravel(order='F') - y).max()) for x, y in zip(g, w)))
rf = rng.normal(0, 0.03, 200) + 1j * rng.normal(0, 0.03, 200)
x = np.linspace(-8, 8, 400001)
lib.mbrf_set_fanout(ndev)
t0 = time.perf_counter(); a1 = m.abrx(rf, np.ones(200) * 0.1, x); t1 = time.perf_counter() - t0
lib.mbrf_set_fanout(1)
a2 = m.abrx(rf, np.ones(200) * 0.1, x)
print("abrx fanout vs single:", float(np.abs(a1[0] - a2[0]).max()), float(np.abs(a1[1] - a2[1]).max()), f"{t1*1e3:.2f} ms")
