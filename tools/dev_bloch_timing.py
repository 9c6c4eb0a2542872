#!/usr/bin/env python
"""Developer timing sweep of the Bloch kernel on one GPU (not the bench): residency x spins/thread."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import multiband_rf_pulse_design_b200 as m  # noqa: E402
from multiband_rf_pulse_design_b200._lib import check  # noqa: E402

lib = m.lib()
dev = torch.device("cuda:0")
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "pulses.npz"))
b1 = g["b1_cfg2_gauss"]
nt = b1.size
dt = 8e-3 / nt


def T(a):
    return torch.tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)


tf = C.c_double()
ms = C.c_double()
check(lib.mbrf_measure_fp64_peak(C.byref(tf), C.byref(ms)))
print(f"fp64 peak (DFMA kernel): {tf.value:.2f} TFLOP/s ({ms.value:.3f} ms)")

nf, npos = 1000, 1000
b1r, b1i = T(b1.real), T(b1.imag)
gx = T(np.full(nt, 0.05))
dts = T(np.full(nt, dt))
df = T(np.linspace(-5000, 5000, nf))
dx = T(np.linspace(-5, 5, npos))
ns = nf * npos
out = [torch.empty(ns, dtype=torch.float64, device=dev) for _ in range(3)]
ws = torch.empty(int(lib.mbrf_bloch_workspace_bytes(nt)), dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def run(ngrad=1, mode=0):
    check(lib.mbrf_bloch_device(b1r.data_ptr(), b1i.data_ptr(), gx.data_ptr() if ngrad else None, None, None,
                                dts.data_ptr(), nt, 1e3, 1e3, df.data_ptr(), nf, dx.data_ptr(), None, None, npos,
                                0, ns, None, None, None, 1, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                mode, m.GAMMA_C13, ws.data_ptr(), stream))


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for spt in (1, 2):
    for bps in (0, 6, 8, 1000):
        check(lib.mbrf_bloch_set_tuning(bps, spt))
        for ngrad in (1, 0):
            t = timeit(lambda: run(ngrad))
            print(f"spt={spt} ctas/sm={bps:2d} ngrad={ngrad}: {t:.3f} ms  {ns * nt / t / 1e6:.1f} Gspin-steps/s")
check(lib.mbrf_bloch_set_tuning(0, 0))

# modes 1, 2, 3 at a smaller spin count (mode 2/3 write ntime values per spin)
nf2 = 200
ns2 = nf2 * npos
out2 = [torch.empty(ns2 * nt, dtype=torch.float64, device=dev) for _ in range(3)]
for mode in (2, 1, 3):
    def run_m():
        check(lib.mbrf_bloch_device(b1r.data_ptr(), b1i.data_ptr(), gx.data_ptr(), None, None, dts.data_ptr(), nt, 1e3, 1e3,
                                    df.data_ptr(), nf2, dx.data_ptr(), None, None, npos, 0, ns2, None, None, None, 1,
                                    out2[0].data_ptr(), out2[1].data_ptr(), out2[2].data_ptr(), mode, m.GAMMA_C13,
                                    ws.data_ptr(), stream))
    t = timeit(run_m, 3)
    gb = ns2 * nt * 24 / 1e9 if mode & 2 else 0
    print(f"mode {mode}: {ns2} spins x {nt}: {t:.3f} ms  {ns2 * nt / t / 1e6:.1f} Gspin-steps/s" + (f"  store {gb / t * 1e3:.0f} GB/s" if gb else ""))
