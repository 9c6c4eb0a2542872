#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference C.

Runs only in the build container (needs /root/reference compiled into oracle/_ref by
oracle/Makefile).  Outputs small .npz files next to this script:

  pulses.npz      the workload pulses of BASELINE configs 1 and 2, made with the reference's
                  own recipe  rf = b2rf(sqrt(1/2)*msinc(n, 2))  (dzrf.m:40,62,81; msinc.m:12-15;
                  rf_tools/mex5/b2rf.c) and scaled to Gauss by rfscaleg (rfscaleg.m:10-12)
  bloch_*.npz     inputs + outputs of the reference mexFunction (blochC.c:514 / blochH.c:514)
  abrx_*.npz      inputs + outputs of the reference abrx mexFunction (abrx.c:35)

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref  # noqa: E402

ref.build()
assert ref.have_ref(), "oracle/_ref missing: this script needs /root/reference"


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print("wrote", name, {k: np.asarray(v).shape for k, v in kw.items()})


# ---- workload pulses -------------------------------------------------------
rf256 = ref.dzrf_ms_ex(256, 8)          # radians, sum = pi/2
rf512 = ref.dzrf_ms_ex(512, 8)
save("pulses.npz",
     rf256_rad=rf256, rf512_rad=rf512,
     b1_cfg1_gauss=ref.rfscaleg(rf256, 4.0, 1.0705),       # T = 4 ms, gamma(13C) = 1.0705 kHz/G
     b1_cfg2_gauss=ref.rfscaleg(rf512, 8.0, 1.0705))       # T = 8 ms

# ---- config 1: dzrf 256-sample pulse, 2000 offsets, blochC mode 0 -----------
b1 = ref.rfscaleg(rf256, 4.0, 1.0705)
df = np.linspace(-4000, 4000, 2000)
mx, my, mz = ref.bloch_mex_ref("C-13", b1.reshape(-1, 1), np.zeros((256, 1)), 4e-3 / 256, 1e3, 1e3,
                               df.reshape(-1, 1), 0.0, 0)
save("bloch_cfg1.npz", b1=b1, df=df, dt=4e-3 / 256, t1=1e3, t2=1e3, mx=mx, my=my, mz=mz)

# ---- randomised parity set: seed 0 ----------------------------------
rng = np.random.default_rng(0)
case = 0
for nuc in ("C-13", "H-1"):
    for nt, dt, t1, t2 in ((96, 4e-6, 1e3, 1e3), (160, 2e-5, 0.5, 0.05)):
        for mode in (0, 2):
            for ngr, npd in ((1, 1), (3, 3), (2, 2)):
                b1 = rng.normal(0, 0.05, nt) + 1j * rng.normal(0, 0.05, nt)
                gr = rng.normal(0, 0.4, (nt, ngr))
                df = rng.uniform(-4000, 4000, 9)
                dp = rng.uniform(-4, 4, (5, npd))
                m0 = rng.normal(0, 0.5, (3, 5, 9))
                args = [b1.reshape(-1, 1), gr, dt, t1, t2, df, dp, mode]
                if case % 2:
                    args += [m0[0], m0[1], m0[2]]
                mx, my, mz = ref.bloch_mex_ref(nuc, *args)
                save(f"bloch_rand_{case:02d}.npz", nucleus=nuc, b1=b1, gr=gr, dt=dt, t1=t1, t2=t2, df=df, dp=dp,
                     mode=mode, use_m0=case % 2, m0=m0, mx=mx, my=my, mz=mz)
                case += 1

# time-vector semantics (blochC.c:660-681): end times vs intervals, real b1, 1xN positions
nt = 64
b1 = rng.normal(0, 0.05, nt)
gr = rng.normal(0, 0.4, (nt, 1))
iv = rng.uniform(2e-6, 3e-5, nt)
df = rng.uniform(-2000, 2000, 6)
dp = rng.uniform(-3, 3, (1, 7))
for name, tp in (("endtimes", np.cumsum(iv)), ("intervals", iv), ("equal_intervals", np.full(nt, 1e-5))):
    mx, my, mz = ref.bloch_mex_ref("C-13", b1, gr, tp, 0.8, 0.08, df, dp, 0)
    save(f"bloch_time_{name}.npz", b1=b1, gr=gr, tp=tp, t1=0.8, t2=0.08, df=df, dp=dp, mx=mx, my=my, mz=mz)

# large rotations per sample (phi up to several turns): exercises every polynomial tier + sincos path
nt = 48
b1 = (rng.normal(0, 1.0, nt) + 1j * rng.normal(0, 1.0, nt)) * np.linspace(0.01, 3.0, nt)
df = np.concatenate([np.linspace(-100, 100, 5), np.linspace(-2e4, 2e4, 6), [1e5, -3e5]])
mx, my, mz = ref.bloch_mex_ref("C-13", b1.reshape(-1, 1), np.zeros((nt, 1)), 2e-5, 1e3, 1e3, df, 0.0, 0)
save("bloch_bigangle.npz", b1=b1, df=df, dt=2e-5, mx=mx, my=my, mz=mz)

# ---- abrx ---------------------------------------------------------------------
for i, (ns, cplx_rf, use_y) in enumerate(((128, True, False), (200, False, False), (96, True, True))):
    rf = rng.normal(0, 0.03, ns) + (1j * rng.normal(0, 0.03, ns) if cplx_rf else 0)
    g = rng.normal(0, 0.2, ns) + 1j * rng.normal(0, 0.2, ns)
    x = np.linspace(-8, 8, 33)
    y = np.linspace(-3, 3, 5) if use_y else None
    (a, b), err = ref.abrx_mex_ref(rf, g if use_y else g.real, x, y)
    assert err is None
    save(f"abrx_{i}.npz", rf=rf, g=g if use_y else g.real, x=x, y=(y if use_y else np.zeros(0)), use_y=use_y,
         alpha=a, beta=b)
# the SLR design identity: an 'ex' pulse's beta profile (default gradient 2*pi/N, abr.m:25)
g = np.ones(256) * 2 * np.pi / 256
x = np.linspace(-16, 16, 257)
(a, b), err = ref.abrx_mex_ref(rf256, g, x)
save("abrx_dzrf256.npz", rf=rf256, g=g, x=x, alpha=a, beta=b)
