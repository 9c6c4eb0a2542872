"""Timing of the batched zero flipping (mbrf_flip_zero_batch) at the reference's cap of 2^12 patterns, N = 256 taps,
against the numpy restatement of the reference's loop (oracle/fir_post.py) on one host core for a sample of the patterns."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from multiband_rf_pulse_design_b200 import fir_post as P
from oracle import fir_post as O
rng = np.random.default_rng(0)
nsb, npb = 243, 12
zs = np.exp(1j * np.linspace(0.25 * np.pi, 1.75 * np.pi, nsb))
zp = rng.uniform(0.6, 0.9, npb) * np.exp(1j * rng.uniform(-0.2, 0.2, npb) * np.pi)
Z = np.concatenate([zs, zp])[rng.permutation(nsb + npb)]
idx = np.nonzero(np.abs(np.abs(Z) - 1) > 1e-2)[0]
mask = P.flip_patterns(idx.size)
P.flip_zero_candidates(Z, idx, mask, 1.0)
t = time.perf_counter()
for _ in range(5):
    r = P.flip_zero_candidates(Z, idx, mask, 1.0)
gpu = (time.perf_counter() - t) / 5
t = time.perf_counter()
for i in range(64):
    Ze = Z.copy(); Ze[idx] = np.where(mask[i].astype(bool), 1 / np.abs(Z[idx]) * np.exp(1j * np.angle(Z[idx])), Z[idx])
    O.poly_reference(Ze)
cpu = (time.perf_counter() - t) / 64 * mask.shape[0]
print(f"flip_zero N={Z.size+1} patterns={mask.shape[0]}: GPU call {gpu*1e3:.2f} ms, numpy loop (extrapolated from 64) {cpu:.2f} s, best={r['best']}")
