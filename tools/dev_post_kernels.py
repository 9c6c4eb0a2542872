"""The kernels added for SURVEY.md 8(f) rows 3-4, run once each at their full sizes (for ncu and for CUDA-event-free wall
timing): device-side assembly of a 512-design fir_ap_cvx batch at N = 256 (mbrf_fir_ap_assemble) and the zero flipping of
a 256-tap filter over 2^12 patterns (mbrf_flip_zero_batch)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from multiband_rf_pulse_design_b200 import fir, fir_post as P
from oracle.fir_problems import H1_DUALBAND as S

f = np.array(S["f"]); df = float((f[2:-1:2] - f[1:-2:2]).min())
fl, ol, pl = fir.sweep_grid(f, np.logspace(-2, 4, 8), np.logspace(-4, -2, 8), np.linspace(0, 0.9 * df / 2, 8))
B = len(fl)
fir.assemble_fir_ap_device(256, fl[:8], [S["a"]] * 8, [S["d"]] * 8, ol[:8], pl[:8])
t = time.perf_counter()
dev = fir.assemble_fir_ap_device(256, fl, [S["a"]] * B, [S["d"]] * B, ol, pl)
t_dev = time.perf_counter() - t
t = time.perf_counter()
host = fir._assemble_batch_ap(256, [fir.assemble_fir_ap(256, fl[i], S["a"], S["d"], ol[i], pl[i]) for i in range(B)])
t_host = time.perf_counter() - t
same = all(np.array_equal(dev[k], host[k]) for k in ("w_row", "lo", "hi", "c", "bl", "bu", "rho", "sw"))
print(f"assembly B={B} N=256 rows={dev['M']} ({dev['M1']} grid + {dev['ns']} stop): device call incl. D2H of the arrays {t_dev*1e3:.1f} ms, "
      f"numpy {t_host*1e3:.1f} ms, bit-identical={same}")
rng = np.random.default_rng(0)
zs = np.exp(1j * np.linspace(0.25 * np.pi, 1.75 * np.pi, 243))
zp = rng.uniform(0.6, 0.9, 12) * np.exp(1j * rng.uniform(-0.2, 0.2, 12) * np.pi)
Z = np.concatenate([zs, zp])[rng.permutation(255)]
idx = np.nonzero(np.abs(np.abs(Z) - 1) > 1e-2)[0]
mask = P.flip_patterns(idx.size)
P.flip_zero_candidates(Z, idx, mask, 1.0)
t = time.perf_counter()
r = P.flip_zero_candidates(Z, idx, mask, 1.0)
print(f"flip_zero N=256 patterns={mask.shape[0]}: {1e3*(time.perf_counter()-t):.2f} ms per call")
