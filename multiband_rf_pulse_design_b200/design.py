"""Host-side mirror of the reference's design driver:

    [rf_pulse, b, rf_spec, b_spec] = dzrf_mb(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype, ftype, nucleus, flip_zero,
                                              downsampling, Peak, dbg, min_order, min_tran, shift_f, name_cell)      dzrf_mb.m

The driver is orchestration (specification -> FIR design -> zero flipping -> inverse SLR -> scaling); every
stage it calls is one of this package's GPU-backed mirrors, so the example scripts of the toolbox (specsat_H1_dualband.m,
bSSFP_pulse_lp_ap.m, ...) have a one-call equivalent:  spec.multiband_spec  ->  fir_ap_cvx / fir_ap / fir_min_order_linprog /
fir_qp_cvx  ->  fir_flip_zero  ->  b2a + ab2rf  ->  rfscaleg.
"""
from __future__ import annotations

import numpy as np

from . import fir, fir_post, slr, spec

GAMMA_KHZ_PER_G = {"H-1": 4.2576, "C-13": 1.0705}                         # dzrf_mb.m:81-88


def rfscaleg(rf, t, gamma):
    """rfs = rfscaleg(rf, t, gamma) — rf_tools/rfscaleg.m:10-12: radians -> Gauss for a pulse of duration t (ms), gamma in kHz/G."""
    rf = np.asarray(rf)
    return rf / (2 * np.pi * gamma * (t / rf.size))


def dzrf_mb(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype="sat", ftype="ap_cvx", nucleus="C-13", flip_zero=0, downsampling=1,
            Peak=1e-3, dbg=0, min_order=0.9, min_tran=0.85, shift_f=0, name_cell=None, **solver_kw):
    """[rf_pulse, b, rf_spec, b_spec] = dzrf_mb(...) — dzrf_mb.m:1-300.  Same positional arguments and defaults (:60-73).

    ftype: 'ap_cvx' (fir_ap_cvx, obj = 1, :167), 'ap_minstopripple_cvx' (obj = 1e4, :170), 'ap_minorder_cvx' (:172-185),
    'ap_mintran_cvx' (:187-200), 'lp_minorder' (fir_min_order_linprog, :207-208), 'qp_cvx' (fir_qp_cvx, k = 120, obj = 1e6,
    :210-213).  Not mirrored: 'ms' (references an undefined TBW, :165), 'ap_mintran_minorder_cvx' (passes its arguments shifted
    by one, :203) and downsampling >= 2 (fir_upsample needs MATLAB's resample).  Returns empty arrays when the
    filter design fails, as the reference does (:216-219)."""
    if nucleus not in GAMMA_KHZ_PER_G:
        raise ValueError("No such option for nucleus. Options are H-1 and C-13")                  # :81-88
    if downsampling >= 2:
        raise NotImplementedError("downsampling >= 2 needs fir_upsample (MATLAB resample); design at the full rate instead")
    gamma = GAMMA_KHZ_PER_G[nucleus]
    s = spec.multiband_spec(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype, shift_f, downsampling)   # :92-157
    n, dt, f, a, d = s["n"], s["dt"], s["f"], s["a"], s["d"]
    b_spec, rf_spec = s["b_spec"], s["rf_spec"]
    shift_fwd = s["shift_f_back"] / (0.5 * (1 / dt))
    if Peak is None or np.size(Peak) == 0:
        Peak = 1e-3                                                         # :67
    if ftype == "ap_cvx":
        b, status = fir.fir_ap_cvx(n, f, a, d, 1, Peak, dbg, **solver_kw)                          # :167
    elif ftype == "ap_minstopripple_cvx":
        b, status = fir.fir_ap_cvx(n, f, a, d, 1e4, Peak, dbg, **solver_kw)                        # :170
    elif ftype == "ap_minorder_cvx":
        if min_order <= 1:                                                                         # :173-174
            b, status, _, _ = fir.fir_ap(n, f, a, d, Peak, min_order, 0, 0, dbg, **solver_kw)
        else:                                                                                      # :175-179: fixed n = min_order
            b, status = fir.fir_ap_cvx(int(min_order), f, a, d, 0.1, Peak, **solver_kw)
    elif ftype == "ap_mintran_cvx":
        b, status, _, f_new = fir.fir_ap(n, f, a, d, Peak, 0, min_tran, 0, dbg, **solver_kw)       # :188
        b_spec = dict(b_spec, f=(f_new + shift_fwd) / downsampling)                                # :198-199
        rf_spec = dict(rf_spec, f=(f_new + shift_fwd) / downsampling)
    elif ftype == "lp_minorder":
        b, status = fir.fir_min_order_linprog(n, f, a, d, 0, dbg, **solver_kw)                     # :207-208
    elif ftype == "qp_cvx":
        b, status = fir.fir_qp_cvx(n, f, a, d, 120, 1e6, dbg, **solver_kw)                         # :210-213
    else:
        raise ValueError(f"ftype {ftype!r} is not mirrored (options: ap_cvx, ap_minstopripple_cvx, ap_minorder_cvx, "
                         "ap_mintran_cvx, lp_minorder, qp_cvx)")
    if status == "Failed":                                                                         # :216-219
        print("Filter design failed.")
        return np.zeros(0), np.zeros(0), rf_spec, b_spec
    b = np.asarray(b)[::-1]                                                                        # :220
    if flip_zero:
        b = fir_post.fir_flip_zero(b, dbg)                                                         # :224-226
    b = np.asarray(b).ravel()                                                                      # :235
    if ptype == "st":
        rf = b                                                                                     # :236-237
    else:
        rf = slr.ab2rf(slr.b2a(b), b)                                                              # :239-240
    rf_pulse = rfscaleg(rf, dt * rf.size, gamma)                                                   # :244
    t_axis = np.arange(1, rf.size + 1) * dt                                                        # ms, :273
    rf_pulse = rf_pulse * np.exp(1j * 2 * np.pi * s["shift_f_back"] * t_axis)                      # :274
    return rf_pulse, b, rf_spec, b_spec
