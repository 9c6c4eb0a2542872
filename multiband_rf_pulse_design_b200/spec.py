"""Host-side mirror of the reference's specification builders: from flip angles, magnetisation
ripples and chemical shifts to the band specification (f, a, d) that the FIR design step takes.

    range_B, ripple_B = rf_ripple_GFA(FA, ripple_M, ptype, appro, dbg)          rf_ripple_GFA.m
    range_M           = rf_Mrange_desired(FA, ripple_M, ptype)                   rf_Mrange_desired.m
    fn                = rf_bandedge(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype, dbg)   rf_bandedge.m
    d                 = dinf(d1, d2)                                             dinf.m
    f, names          = spectrum_C13(B0, dbg)                                    spectrum_C13.m
    b_spec, rf_spec, shift = multiband_spec(...)                                 dzrf_mb.m:95-157 (the part of the driver before the solve)

O(number of bands) scalar arithmetic per design: it stays on the host (it is what produces the few numbers per design that
`mbrf_fir_ap_solve` takes).  `rf_ripple_GFA` and `rf_Mrange_desired` accept arrays, so that a sweep over flip angles / ripples
builds its (a, d) columns in one vectorised call.
"""
from __future__ import annotations

import numpy as np

PTYPES = ("st", "ex", "se", "inv", "sat")


def _check_fa(FA):
    FA = np.asarray(FA, float)
    if np.any(FA < 0) or np.any(FA > 180):                                 # rf_ripple_GFA.m:170-174, rf_Mrange_desired.m
        raise ValueError("Flip angle should be in the range of [0 180] degree")
    return FA * np.pi / 180


def _fa2beta(rfa_r, rfa_l, FA):
    """rf_ripple_FA2Beta, rf_ripple_GFA.m:263-295: flip-angle range -> |beta| = sin(theta/2) range."""
    over = rfa_r > np.pi                                                   # :281-286: the range reaches 180 degrees
    min_B = np.where(over, np.minimum(np.sin(rfa_l / 2), np.sin(rfa_r / 2)), np.sin(rfa_l / 2))
    max_B = np.where(over, 1.0, np.sin(rfa_r / 2))
    mid_B = np.sin(FA / 2)
    rip_lo = np.where(over, 1.0 - min_B, np.abs(min_B - mid_B))
    rip_hi = np.where(over, 1.0 - min_B, np.abs(max_B - mid_B))
    return np.stack([min_B, max_B], -1), np.stack([rip_lo, rip_hi], -1)


def rf_ripple_GFA(FA, ripple_M, ptype, appro=0, dbg=0):
    """[range_B, ripple_B] = rf_ripple_GFA(FA, ripple_M, ptype, appro, dbg) — rf_ripple_GFA.m.

    Range of |beta| (the FIR amplitude) that keeps the magnetisation of a band with flip angle FA (degrees) within
    +- ripple_M, exact (asin / acos form, :166-260) or with the quadratic approximation (appro != 0, :84-163).  Arrays
    broadcast; the last axis of the results is (min, max).  When the range reaches 180 degrees the reference returns a
    scalar ripple_B = 1 - min_B (:281-286); here both entries carry that value."""
    if ptype not in PTYPES:
        raise ValueError(f"Unrecognized Pulse Type -- {ptype}; recognized types are st, ex, se, inv, and sat")
    fa = _check_fa(FA)
    rip = np.asarray(ripple_M, float)
    fa, rip = np.broadcast_arrays(fa, rip)
    if ptype == "st":
        # the asin branch of the reference reads mid_M before assigning it (rf_ripple_GFA.m:201-203): an error
        # there; the quadratic branch defines it as sin(FA) (:99), which is what the small-tip model means
        mid = np.sin(fa)
        rng = np.stack([mid - rip, np.minimum(1.0, mid + rip)], -1)
        return rng, np.abs(rng - mid[..., None])
    if ptype == "se":                                                      # :243-251
        mid_M = np.sin(fa / 2) ** 2
        rng = np.sqrt(np.stack([np.clip(mid_M - rip, 0, 1), np.clip(mid_M + rip, 0, 1)], -1))
        return rng, np.abs(rng - np.sin(fa / 2)[..., None])
    if appro:
        # second-order expansion of sin / cos around FA: roots of -1/2 f(FA) x^2 +- f'(FA) x +- ripple (:104-147)
        s, c = (np.sin(fa), np.cos(fa)) if ptype == "ex" else (np.cos(fa), np.sin(fa))
        if ptype == "ex":
            hit = s + rip >= 1
            sr = np.where(hit, 1.0, np.where(fa <= np.pi / 2, -1.0, 1.0))
            sl = np.where(hit, 1.0, np.where(fa <= np.pi / 2, 1.0, -1.0))
            br, bl = c, -c
        else:
            hi, lo = s + rip > 1, s - rip < -1
            sr = np.where(hi, 1.0, -1.0)
            sl = np.where(hi, 1.0, np.where(lo, -1.0, 1.0))
            br, bl = c, -c

        def small_root(a2, a1, a0):
            out = np.full(a2.shape, np.nan)
            for idx in np.ndindex(a2.shape):
                r = np.roots([a2[idx], a1[idx], a0[idx]])
                r = r[np.isreal(r) & (r.real > 0)].real
                if r.size:
                    out[idx] = r.min()
            return out
        dr = small_root(-0.5 * s, br, sr * rip)
        dl = small_root(-0.5 * s, bl, sl * rip)
        return _fa2beta(fa + dr, fa - dl, fa)
    if ptype == "ex":                                                      # :205-217
        with np.errstate(invalid="ignore"):
            lo_, hi_ = np.arcsin(np.sin(fa) - rip), np.arcsin(np.minimum(np.sin(fa) + rip, 1.0))
        hit = np.sin(fa) + rip >= 1
        acute = fa <= np.pi / 2
        rfa_l = np.where(hit, lo_, np.where(acute, lo_, np.pi - hi_))
        rfa_r = np.where(hit, np.pi - lo_, np.where(acute, hi_, np.pi - lo_))
        return _fa2beta(rfa_r, rfa_l, fa)
    # 'sat', 'inv' (:221-241)
    with np.errstate(invalid="ignore"):
        up = np.arccos(np.clip(np.cos(fa) - rip, -1, 1))
        dn = np.arccos(np.clip(np.cos(fa) + rip, -1, 1))
    hi, lo = np.cos(fa) + rip > 1, np.cos(fa) - rip < -1
    rfa_l = np.where(hi, -up, np.where(lo, dn, dn))
    rfa_r = np.where(hi, up, np.where(lo, 2 * np.pi - dn, up))
    return _fa2beta(rfa_r, rfa_l, fa)


def rf_Mrange_desired(FA, ripple_M, ptype):
    """range_M = rf_Mrange_desired(FA, ripple_M, ptype) — rf_Mrange_desired.m: the magnetisation range of a band."""
    if ptype not in PTYPES:
        raise ValueError(f"Unrecognized Pulse Type -- {ptype}")
    fa = _check_fa(FA)
    rip = np.asarray(ripple_M, float)
    if ptype in ("st", "ex"):
        return np.stack(np.broadcast_arrays(np.sin(fa) - rip, np.minimum(np.sin(fa) + rip, 1.0)), -1)
    if ptype in ("sat", "inv"):
        return np.stack(np.broadcast_arrays(np.maximum(np.cos(fa) - rip, -1.0), np.minimum(np.cos(fa) + rip, 1.0)), -1)
    m = np.sin(fa / 2) ** 2
    return np.stack(np.broadcast_arrays(np.maximum(m - rip, 0.0), np.minimum(m + rip, 1.0)), -1)


def dinf(d1, d2):
    """d = dinf(d1, d2) — dinf.m: the D-infinity measure of the Parks-McClellan length estimate (transition width x duration)."""
    a1, a2, a3, a4, a5, a6 = 5.309e-3, 7.114e-2, -4.761e-1, -2.66e-3, -5.941e-1, -4.278e-1
    l1, l2 = np.log10(np.asarray(d1, float)), np.log10(np.asarray(d2, float))
    return (a1 * l1 * l1 + a2 * l1 + a3) * l2 + (a4 * l1 * l1 + a5 * l1 + a6)


def rf_bandedge(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype, dbg=0):
    """fn = rf_bandedge(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype, dbg) — rf_bandedge.m: band edges normalised to [-1, 1].

    mb_cf: per band a centre frequency (scalar) or a (low, high) pair, kHz.  mb_range (kHz per band): each band is its centre
    (or range) widened by mb_range/2 on both sides (:133-144); empty / None: the bands fill the axis, separated by the
    transition width the D-infinity estimate gives for the pulse duration (:36-131)."""
    T = n * dt
    mb_FA = np.asarray(mb_FA, float)
    mb_ripple = np.asarray(mb_ripple, float)
    m_band = mb_FA.size
    fs = 1.0 / dt
    lo = np.array([np.atleast_1d(c)[0] for c in mb_cf], float)
    hi = np.array([np.atleast_1d(c)[-1] for c in mb_cf], float)
    f = np.empty(2 * m_band)
    if mb_range is None or np.size(mb_range) == 0:
        ref_on = {"st": 90, "ex": 90, "sat": 90, "inv": 180, "se": 180}
        if ptype not in ref_on:
            raise ValueError(f"Unrecognized Pulse Type -- {ptype}")
        d1 = mb_ripple[np.argmin(np.abs(mb_FA - ref_on[ptype]))]           # :38-90
        d2 = mb_ripple[np.argmin(np.abs(mb_FA))]
        delta1, delta2 = {"st": (np.sqrt(d1 / 2), d2 / np.sqrt(2)), "ex": (np.sqrt(d1 / 2), d2 / np.sqrt(2)),
                          "inv": (d1 / 8, np.sqrt(d2 / 2)), "sat": (d1 / 2, np.sqrt(d2)), "se": (d1 / 4, np.sqrt(d2))}[ptype]
        df = float(dinf(delta1, delta2)) / T                               # :98
        f[0::2], f[1::2] = lo, hi                                          # :100-111
        f1 = f.copy()
        mids = (f[1:-1:2] + f[2::2]) / 2                                   # :113-117
        f1[1:-1:2] = mids
        f1[2::2] = mids
        f1[0] = f[0] - (f1[1] - f[1])                                      # :118-119
        f1[-1] = f[-1] + (f[-2] - f1[-2])
        f = f1.copy()
        f[0::2] += df / 2                                                  # :121-125
        f[1::2] -= df / 2
    else:
        rng = np.asarray(mb_range, float)
        f[0::2], f[1::2] = lo - rng / 2, hi + rng / 2                      # :133-144
    if np.any(np.argsort(f, kind="stable") != np.arange(2 * m_band)):      # :146-151
        raise ValueError("Incompatible spec of frequency range: f is not monotonically increasing. "
                         "Try reducing mb_range (width for each band) or reorder bands.")
    if f[0] < -fs / 2 or f[-1] > fs / 2:                                   # :153-155
        raise ValueError("the sampling rate is not enough, increase n")
    return f / (fs / 2)                                                    # :157


def spectrum_C13(B0, dbg=0):
    """[f, name_cell] = spectrum_C13(B0, dbg) — spectrum_C13.m: C-13 resonances relative to pyruvate, Hz."""
    gamma = 10.705 * 1e6                                                   # :27
    cs = np.array([170.60, 182.98, 176.32, 178.91, 160.9, 163.13])         # :28-34
    names = ["Pyruvate", "Lactate", "Alanine", "Pyruvate-H_2O", "Bicarbonate", "Urea"]
    f0 = gamma * B0 * (1 + cs * 1e-6)                                      # :37
    return f0 - f0[0], names                                               # :38-39


def multiband_spec(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype="sat", shift_f=0, downsampling=1):
    """The specification part of dzrf_mb.m (:92-157), i.e. everything the driver does before it calls the FIR design step:
    downsampling of n / dt (:92-97), band edges (:100), (a, d) from the |beta| range of every band (:102-110), the
    magnetisation spec (:112-119) and the frequency shift that centres the design (:133-157).
    Returns dict(n, dt, f, a, d: what the design step is called with; b_spec, rf_spec: the structs the reference returns
    (UNSHIFTED edges divided by `downsampling`, :120-121); shift_f_back: kHz to shift the designed pulse back by)."""
    n = n / downsampling                                                   # :92-97
    dt = dt * downsampling
    if abs(n - round(n)) > 1e-10:
        raise ValueError("n/downsampling is not an integer")
    n = int(round(n))
    mb_FA = np.asarray(mb_FA, float)
    mb_ripple = np.asarray(mb_ripple, float)
    f = rf_bandedge(n, dt, mb_cf, mb_range, mb_FA, mb_ripple, ptype)       # :100
    rB, _ = rf_ripple_GFA(mb_FA, mb_ripple, ptype, 0)
    rM = rf_Mrange_desired(mb_FA, mb_ripple, ptype)
    a = np.repeat((rB[:, 1] + rB[:, 0]) / 2, 2)                            # :105-110
    d = (rB[:, 1] - rB[:, 0]) / 2
    a_M = np.repeat((rM[:, 1] + rM[:, 0]) / 2, 2)                          # :114-119
    d_M = (rM[:, 1] - rM[:, 0]) / 2
    b_spec = dict(f=f / downsampling, a=a, d=d)                            # :120-121 (before the shift)
    rf_spec = dict(f=f / downsampling, a=a_M, d=d_M)
    if shift_f == 0:                                                       # :134-136
        fwd = 0.0
    elif shift_f == 1:                                                     # :138-143: centre of the high-flip-angle bands at f = 0
        idx = np.nonzero(mb_FA > 60)[0]
        fwd = (f[2 * idx[0]] + f[2 * idx[-1] + 1]) / 2
    elif shift_f == 2:                                                     # :145-149: the highest-flip-angle band at f = 0
        k = int(np.argmax(mb_FA))
        fwd = (f[2 * k] + f[2 * k + 1]) / 2
    else:
        raise ValueError(f"shift_f = {shift_f} is not an option. Options are 0,1,2")
    return dict(n=n, dt=dt, f=f - fwd, a=a, d=d, b_spec=b_spec, rf_spec=rf_spec, shift_f_back=fwd * 0.5 * (1 / dt))
