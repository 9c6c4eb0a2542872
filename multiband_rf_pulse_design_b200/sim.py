"""Host-side mirror of the reference's verification scripts' compute part.

    sim_rf_scale(rf, dt, rfname, scale, nucleus, ptype, f, a, d)         sim_rf_scale.m

The reference runs one blochC / blochH call per B1 scaling (sim_rf_scale.m:82-89) and plots; here all scalings are ONE launch
of the sweep kernel (`mbrf_bloch_scale_sweep`: spin = (off-resonance, scale)), and the arrays the plots are drawn from are
returned instead of drawn.
"""
from __future__ import annotations

import numpy as np

from ._lib import c_double_p, check, lib
from .bloch import GAMMA_C13, GAMMA_H1


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def bloch_scale_sweep(b1, tp, t1, t2, df, scales, gamma=GAMMA_C13):
    """Magnetisation after the pulse b1 (Gauss, complex or real, time step tp seconds) for every off-resonance df (Hz) and every
    B1 scaling: returns mx, my, mz of shape [len(scales), len(df)], starting from (0, 0, 1).  One kernel launch."""
    b1 = np.asarray(b1).ravel()
    br = np.ascontiguousarray(b1.real, dtype=np.float64)
    bi = np.ascontiguousarray(b1.imag, dtype=np.float64) if np.iscomplexobj(b1) else None
    df = np.ascontiguousarray(np.asarray(df, float).ravel())
    sc = np.ascontiguousarray(np.asarray(scales, float).ravel())
    out = [np.empty((sc.size, df.size)) for _ in range(3)]
    check(lib().mbrf_bloch_scale_sweep(_dp(br), _dp(bi) if bi is not None else None, br.size, float(tp), float(t1), float(t2),
                                       _dp(df), df.size, _dp(sc), sc.size, _dp(out[0]), _dp(out[1]), _dp(out[2]), float(gamma)))
    return tuple(out)


def sim_rf_scale(rf, dt, rfname=None, scale=None, nucleus="C-13", ptype="ex", f=None, a=None, d=None):
    """sim_rf_scale(rf, dt, rfname, scale, nucleus, ptype, f, a, d) — sim_rf_scale.m:1-150 without the figures.

    rf in Gauss, dt in ms.  f alone (a, d None): the passband bandwidth in kHz, simulated range +-3 BW (:33-37); f, a, d: the
    multiband specification in kHz, range f(1) - 300 Hz .. f(end) + 300 Hz (:38-42).  scale defaults to [0.8 0.9 1 1.1 1.2]
    (:47-49).  Returns dict(df [Hz, 2048 points, :77], scale, mxy [nscale x 2048] complex, mz [nscale x 2048]) -- the arrays
    the reference plots (:92-140)."""
    if f is None:
        raise ValueError("Number of input should be either 7 or 9")       # :43-45
    if a is None and d is None:
        bw = float(np.ravel(f)[0]) * 1e3                                   # :35-36
        wrange = (-3 * bw, 3 * bw)
    elif a is not None and d is not None:
        fh = np.asarray(f, float).ravel() * 1e3                            # :40-41
        wrange = (fh[0] - 300, fh[-1] + 300)
    else:
        raise ValueError("Number of input should be either 7 or 9")
    if scale is None or np.size(scale) == 0:
        scale = [0.8, 0.9, 1, 1.1, 1.2]                                    # :47-49
    if nucleus == "H-1":                                                   # :54-61
        gamma = GAMMA_H1
    elif nucleus == "C-13":
        gamma = GAMMA_C13
    else:
        raise ValueError("No such option for nucleus. Options are H-1 and C-13")
    df = np.linspace(wrange[0], wrange[1], 256 * 8)                        # :76-77
    mx, my, mz = bloch_scale_sweep(np.asarray(rf).ravel(), dt * 1e-3, 1e3, 1e3, df, scale, gamma)   # :82-89
    return dict(df=df, scale=np.asarray(scale, float), mxy=mx + 1j * my, mz=mz)


def sim_rf_spectral(rf, dt, rfname=None, nucleus="C-13", ptype="ex", f=None, a=None, d=None, name_cell=None):
    """sim_rf_spectral(rf, dt, rfname, nucleus, ptype, f, a, d, name_cell) — sim_rf_spectral.m without the figures: the spectral
    profile of one pulse (or of a list of pulses: the cell-array form of the reference) over 2048 off-resonances, G = 0,
    T1 = T2 = 1e3 s, M0 = (0, 0, 1) (:62-78).  rf in Gauss, dt in ms.  f alone: passband bandwidth in kHz, range +-3 BW
    (:33-36); f, a, d: the multiband specification in kHz, range f(1) - 500 Hz .. f(end) + 500 Hz (:37-41).
    Returns dict(df [Hz], mxy, mz) -- arrays of 2048 points, or lists of them for a list of pulses."""
    from .bloch import bloch
    if f is None:
        raise ValueError("Number of input should be either 4 or 6")       # :46-48
    if a is None and d is None:
        bw = float(np.ravel(f)[0]) * 1e3
        wrange = (-3 * bw, 3 * bw)
    else:
        fh = np.asarray(f, float).ravel() * 1e3
        wrange = (fh[0] - 500, fh[-1] + 500)
    if nucleus not in ("H-1", "C-13"):
        raise ValueError("No such option for nucleus. Options are H-1 and C-13")
    gamma = GAMMA_H1 if nucleus == "H-1" else GAMMA_C13
    df = np.linspace(wrange[0], wrange[1], 256 * 8)                        # :62-63
    pulses = list(rf) if isinstance(rf, (list, tuple)) else [rf]
    dts = list(dt) if isinstance(dt, (list, tuple)) else [dt] * len(pulses)
    mxy, mzs = [], []
    for p, t in zip(pulses, dts):
        p = np.asarray(p).ravel()
        mx, my, mz = bloch(p, np.zeros(p.size), t * 1e-3, 1e3, 1e3, df, 0.0, 0, gamma=gamma)      # :70-78
        mxy.append((np.asarray(mx) + 1j * np.asarray(my)).ravel())
        mzs.append(np.asarray(mz).ravel())
    if isinstance(rf, (list, tuple)):
        return dict(df=df, mxy=mxy, mz=mzs)
    return dict(df=df, mxy=mxy[0], mz=mzs[0])
