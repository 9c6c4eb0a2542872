"""TEST INFRASTRUCTURE ONLY — scalar, line-by-line restatement of the reference's specification builders, one band at a time
as the .m files are written (the product's mirror, multiband_rf_pulse_design_b200/spec.py, is vectorised over bands / designs):

  ripple_asin   : rf_ripple_GFA.m:166-260 (rf_ripple_asin) with rf_ripple_FA2Beta (:263-295)
  ripple_quad   : rf_ripple_GFA.m:84-163 (rf_ripple_quad)
  mrange        : rf_Mrange_desired.m
  measured_range: the reference's own numeric self-check, rf_ripple_GFA.m:42-79 (dbg >= 1): the magnetisation reached over
                  the computed |beta| range, to be compared with the desired range

Parity status: no MATLAB here, so these are pinned by (i) that self-check, which the reference itself prints, and (ii) the
dual-band H-1 specification of specsat_H1_dualband.m as probed in SURVEY.md 8(d) (f, a, d to six digits).
"""
from __future__ import annotations

import math


def fa2beta(rfa_r, rfa_l, FA):                                             # rf_ripple_GFA.m:263-295
    if rfa_r > math.pi:
        min_B = min(math.sin(rfa_l / 2), math.sin(rfa_r / 2))
        return (min_B, 1.0), (1 - min_B, 1 - min_B)
    min_B, max_B, mid_B = math.sin(rfa_l / 2), math.sin(rfa_r / 2), math.sin(FA / 2)
    return (min_B, max_B), (abs(min_B - mid_B), abs(max_B - mid_B))


def ripple_asin(FA_deg, ripple_M, ptype):                                  # rf_ripple_GFA.m:166-260
    FA = FA_deg * math.pi / 180
    if ptype == "ex":
        if math.sin(FA) + ripple_M >= 1:
            rfa = math.asin(math.sin(FA) - ripple_M)
            rfa_l, rfa_r = rfa, math.pi - rfa
        elif FA <= math.pi / 2:
            rfa_l, rfa_r = math.asin(math.sin(FA) - ripple_M), math.asin(math.sin(FA) + ripple_M)
        else:
            rfa_r, rfa_l = math.pi - math.asin(math.sin(FA) - ripple_M), math.pi - math.asin(math.sin(FA) + ripple_M)
        return fa2beta(rfa_r, rfa_l, FA)
    if ptype in ("sat", "inv"):
        if math.cos(FA) + ripple_M > 1:
            rfa = math.acos(math.cos(FA) - ripple_M)
            rfa_l, rfa_r = -rfa, rfa
        elif math.cos(FA) - ripple_M < -1:
            rfa = math.acos(math.cos(FA) + ripple_M)
            rfa_l, rfa_r = rfa, 2 * math.pi - rfa
        else:
            rfa_r, rfa_l = math.acos(math.cos(FA) - ripple_M), math.acos(math.cos(FA) + ripple_M)
        return fa2beta(rfa_r, rfa_l, FA)
    if ptype == "se":
        mid_M = math.sin(FA / 2) ** 2
        max_M, min_M = max(0, min(1, mid_M + ripple_M)), max(0, min(1, mid_M - ripple_M))
        rng = (math.sqrt(min_M), math.sqrt(max_M))
        return rng, tuple(abs(v - math.sin(FA / 2)) for v in rng)
    raise ValueError(ptype)


def ripple_quad(FA_deg, ripple_M, ptype):                                  # rf_ripple_GFA.m:84-163
    import numpy as np
    FA = FA_deg * math.pi / 180
    if ptype == "ex":
        if math.sin(FA) + ripple_M >= 1:
            C_r, C_l = [-0.5 * math.sin(FA), math.cos(FA), ripple_M], [-0.5 * math.sin(FA), -math.cos(FA), ripple_M]
        elif FA <= math.pi / 2:
            C_r, C_l = [-0.5 * math.sin(FA), math.cos(FA), -ripple_M], [-0.5 * math.sin(FA), -math.cos(FA), ripple_M]
        else:
            C_r, C_l = [-0.5 * math.sin(FA), math.cos(FA), ripple_M], [-0.5 * math.sin(FA), -math.cos(FA), -ripple_M]
    elif ptype in ("sat", "inv"):
        if math.cos(FA) + ripple_M > 1:
            C_r, C_l = [-0.5 * math.cos(FA), math.sin(FA), ripple_M], [-0.5 * math.cos(FA), -math.sin(FA), ripple_M]
        elif math.cos(FA) - ripple_M < -1:
            C_r, C_l = [-0.5 * math.cos(FA), math.sin(FA), -ripple_M], [-0.5 * math.cos(FA), -math.sin(FA), -ripple_M]
        else:
            C_r, C_l = [-0.5 * math.cos(FA), math.sin(FA), -ripple_M], [-0.5 * math.cos(FA), -math.sin(FA), ripple_M]
    else:
        raise ValueError(ptype)
    pos = lambda C: min(r.real for r in np.roots(C) if abs(r.imag) == 0 and r.real > 0)   # noqa: E731
    return fa2beta(FA + pos(C_r), FA - pos(C_l), FA)


def mrange(FA_deg, ripple_M, ptype):                                       # rf_Mrange_desired.m
    FA = FA_deg * math.pi / 180
    if ptype in ("st", "ex"):
        return (math.sin(FA) - ripple_M, min(math.sin(FA) + ripple_M, 1))
    if ptype in ("sat", "inv"):
        return (max(math.cos(FA) - ripple_M, -1), min(math.cos(FA) + ripple_M, 1))
    m = math.sin(FA / 2) ** 2
    return (max(m - ripple_M, 0), min(m + ripple_M, 1))


def measured_range(range_B, ptype, N0=1000):
    """rf_ripple_GFA.m:42-79: sample |beta| over range_B and return the magnetisation range reached."""
    import numpy as np
    B = np.linspace(range_B[0], range_B[1], N0)
    if ptype == "ex":
        M = 2 * np.sqrt(1 - B ** 2) * B                                    # :52
    elif ptype in ("sat", "inv"):
        M = 1 - 2 * B ** 2                                                 # :61
    elif ptype == "se":
        M = B ** 2
    else:
        M = B
    return float(M.min()), float(M.max())
