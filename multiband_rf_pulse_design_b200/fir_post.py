"""Host-side mirror of the steps around the convex FIR design step.

    h_new = fir_flip_zero(h, dbg)                                   fir_flip_zero.m
    [h, status] = fir_qprog_phs(n, f, a, d, x0, dbg)                ss/fir_qprog_phs.m
    [h, status] = fir_min_order_qprog_phs(n, f, a, d, even_odd, dbg)  ss/fir_min_order_qprog_phs.m

fir_flip_zero: the reference expands up to 2^12 polynomials one after the other (poly() in a MATLAB loop); here every flip
pattern is one CTA of `mbrf_flip_zero_batch` (csrc/flipzero.cu).  roots(h) and the choice of patterns stay on the host, as
they are O(N^3)-once / index work.  fir_qprog_phs: the QP the reference hands to quadprog is posed to the GPU's first-order
solver (csrc/pdhg.cu) through `mbrf_fir_pdhg_solve2`.  No CPU solve exists.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import PdhgBlocks, c_double_p, check, lib

c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)
MAX_PATTERNS = 2 ** 12                                                     # fir_flip_zero.m:54,62


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def flip_patterns(n_z, rng=None):
    """The flip patterns fir_flip_zero.m:45-64 tries, as a [Num x n_z] uint8 array (row = pattern, 1 = flip).

    n_z <= 12: all 2^n_z patterns in the column order of `combination_2power` (:119-138) — column c (0-based) flips
    passband zero r (0-based) iff bit (n_z-1-r) of c is CLEAR.  12 < n_z <= 19: 2^12 distinct patterns drawn without
    replacement and sorted (:55-58); n_z > 19: 2^12 independent uniform patterns (:60-63, `combination_MC`).  The
    reference draws from MATLAB's unseeded global stream, so those two cases are not reproducible there either; here
    `rng` (a numpy Generator or seed) makes them so."""
    if n_z == 0:
        return np.zeros((1, 0), np.uint8)
    if n_z <= 19:
        cols = np.arange(2 ** n_z, dtype=np.int64)
        if n_z > 12:
            cols = np.sort(np.random.default_rng(rng).permutation(2 ** n_z)[:MAX_PATTERNS])
        shifts = n_z - 1 - np.arange(n_z)
        return (1 - ((cols[:, None] >> shifts[None, :]) & 1)).astype(np.uint8)
    return np.round(np.random.default_rng(rng).random((MAX_PATTERNS, n_z))).astype(np.uint8)


def flip_zero_candidates(Z, idx_pb, mask, hsum, want_all=False):
    """All candidates of fir_flip_zero.m:66-93 on the GPU.  Z: zeros (complex [nroots]); idx_pb: indices of the zeros that
    may flip; mask [Num x n_pb]; hsum = sum(h).  Returns dict(best, h_new [N], peak [Num], power [Num][, h_array [Num x N]])."""
    Z = np.asarray(Z, complex).ravel()
    zr = np.ascontiguousarray(Z.real)
    zi = np.ascontiguousarray(Z.imag)
    idx = np.ascontiguousarray(idx_pb, dtype=np.int32)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    num = mask.shape[0]
    if mask.ndim != 2 or mask.shape[1] != idx.size:
        raise ValueError("mask must be [Num x number of passband zeros]")
    N = Z.size + 1
    h_re, h_im = np.empty(N), np.empty(N)
    peak, power = np.empty(num), np.empty(num)
    best = C.c_int(-1)
    all_re = np.empty((num, N)) if want_all else None
    all_im = np.empty((num, N)) if want_all else None
    hs = complex(hsum)
    check(lib().mbrf_flip_zero_batch(_dp(zr), _dp(zi), Z.size, idx.ctypes.data_as(c_int_p), idx.size,
                                     mask.ctypes.data_as(c_ubyte_p), num, hs.real, hs.imag, C.byref(best), _dp(h_re), _dp(h_im),
                                     _dp(peak), _dp(power), _dp(all_re) if want_all else None,
                                     _dp(all_im) if want_all else None))
    out = dict(best=int(best.value), h_new=h_re + 1j * h_im, peak=peak, power=power)
    if want_all:
        out["h_array"] = all_re + 1j * all_im
    return out


def _conjugate_closed(e):
    """poly.m's test for a real result: the roots with positive imaginary part are the conjugates of those with negative."""
    pos = np.sort_complex(e[e.imag > 0])
    neg = np.sort_complex(np.conj(e[e.imag < 0]))
    return pos.size == neg.size and np.array_equal(pos, neg)


def fir_flip_zero(h, dbg=0, rng=None, return_info=False):
    """h_new = fir_flip_zero(h, dbg) — fir_flip_zero.m:1-110.

    Zeros of h outside the annulus 0.99 <= |z| <= 1.01 are the passband zeros (:28); every tried combination of reflecting
    them about the unit circle keeps |H| and changes the phase; the one with the smallest peak max|h| is returned (:96-99).
    Like poly(), the result is real when the chosen zeros are closed under conjugation."""
    h = np.asarray(h).ravel()
    Z = np.roots(h)                                                        # :25
    if Z.size == 0:
        return (h.copy(), dict(best=0, peak=np.abs(h[:1]), power=np.abs(h[:1]) ** 2)) if return_info else h.copy()
    absZ = np.abs(Z)
    idx_pb = np.nonzero((absZ > 1 + 1e-2) | (absZ < 1 - 1e-2))[0]          # :28
    mask = flip_patterns(idx_pb.size, rng)
    res = flip_zero_candidates(Z, idx_pb, mask, h.sum())
    h_new = res["h_new"]
    Zb = Z.copy()
    sel = mask[res["best"]].astype(bool)
    Zp = Z[idx_pb]
    Zb[idx_pb[sel]] = (1.0 / np.abs(Zp[sel])) * np.exp(1j * np.angle(Zp[sel]))
    if _conjugate_closed(Zb) and np.isrealobj(h):
        h_new = h_new.real.copy()
    if dbg >= 1:                                                           # :101-103
        p0, p1 = np.abs(h).max(), np.abs(h_new).max()
        print(f"reduce peak amplitude from {p0:6.4f} to {p1:6.4f} by {(p0 - p1) / p0:6.4f}")
    if return_info:
        return h_new, dict(res, Z=Z, idx_pb=idx_pb, mask=mask)
    return h_new


# --------------------------------------------------------------------------------------------
# ss/fir_qprog_phs.m — minimum-energy FIR with magnitude AND phase bounds per band (QP)
# --------------------------------------------------------------------------------------------
QP_EPS_PR = 8e-7
QP_EPS_DR = 1e-5
QP_EPS_GAP = 2e-5
QP_MAX_ITER = 400000


def assemble_fir_qprog_phs(n, f, ac, dc):
    """ss/fir_qprog_phs.m:48-322 -> rows of `A x <= B` as (w, phase, sign, bound): row = sign * Re(e^{-j(w q + phase)} h),
    h = x(1:n) + j x(n+1:2n), q the tap offsets (:226-230).  The reference's in-phase rows `[real(Wtmp) -imag(Wtmp)]`
    (:250,259,313) are phase phi, its quadrature rows `[imag(Wtmp) real(Wtmp)]` (:264,270) are phase phi + pi/2.
    Returns None where the reference returns 'Failed' up front (:190-201)."""
    f = np.asarray(f, float).ravel()
    ac = np.asarray(ac, complex).ravel()
    dc = np.asarray(dc, complex).ravel()
    nband = f.size // 2                                                    # :51
    if any(ac[2 * b] != ac[2 * b + 1] for b in range(nband)):
        raise ValueError("Does not support sloped bands")                  # :55-59
    a = np.abs(ac[0::2])                                                   # :63-66
    aphs = np.angle(a)                                                     # :67 — taken AFTER abs(): zero, as in the reference
    d, dphs = np.abs(dc), np.angle(dc)                                     # :68-69
    for b in range(nband):                                                 # :76-82
        if (a[b] + d[b]) * (a[b] - d[b]) < 0 and (a[b] != 0 or dphs[b] != 0):
            raise ValueError("Bands straddling 0 must have a = 0, angle(d) = 0")
    err_tol = 0.05                                                         # :87
    for b in range(nband):                                                 # :88-100
        if a[b] != 0 and (a[b] - d[b]) * (1.0 / np.cos(dphs[b]) - 1) >= 2 * d[b]:
            import warnings
            warnings.warn("Reducing phase ripple to that feasible")
            dphs[b] = 0.99 * np.arccos((a[b] - d[b]) / (a[b] + d[b]))
    n_seg = int(np.ceil(2 * np.pi / np.arccos(1 - err_tol)))               # :105
    amax = float(np.max(a + d))
    full_circle = np.arange(n_seg + 1) / n_seg * 2 * np.pi
    tran_pts = [full_circle]
    band_pts = []
    for b in range(nband):                                                 # :108-124
        if a[b] == 0:
            band_pts.append(full_circle)
            continue
        n_phs = int(np.ceil(2 * dphs[b] / np.arccos(1 - err_tol * 2 * d[b])))
        with np.errstate(invalid="ignore", divide="ignore"):               # zero phase ripple: 0/0 = NaN, as in MATLAB (:113)
            band_pts.append((np.arange(n_phs + 1) / n_phs * 2 - 1) * dphs[b] + aphs[b])
        if a[b] + d[b] >= amax * (1 - err_tol):
            tran_pts.append(np.array([aphs[b] - dphs[b], aphs[b] + dphs[b]]))
    tran_pts = np.unique(np.concatenate([np.mod(np.concatenate(tran_pts), 2 * np.pi), [0.0, 2 * np.pi]]))   # :128-129
    fw = f * np.pi                                                         # :178
    odd = (n & 1) == 1
    if not odd and np.any(np.abs(ac[np.abs(fw) == np.pi]) != 0):           # :190-201
        return None
    nhalf = int(np.ceil(n / 2))                                            # :205
    w = np.sort(np.concatenate([np.linspace(-np.pi, np.pi, 2 * 15 * n), fw]))   # :212-222
    q = np.arange(-(nhalf - 1), nhalf, dtype=float) if odd else np.arange(-nhalf, nhalf) + 0.5   # :226-230
    in_band = np.zeros(w.size, bool)
    up, low = [], []                                                       # (w rows, phase, bound) of Au x <= Bu / Al x >= Bl
    for b in range(nband):                                                 # :237-272
        idx = np.nonzero((w >= fw[2 * b]) & (w <= fw[2 * b + 1]))[0]
        in_band[idx] = True
        pts = band_pts[b]
        step = np.angle(np.exp(1j * pts[1]) * np.exp(-1j * pts[0]))        # :244-245
        for k in range(pts.size - 1):                                      # outer polygon of the magnitude bound, :247-253
            up.append((w[idx], pts[k] + step / 2, (a[b] + d[b]) * np.cos(step / 2)))
        if a[b] != 0:
            low.append((w[idx], aphs[b], a[b] - d[b]))                     # inner chord, :257-260
            up.append((w[idx], pts[-1] + np.pi / 2, 0.0))                  # phase <= upper edge, :264-265
            low.append((w[idx], pts[0] + np.pi / 2, 0.0))                  # phase >= lower edge, :269-271
    wt = w[~in_band]                                                       # :276-282
    for k in range(tran_pts.size - 1):                                     # :306-317
        step = tran_pts[k + 1] - tran_pts[k]
        up.append((wt, tran_pts[k] + step / 2, amax * np.cos(step / 2)))
    rows_w = np.concatenate([r[0] for r in up + low])
    rows_phase = np.concatenate([np.full(r[0].size, r[1]) for r in up + low])
    nup = sum(r[0].size for r in up)
    hi = np.full(rows_w.size, np.inf)
    lo = np.full(rows_w.size, -np.inf)
    hi[:nup] = np.concatenate([np.full(r[0].size, r[2]) for r in up])
    if low:
        lo[nup:] = np.concatenate([np.full(r[0].size, r[2]) for r in low])
    return dict(n=n, q=q, w=rows_w, phase=rows_phase, lo=lo, hi=hi, grid=w, amax=amax)


def fir_qprog_phs(n, f, a, d, x0=None, dbg=0, return_info=False, **solver_kw):
    """[h, status] = fir_qprog_phs(n, f, a, d, x0, dbg) — ss/fir_qprog_phs.m:1-400.

    The reference solves  min 1/2 x'x  s.t.  A x <= B  with quadprog (:326-345; x0 is overwritten with [] at :337, so it is
    accepted and unused here too).  Minimising 1/2||x||^2 and minimising ||x|| have the same minimiser, and ||x|| is the
    first-order solver's norm term (the one `fir_qp_cvx` uses for E_total): the rows are one Fourier matrix with a phase per
    row, K[i][j] = {cos, sin}(w_i q_j + phase_i), and the constraints are intervals on K x.  'Solved' needs violation <= 8e-7
    and a relative gap <= 2e-5 on ||x||."""
    n = int(n)
    p = assemble_fir_qprog_phs(n, f, a, d)
    if p is None:
        return (np.zeros(0), "Failed", dict(info=None)) if return_info else (np.zeros(0), "Failed")
    if np.isnan(p["phase"]).any() or np.isnan(p["hi"]).any():             # a pass band without phase ripple: NaN rows (:113)
        return (np.zeros(0), "Failed", dict(info=None)) if return_info else (np.zeros(0), "Failed")
    M, N = p["w"].size, 2 * n
    arr = lambda v: np.ascontiguousarray(v, dtype=np.float64)             # noqa: E731
    col_type = np.concatenate([np.full(n, 1), np.full(n, 2)]).astype(np.int32)
    col_kappa = arr(np.concatenate([p["q"], p["q"]]))
    col_amp = arr(np.ones(N))
    w_row, row_phase = arr(p["w"]), arr(p["phase"])
    lo, hi = arr(p["lo"].reshape(M, 1)), arr(p["hi"].reshape(M, 1))
    # |h_k| <= ||x|| <= sqrt(n) * max|H| bounds every tap; the box only makes the dual bound finite
    fin = np.concatenate([p["hi"][np.isfinite(p["hi"])], p["lo"][np.isfinite(p["lo"])]])
    big = 2.0 * np.sqrt(n) * max(1.0, np.abs(fin).max())
    bl, bu = arr(np.full((N, 1), -big)), arr(np.full((N, 1), big))
    c = np.zeros((N, 1))
    lam = np.array([1.0])
    blocks = PdhgBlocks()
    blocks.norm_coords, blocks.norm_w = N, _dp(lam)
    kw = dict(max_iter=QP_MAX_ITER, check_every=64, eps_pr=QP_EPS_PR, eps_dr=QP_EPS_DR, eps_gap=QP_EPS_GAP)
    kw.update(solver_kw)
    z, info = np.zeros((N, 1)), np.zeros((1, 8))
    # Every grid point is inside a polygon inscribed in the circle |H| = amax (:247-253, :306-317), and the base grid
    # linspace(-pi, pi, 30n) is a DFT grid of 30n - 1 >= n points: Parseval gives ||x|| = ||h|| <= amax for every feasible
    # point.  A dual bound above that certifies infeasibility (status 2 -> 'Failed'), which the order search relies on.
    upper = arr([p["amax"] * (1.0 + 1e-9)])
    check(lib().mbrf_fir_pdhg_solve2(_dp(w_row), _dp(row_phase), None, M, col_type.ctypes.data_as(c_int_p), _dp(col_kappa),
                                     _dp(col_amp), N, 0, None, None, None, None, None, 0, _dp(c), _dp(lo), _dp(hi), _dp(bl),
                                     _dp(bu), None, 1, _dp(upper), C.byref(blocks), int(kw["max_iter"]), int(kw["check_every"]),
                                     float(kw["eps_pr"]), float(kw["eps_dr"]), float(kw["eps_gap"]), _dp(z), _dp(info), None))
    ok = info[0, 0] == 1.0                                                 # exitflag == 1, :388
    x = z[:, 0]
    h = x[:n] + 1j * x[n:] if ok else np.zeros(0)                          # :389
    st = "Solved" if ok else "Failed"
    if return_info:
        return h, st, dict(x=x.copy(), info=info[0].copy(), problem=p)
    return h, st


def fir_min_order_qprog_phs(n, f, a, d, even_odd=0, dbg=0, **solver_kw):
    """[h, status] = fir_min_order_qprog_phs(n, f, a, d, even_odd, dbg) — ss/fir_min_order_qprog_phs.m: the bisection of
    fir_min_order_linprog with fir_qprog_phs probes (:100,142); shorter of odd / even."""
    from .fir import _min_order_search, _retry_kw, UndecidedProbe     # noqa: F401

    def probe(nt, hw):
        h, st, ex = fir_qprog_phs(nt, f, a, d, hw, dbg, return_info=True, **solver_kw)
        if ex.get("info") is not None and int(ex["info"][0]) == 3:         # iteration limit is not infeasibility: retry
            kw = dict(solver_kw, max_iter=4 * int(solver_kw.get("max_iter") or QP_MAX_ITER))
            h, st, ex = fir_qprog_phs(nt, f, a, d, hw, dbg, return_info=True, **kw)
            if int(ex["info"][0]) == 3:
                import warnings
                warnings.warn(f"fir_qprog_phs probe at n = {nt} stayed undecided at the iteration limit", UndecidedProbe,
                              stacklevel=2)
        return h, st
    return _min_order_search(int(n), f, a, d, even_odd, probe, pick_longer=False)
