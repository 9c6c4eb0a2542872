"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's post-design steps and of the one QP it hands to quadprog.

  flip_zero_reference     : fir_flip_zero.m:24-118 (zeros, passband selection, flip patterns, poly / scale / peak loop, argmin)
  combination_2power      : fir_flip_zero.m:119-138 (the recursive pattern table, restated as the same recursion)
  build_fir_qprog_phs     : ss/fir_qprog_phs.m:48-310 (phase-constrained minimum-energy QP: A x <= B, H = I)
  solve_fir_qprog_phs_reference : that QP solved by an independent CPU method (SciPy SLSQP / trust-constr)

Parity status: PARITY UNPINNED by the reference (MATLAB roots/poly/quadprog absent, no golden vectors).
`flip_zero_reference` is pure arithmetic, so what stands in for a golden vector is a property the reference's own comments
state (fir_flip_zero.m:4-6): every candidate has the SAME magnitude response as h — checked in tests/test_oracle_fir.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------------
# fir_flip_zero.m
# --------------------------------------------------------------------------------------------
def combination_2power(n):
    """fir_flip_zero.m:119-138: n x 2^n table, column = one flip pattern (1 flip, 0 keep), by the reference's recursion."""
    if n == 1:
        return np.array([[1, 0]])
    low = combination_2power(n - 1)
    cols = low.shape[1]
    return np.hstack([np.vstack([np.ones((1, cols), int), low]), np.vstack([np.zeros((1, cols), int), low])])


def poly_reference(e, dtype=complex):
    """MATLAB poly() for a vector of roots: c = [1 0 .. 0]; for j: c(2:j+1) = c(2:j+1) - e(j)*c(1:j).
    dtype=np.clongdouble evaluates the same recursion in 80-bit arithmetic (how ill-conditioned was the fp64 one?)."""
    e = np.asarray(e, dtype)
    c = np.zeros(e.size + 1, dtype)
    c[0] = 1.0
    for j in range(e.size):
        c[1:j + 2] = c[1:j + 2] - e[j] * c[0:j + 1]
    return c


def flip_zero_reference(h, Z=None, mask=None, dtype=complex):
    """fir_flip_zero.m:24-99 for N_z <= 12 (or an explicit `mask` [N_z x Num]); returns dict(h_new, best, peak, power,
    h_array [N x Num], Z, idx_pb, mask).  Z may be handed in so that product and oracle factor the same zeros.
    Expanding a polynomial from its zeros in fp64 loses digits (intermediate coefficients grow and cancel) -- in MATLAB as
    here; dtype=np.clongdouble repeats the loop in 80-bit arithmetic so that tests can measure that loss and hold the GPU to
    the accuracy of the reference's own fp64 loop instead of to an arbitrary constant."""
    rdtype = np.longdouble if dtype is not complex else float
    h = np.asarray(h, dtype).ravel()
    if Z is None:
        Z = np.roots(np.asarray(h, complex))                               # :25
    Z = np.asarray(Z, dtype)
    absZ = np.abs(Z)
    idx_pb = np.nonzero((absZ > 1 + 1e-2) | (absZ < 1 - 1e-2))[0]          # :28
    N_z = idx_pb.size
    Z_pb = Z[idx_pb]
    Z_pb_flip = (1.0 / np.abs(Z_pb)) * np.exp(1j * np.angle(Z_pb))         # :33, :112-117
    if mask is None:
        if N_z > 12:
            raise ValueError("the restatement enumerates only the exhaustive case N_z <= 12 (the reference samples at random above)")
        mask = combination_2power(N_z) if N_z else np.zeros((0, 1), int)   # :45-48
    mask = np.asarray(mask)
    Num = mask.shape[1]
    N = Z.size + 1
    h_array = np.zeros((N, Num), dtype)
    power = np.zeros(Num, rdtype)
    peak = np.zeros(Num, rdtype)
    for i in range(Num):                                                   # :66-93
        Z_each = Z.copy()
        Z_each[idx_pb] = Z_pb * (1 - mask[:, i]) + Z_pb_flip * mask[:, i]
        h_each = poly_reference(Z_each, dtype)
        h_each = h_each * h.sum() / h_each.sum()
        h_array[:, i] = h_each
        power[i] = np.sum(np.abs(h_each) ** 2)
        peak[i] = np.max(np.abs(h_each))
    best = int(np.argmin(peak))                                            # :96
    return dict(h_new=h_array[:, best], best=best, peak=peak, power=power, h_array=h_array, Z=Z, idx_pb=idx_pb, mask=mask)


# --------------------------------------------------------------------------------------------
# ss/fir_qprog_phs.m
# --------------------------------------------------------------------------------------------
def build_fir_qprog_phs(n, f, ac, dc):
    """ss/fir_qprog_phs.m:48-310 -> dict(A, B, n, w) of  min 1/2 x'x  s.t.  A x <= B,  x = [real(h); imag(h)].
    Returns None for the case the reference rejects (:190-201).  Raises like the reference's error() calls."""
    f = np.asarray(f, float).ravel()
    ac = np.asarray(ac, complex).ravel()
    dc = np.asarray(dc, complex).ravel()
    nband = f.size // 2                                                    # :51
    for band in range(nband):                                              # :55-59
        if ac[2 * band] != ac[2 * band + 1]:
            raise ValueError("Does not support sloped bands")
    a = ac[0::2]                                                           # :63
    a = np.abs(a)                                                          # :66
    aphs = np.angle(a)                                                     # :67  (angle of the MAGNITUDE: identically 0 in the reference)
    d = np.abs(dc)
    dphs = np.angle(dc)
    for band in range(nband):                                              # :76-82
        if (a[band] + d[band]) * (a[band] - d[band]) < 0:
            if a[band] != 0 or dphs[band] != 0:
                raise ValueError("Bands straddling 0 must have a = 0, angle(d) = 0")
    err_tol = 0.05                                                         # :87
    for band in range(nband):                                              # :88-100
        if a[band] != 0:
            magerr_inner = (a[band] - d[band]) * (1.0 / np.cos(dphs[band]) - 1)
            if magerr_inner >= 2 * d[band]:
                dphs[band] = 0.99 * np.arccos((a[band] - d[band]) / (a[band] + d[band]))
    n_phs_tran = int(np.ceil(2 * np.pi / np.arccos(1 - err_tol)))          # :105
    amax = np.max(a + d)
    phs_tran = list(np.arange(n_phs_tran + 1) / n_phs_tran * 2 * np.pi)
    phs_band = []
    for band in range(nband):                                              # :108-124
        if a[band] == 0:
            phs_band.append(np.arange(n_phs_tran + 1) / n_phs_tran * 2 * np.pi)
        else:
            phs_tol = np.arccos(1 - (err_tol * 2 * d[band]))
            n_phs = int(np.ceil(2 * dphs[band] / phs_tol))
            phs_band.append((np.arange(n_phs + 1) / n_phs * 2 - 1) * dphs[band] + aphs[band])
            if (a[band] + d[band]) >= amax * (1 - err_tol):
                phs_tran += [aphs[band] - dphs[band], aphs[band] + dphs[band]]
    phs_tran = np.mod(np.array(phs_tran), 2 * np.pi)                       # :128-129
    phs_tran = np.unique(np.concatenate([phs_tran, [0, 2 * np.pi]]))
    f = f * np.pi                                                          # :178
    odd_filter = (n & 1) == 1
    if not odd_filter:                                                     # :190-201
        idx = np.nonzero(np.abs(f) == np.pi)[0]
        if np.any(np.abs(ac[idx]) != 0):
            return None
    nhalf = int(np.ceil(n / 2))                                            # :205
    oversamp = 15
    m = 2 * oversamp * n                                                   # :217
    w = np.sort(np.concatenate([np.linspace(-np.pi, np.pi, m), f]))        # :218-222
    q = np.arange(-(nhalf - 1), nhalf) if odd_filter else np.arange(-nhalf, nhalf) + 0.5    # :226-230
    W = np.exp(-1j * np.outer(w, q))
    idx_band = []
    Au, Bu, Al, Bl = [], [], [], []
    inph = lambda Wt: np.hstack([Wt.real, -Wt.imag])                       # noqa: E731  in-phase part of Wtmp*x
    quad = lambda Wt: np.hstack([Wt.imag, Wt.real])                        # noqa: E731  quadrature part
    for band in range(nband):                                              # :237-272
        idx = np.nonzero((w >= f[2 * band]) & (w <= f[2 * band + 1]))[0]
        idx_band.append(idx)
        pb = phs_band[band]
        phs_diff = np.angle(np.exp(1j * pb[1]) * np.exp(-1j * pb[0]))
        a_mid = (a[band] + d[band]) * np.cos(phs_diff / 2)
        for k in range(pb.size - 1):
            phs_mid = pb[k] + phs_diff / 2
            Au.append(inph(W[idx] * np.exp(-1j * phs_mid)))
            Bu.append(np.full(idx.size, a_mid))
        if a[band] != 0:
            Al.append(inph(W[idx] * np.exp(-1j * aphs[band])))
            Bl.append(np.full(idx.size, a[band] - d[band]))
            Au.append(quad(W[idx] * np.exp(-1j * pb[-1])))
            Bu.append(np.zeros(idx.size))
            Al.append(quad(W[idx] * np.exp(-1j * pb[0])))
            Bl.append(np.zeros(idx.size))
    idx_band = np.concatenate(idx_band)
    tmp = np.ones(w.size, bool)                                            # :276-282
    tmp[idx_band] = False
    idx_tran = np.nonzero(tmp)[0]
    for k in range(phs_tran.size - 1):                                     # :306-317
        phs_diff = phs_tran[k + 1] - phs_tran[k]
        phs_mid = phs_tran[k] + phs_diff / 2
        Au.append(inph(W[idx_tran] * np.exp(-1j * phs_mid)))
        Bu.append(np.full(idx_tran.size, amax * np.cos(phs_diff / 2)))
    A = np.vstack(Au + [-v for v in Al])                                   # :321-322
    B = np.concatenate(Bu + [-v for v in Bl])
    return dict(A=A, B=B, n=n, w=w)


def solve_fir_qprog_phs_reference(p, maxiter=500):
    """min 1/2 ||x||^2 s.t. A x <= B (ss/fir_qprog_phs.m:326-345, quadprog with H = I, f = 0) by Hildreth-free means:
    the dual  min_{y>=0} 1/2 ||A'y||^2 + B'y  solved with SciPy L-BFGS-B is ill-conditioned, so use SLSQP on the primal
    from the least-norm point of the most violated rows; small n only."""
    from scipy.optimize import minimize
    A, B = p["A"], p["B"]
    nx = A.shape[1]
    cons = dict(type="ineq", fun=lambda x: B - A @ x, jac=lambda x: -A)
    res = minimize(lambda x: 0.5 * x @ x, np.zeros(nx), jac=lambda x: x, constraints=[cons], method="SLSQP",
                   options=dict(maxiter=maxiter, ftol=1e-14))
    return res
