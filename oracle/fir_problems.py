"""TEST INFRASTRUCTURE ONLY — numpy restatement of how the reference BUILDS its convex FIR design
problems (the solve itself happens in third-party code that is not under /root/reference:
CVX 2.0 beta -> SeDuMi/SDPT3, MATLAB linprog).

  build_fir_ap   : fir_ap_cvx.m:44-142 (grid, bands, A, bounds, stop rows, peak cones)
  build_fir_lp   : ss/fir_linprog.m:46-240 (real / complex-Hermitian linear-phase LP)
  build_fir_qp   : fir_qp_cvx.m:34-139

Parity status: PARITY UNPINNED by the reference (no solver, no golden vectors).
What is pinned: the known-answer of the cone-free fir_ap_cvx LP at N=256 on the dual-band H-1 spec
(15 490 rows x 513 vars, objective 0.0160332 with HiGHS) — checked in tests/test_oracle_fir.py.
"""
from __future__ import annotations

import numpy as np

# dual-band H-1 saturation spec after shift_f (specsat_H1_dualband.m:5-32 -> dzrf_mb.m:100-147), SURVEY.md 8(d)
H1_DUALBAND = dict(
    f=np.array([-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006]),
    a=np.array([0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886]),
    d=np.array([0.014436, 0.022361, 0.017683]),
)


def _bands(w, f, a, d):
    """fir_ap_cvx.m:51-82 / fir_qp_cvx.m:41-73: band membership, interpolated amplitude, bounds."""
    nband = len(f) // 2
    idx_band, U, L, M, D = [], [], [], [], []
    for b in range(nband):
        lo, hi = f[2 * b], f[2 * b + 1]
        idx = np.nonzero((w >= lo) & (w <= hi))[0]
        idx_band.append(idx)
        if lo == hi:
            amp = np.full(idx.size, a[2 * b])
        else:
            amp = a[2 * b] + (a[2 * b + 1] - a[2 * b]) * ((w[idx] - lo) / (hi - lo))
        U.append(amp + d[b]); L.append(amp - d[b]); M.append(amp); D.append(np.full(idx.size, d[b]))
    idx_band = np.concatenate(idx_band)
    U, L, M, D = map(np.concatenate, (U, L, M, D))
    mask = np.ones(w.size, bool)
    mask[idx_band] = False
    idx_tran = np.nonzero(mask)[0]
    return idx_band, idx_tran, U, L, M, D


def build_fir_ap(n, f, a, d, obj=0.0, peak=1e-3, oversamp=15):
    """fir_ap_cvx.m.  Variables z = [x (2n-1); ripple_stop].  Returns dict with
       w (m,)           reordered grid, band rows first then transition rows (:86-91)
       lo, hi (m,)      L_b <= A x <= U_b                                        (:103-120)
       stop (k,) int    rows with A x <= ripple_stop                               (:125)
       c (2n,)          objective x1 + obj*ripple_stop                             (:163)
       radius (n,)      |x1| <= n*Peak ; ||(x_i, x_{n+i-1})|| <= (n-i+1)*Peak       (:166-168)
    """
    f = np.asarray(f, float) * np.pi                       # :44
    a = np.asarray(a, float); d = np.asarray(d, float)
    m = 2 * n * oversamp                                   # :45-46
    w = np.sort(np.concatenate([np.linspace(-np.pi, np.pi, m), f]))   # :47-48
    idx_band, idx_tran, U, L, _, _ = _bands(w, f, a, d)
    if idx_tran.size:                                      # :67-75
        U_tran = np.full(idx_tran.size, U.max())
        L_tran = np.full(idx_tran.size, min(0.0, L.min()))
    else:
        U_tran = L_tran = np.zeros(0)
    w = np.concatenate([w[idx_band], w[idx_tran]])         # :86-91
    U_b = np.concatenate([U, U_tran]) ** 2                 # :103-106
    L_b = np.concatenate([L, L_tran])
    L_b[L_b < 0] = 0                                       # :110-112
    L_b = L_b ** 2
    L_b[L_b < 1e-20] = 1e-20                               # :115-116  epsilon^2
    stop = np.nonzero(np.sqrt(U_b) < np.sqrt(U_b).min() + 1e-2)[0]   # :125
    c = np.zeros(2 * n)
    c[0] = 1.0
    c[-1] = obj
    radius = (n - np.arange(1, n + 1) + 1) * peak         # :166-168, i = 1..n
    return dict(kind="ap", n=n, w=w, lo=L_b, hi=U_b, stop=stop, c=c, radius=radius, nband_rows=idx_band.size)


def matrix_fir_ap(w, n):
    """A = [1, 2cos(w k), 2sin(w k)], k = 1..n-1   (fir_ap_cvx.m:100)."""
    k = np.arange(1, n)
    wk = np.outer(w, k)
    return np.hstack([np.ones((w.size, 1)), 2 * np.cos(wk), 2 * np.sin(wk)])


def violation_fir_ap(p, z):
    """Max constraint violation of z = [x; ripple_stop] in the problem's own units (absolute)."""
    n = p["n"]
    x, t = z[:2 * n - 1], z[-1]
    S = matrix_fir_ap(p["w"], n) @ x
    v = max(0.0, (S - p["hi"]).max(), (p["lo"] - S).max(), (S[p["stop"]] - t).max())
    nx = np.abs(x[0])
    nr = np.hypot(x[1:n], x[n:2 * n - 1])
    v = max(v, nx - p["radius"][0], (nr - p["radius"][1:]).max())
    return v


def solve_fir_ap_highs(p, cone_sides=0):
    """Independent CPU solve of the fir_ap_cvx problem with HiGHS.
    cone_sides = 0 drops the 2-D peak cones (the cone-free LP of SURVEY.md 8d);
    cone_sides = K > 0 replaces each disk by its circumscribed (outer) K-gon -> a lower bound on the
    optimum; the inscribed K-gon (radius*cos(pi/K)) gives an upper bound (returned as second value).
    """
    from scipy.optimize import linprog
    n = p["n"]
    A = matrix_fir_ap(p["w"], n)
    m = A.shape[0]
    nv = 2 * n
    rows = [np.hstack([A, np.zeros((m, 1))]), np.hstack([-A, np.zeros((m, 1))])]
    rhs = [p["hi"], -p["lo"]]
    S = np.hstack([A[p["stop"]], -np.ones((p["stop"].size, 1))])
    rows.append(S); rhs.append(np.zeros(p["stop"].size))

    def run(scale):
        r2, b2 = list(rows), list(rhs)
        bounds = [(None, None)] * nv
        if cone_sides:
            bounds = list(bounds)
            bounds[0] = (-p["radius"][0], p["radius"][0])
            ang = 2 * np.pi * np.arange(cone_sides) / cone_sides
            for i in range(1, n):
                blk = np.zeros((cone_sides, nv))
                blk[:, i] = np.cos(ang); blk[:, n + i - 1] = np.sin(ang)
                r2.append(blk); b2.append(np.full(cone_sides, p["radius"][i] * scale))
        res = linprog(p["c"], A_ub=np.vstack(r2), b_ub=np.concatenate(b2), bounds=bounds, method="highs")
        return res
    outer = run(1.0)
    if not cone_sides:
        return outer, None
    inner = run(np.cos(np.pi / cone_sides))
    return outer, inner


# --------------------------------------------------------------------------------------------
# ss/fir_linprog.m
# --------------------------------------------------------------------------------------------
def build_fir_lp(n, f, a, d):
    """ss/fir_linprog.m:46-240: returns dict(A (m, nx), lo, hi, c=fmin, meta) or None when the reference
    refuses the spec (even n with amplitude 1 at fs/2, :63-75)."""
    f = np.asarray(f, float) * np.pi
    a = np.asarray(a, float); d = np.asarray(d, float)
    real_filter = f.min() >= 0                               # :48-52
    odd = n % 2 == 1                                         # :56-60
    if not odd and np.any(a[np.abs(f) == np.pi] == 1):       # :63-75
        return None
    nhalf = -(-n // 2)                                       # :79
    w = np.linspace(0, np.pi, 15 * n) if real_filter else np.linspace(-np.pi, np.pi, 30 * n)   # :92-101
    w = np.sort(np.concatenate([w, f]))                      # :107
    idx_band, idx_tran, U, L, _, _ = _bands(w, f, a, d)
    U_tran = np.full(idx_tran.size, U.max())                 # :163-171
    L_tran = np.full(idx_tran.size, min(0.0, L.min()))
    w = np.concatenate([w[idx_band], w[idx_tran]])           # :175-180
    if odd:                                                  # :195-217
        k = np.arange(1, nhalf)
        Acos = np.hstack([np.ones((w.size, 1)), 2 * np.cos(np.outer(w, k))])
        Asin = 2 * np.sin(np.outer(w, k))
    else:
        k = np.arange(0, nhalf) + 0.5
        Acos = 2 * np.cos(np.outer(w, k))
        Asin = 2 * np.sin(np.outer(w, k))
    A = Acos if real_filter else np.hstack([Acos, Asin])
    c = A[idx_band.size:].sum(0)                             # :231
    return dict(kind="lp", n=n, A=A, w=w, lo=np.concatenate([L, L_tran]), hi=np.concatenate([U, U_tran]), c=c,
                real=real_filter, odd=odd, nhalf=nhalf)


def solve_fir_lp_highs(p):
    from scipy.optimize import linprog
    A = p["A"]
    return linprog(p["c"], A_ub=np.vstack([A, -A]), b_ub=np.concatenate([p["hi"], -p["lo"]]),
                   bounds=[(None, None)] * A.shape[1], method="highs")


# --------------------------------------------------------------------------------------------
# fir_qp_cvx.m (variant with scalar obj, fir_qp_cvx.m:145-166)
# --------------------------------------------------------------------------------------------
def build_fir_qp(n, f, a, d, k=100.0, obj=0.0, oversamp=10):
    """fir_qp_cvx.m:34-139.  Variables x = [Re h; Im h] (2n), E_total, Peak.
       minimise E_total + obj*Peak
       s.t. ||A_i x - Hd_i|| <= D_i (band), ||A_i x|| <= 1 + 5*max(d) (transition),
            ||(x_i, x_{n+i})|| <= Peak, ||x|| <= E_total.
    Returns dict(w (m,), center (m,) complex, radius (m,), nband, n, obj)."""
    f = np.asarray(f, float) * np.pi                        # :34
    a = np.asarray(a, float); d = np.asarray(d, float)
    m = n * oversamp                                        # :35-36
    w = np.sort(np.concatenate([np.linspace(-np.pi, np.pi, m), f]))   # :37-38
    idx_band, idx_tran, U, L, M, D = _bands(w, f, a, d)
    wband, wtran = w[idx_band], w[idx_tran]                 # :75-76
    Hd = M * np.exp(1j * (k * wband ** 2 - wband * (n - 1) / 2))      # :113-121
    center = np.concatenate([Hd, np.zeros(wtran.size, complex)])
    radius = np.concatenate([D, np.full(wtran.size, 1 + d.max() * 5)])   # :151,156
    return dict(kind="qp", n=n, w=np.concatenate([wband, wtran]), center=center, radius=radius, nband=wband.size,
                obj=float(obj))


def response_fir_qp(w, n, x):
    """H(w) = sum_k h_k exp(-j w k), h = x[:n] + j x[n:]: rows [cos sin; -sin cos] of fir_qp_cvx.m:96-109."""
    kk = np.arange(n)
    E = np.exp(-1j * np.outer(w, kk))
    return E @ (x[:n] + 1j * x[n:])


def objective_fir_qp(p, x):
    n = p["n"]
    return np.linalg.norm(x) + p["obj"] * np.hypot(x[:n], x[n:]).max()


def violation_fir_qp(p, x):
    H = response_fir_qp(p["w"], p["n"], x)
    return max(0.0, (np.abs(H - p["center"]) - p["radius"]).max())


def objective_fir_qp_minimax(p, x, obj2):
    """fir_qp_cvx.m:170-191 with the epigraph variables at their optimal values:
    delta = max_i |H(w_i) - Hd_i| / D_i over the band rows, E_total = ||x||, Peak = max_i |h_i|."""
    n, nb = p["n"], p["nband"]
    H = response_fir_qp(p["w"][:nb], n, x)
    delta = (np.abs(H - p["center"][:nb]) / p["radius"][:nb]).max()
    return delta + obj2[0] * np.linalg.norm(x) + obj2[1] * np.hypot(x[:n], x[n:]).max()


def violation_fir_qp_minimax(p, x):
    """Only the transition rows are constraints in the minimax form: |H(w_i)| <= 1 + 0.1 (:180)."""
    nb = p["nband"]
    if p["w"].size == nb:
        return 0.0
    H = response_fir_qp(p["w"][nb:], p["n"], x)
    return max(0.0, (np.abs(H) - 1.1).max())


def solve_fir_qp_minimax_reference(p, obj2, maxiter=30000):
    """Independent CPU solve of the minimax form (SciPy trust-constr, smooth squared-norm constraints, epigraph variables
    delta, E_total, Peak).  v = [x (2n), delta, E, P]."""
    from scipy.optimize import NonlinearConstraint, minimize
    n, nb = p["n"], p["nband"]
    kk = np.arange(n)
    Ew = np.exp(-1j * np.outer(p["w"], kk))
    Ar = np.hstack([Ew.real, -Ew.imag])
    Ai = np.hstack([Ew.imag, Ew.real])
    cr, ci, D = p["center"].real[:nb], p["center"].imag[:nb], p["radius"][:nb]
    nt = p["w"].size - nb

    def fun(v):
        return v[2 * n] + obj2[0] * v[2 * n + 1] + obj2[1] * v[2 * n + 2]

    def grad(v):
        g = np.zeros_like(v); g[2 * n] = 1; g[2 * n + 1] = obj2[0]; g[2 * n + 2] = obj2[1]; return g

    def cons(v):
        x, dl, E, P = v[:2 * n], v[2 * n], v[2 * n + 1], v[2 * n + 2]
        hr, hi = Ar @ x, Ai @ x
        band = (D * dl) ** 2 - (hr[:nb] - cr) ** 2 - (hi[:nb] - ci) ** 2
        tran = 1.1 ** 2 - hr[nb:] ** 2 - hi[nb:] ** 2
        return np.concatenate([band, tran, P ** 2 - x[:n] ** 2 - x[n:] ** 2, [E ** 2 - x @ x], [dl, E, P]])

    def jac(v):
        x, dl, E, P = v[:2 * n], v[2 * n], v[2 * n + 1], v[2 * n + 2]
        hr, hi = Ar @ x, Ai @ x
        Jb = np.hstack([-2 * ((hr[:nb] - cr)[:, None] * Ar[:nb] + (hi[:nb] - ci)[:, None] * Ai[:nb]),
                        (2 * D ** 2 * dl)[:, None], np.zeros((nb, 2))])
        Jt = np.hstack([-2 * (hr[nb:, None] * Ar[nb:] + hi[nb:, None] * Ai[nb:]), np.zeros((nt, 3))])
        J2 = np.zeros((n, 2 * n + 3))
        J2[np.arange(n), np.arange(n)] = -2 * x[:n]; J2[np.arange(n), n + np.arange(n)] = -2 * x[n:]; J2[:, 2 * n + 2] = 2 * P
        J3 = np.concatenate([-2 * x, [0, 2 * E, 0]])[None, :]
        J4 = np.zeros((3, 2 * n + 3)); J4[0, 2 * n] = 1; J4[1, 2 * n + 1] = 1; J4[2, 2 * n + 2] = 1
        return np.vstack([Jb, Jt, J2, J3, J4])
    x0 = np.linalg.lstsq(np.vstack([Ar[:nb], Ai[:nb]]), np.concatenate([cr, ci]), rcond=None)[0]
    H0 = response_fir_qp(p["w"][:nb], n, x0)
    v0 = np.concatenate([x0, [(np.abs(H0 - p["center"][:nb]) / D).max() * 1.01 + 1e-3, np.linalg.norm(x0) * 1.01 + 1e-3,
                              np.hypot(x0[:n], x0[n:]).max() * 1.01 + 1e-3]])
    return minimize(fun, v0, jac=grad, method="trust-constr", constraints=[NonlinearConstraint(cons, 0, np.inf, jac=jac)],
                    options=dict(maxiter=maxiter, gtol=1e-10, xtol=1e-12, barrier_tol=1e-12, verbose=0))


def solve_fir_qp_reference(p, x0=None, maxiter=3000):
    """Independent CPU solve (SciPy trust-constr on the smooth squared-norm form with epigraph variables)."""
    from scipy.optimize import NonlinearConstraint, minimize
    n = p["n"]
    kk = np.arange(n)
    Ew = np.exp(-1j * np.outer(p["w"], kk))
    Ar = np.hstack([Ew.real, -Ew.imag])          # Re H = Ar x ; Im H = Ai x   (H = Ew (xr + j xi))
    Ai = np.hstack([Ew.imag, Ew.real])
    cr, ci, R2 = p["center"].real, p["center"].imag, p["radius"] ** 2

    def fun(v):
        return v[2 * n] + p["obj"] * v[2 * n + 1]

    def grad(v):
        g = np.zeros_like(v); g[2 * n] = 1; g[2 * n + 1] = p["obj"]; return g

    def cons(v):
        x, E, P = v[:2 * n], v[2 * n], v[2 * n + 1]
        hr, hi = Ar @ x - cr, Ai @ x - ci
        return np.concatenate([R2 - hr ** 2 - hi ** 2, P ** 2 - x[:n] ** 2 - x[n:] ** 2, [E ** 2 - x @ x], [E, P]])

    def jac(v):
        x, E, P = v[:2 * n], v[2 * n], v[2 * n + 1]
        hr, hi = Ar @ x - cr, Ai @ x - ci
        J1 = np.hstack([-2 * (hr[:, None] * Ar + hi[:, None] * Ai), np.zeros((hr.size, 2))])
        J2 = np.zeros((n, 2 * n + 2))
        J2[np.arange(n), np.arange(n)] = -2 * x[:n]; J2[np.arange(n), n + np.arange(n)] = -2 * x[n:]; J2[:, 2 * n + 1] = 2 * P
        J3 = np.concatenate([-2 * x, [2 * E, 0]])[None, :]
        J4 = np.zeros((2, 2 * n + 2)); J4[0, 2 * n] = 1; J4[1, 2 * n + 1] = 1
        return np.vstack([J1, J2, J3, J4])
    if x0 is None:
        # least-squares fit of the band targets as a start
        x0 = np.linalg.lstsq(np.vstack([Ar[:p["nband"]], Ai[:p["nband"]]]), np.concatenate([cr[:p["nband"]], ci[:p["nband"]]]), rcond=None)[0]
    v0 = np.concatenate([x0, [np.linalg.norm(x0) * 1.01 + 1e-3, np.hypot(x0[:n], x0[n:]).max() * 1.01 + 1e-3]])
    res = minimize(fun, v0, jac=grad, method="trust-constr", constraints=[NonlinearConstraint(cons, 0, np.inf, jac=jac)],
                   options=dict(maxiter=maxiter, gtol=1e-10, xtol=1e-12, barrier_tol=1e-12, verbose=0))
    return res


# --------------------------------------------------------------------------------------------
# spectral factorisation -- fir_ap_cvx.m:253-304 (fftc, fmp2, mag2mp), numpy restatement line by line.
# Test infrastructure: the product computes this on the GPU (csrc/fmp.cu).
# --------------------------------------------------------------------------------------------
def mag2mp_reference(x):
    nn = x.size                                                           # fir_ap_cvx.m:292-303
    xlf = np.fft.fft(np.log(x))                                           # :294-295
    xlfp = np.zeros(nn, complex)
    xlfp[0] = xlf[0]                                                      # :296
    xlfp[1:nn // 2] = 2 * xlf[1:nn // 2]                                  # :297
    xlfp[nn // 2] = xlf[nn // 2]                                          # :298
    return np.exp(np.fft.ifft(xlfp))                                      # :300-301


def fmp2_reference(r):
    """hmp = fmp2(h), fir_ap_cvx.m:262-283."""
    h = np.asarray(r, complex).ravel()
    ln = h.size
    if ln % 2 == 0:
        raise ValueError("filter length must be odd")                     # :265-268
    lp = int(round(8 * np.exp(np.ceil(np.log(ln) / np.log(2)) * np.log(2))))                            # :269
    hp = np.concatenate([np.zeros(int(np.ceil((lp - ln) / 2))), h, np.zeros(int(np.floor((lp - ln) / 2)))])   # :270
    hpf = np.fft.fftshift(np.fft.fft(np.fft.fftshift(hp)))                # fftc, :253-255, :271
    hpfmp = mag2mp_reference(np.sqrt(np.abs(hpf)))                        # :278
    hpmp = np.fft.ifft(np.fft.fftshift(np.conj(hpfmp)))                   # :279
    return hpmp[:(ln + 1) // 2]                                           # :280


def x_to_h_reference(x, n):
    """fir_ap_cvx.m:185-202: autocorrelation coefficients -> two-sided sequence -> minimum-phase taps."""
    r = np.concatenate([[x[0]], x[1:n] + 1j * x[n:2 * n - 1]])            # :185
    r = np.concatenate([np.conj(r[:0:-1]), r])                            # :186
    return fmp2_reference(r)


def peak_lower_bound_fir_qp(p, y_disk):
    """A RIGOROUS lower bound on Peak = max_i ||(x_i, x_{n+i})|| over the feasible set of the scalar-obj fir_qp_cvx problem
    (fir_qp_cvx.m:145-166), from ANY multipliers y_i in R^2 of the disk constraints ||A_i x - Hd_i|| <= D_i (build_fir_qp
    rows), evaluated here on the CPU -- it does not matter where y came from, only how good it is:
        feasible x:  <y_i, A_i x - Hd_i> >= -D_i ||y_i||   =>   <r, x> >= sum_i (<y_i, Hd_i> - D_i ||y_i||),   r = sum_i A_i' y_i
        <r, x> <= (sum_pairs ||r_pair||) * max_pairs ||x_pair||                                  (Cauchy-Schwarz per pair)
        =>  Peak >= sum_i (<y_i, Hd_i> - D_i ||y_i||) / sum_pairs ||(r_k, r_{n+k})||.
    Since E_total >= 0, obj * (this bound) is a lower bound on the optimal value E_total + obj * Peak.  Both signs of y are
    tried (the bound is valid for either)."""
    n, w = p["n"], p["w"]
    y = np.asarray(y_disk, float).reshape(-1, 2)
    k = np.arange(n)
    wk = np.outer(w, k)
    C, S = np.cos(wk), np.sin(wk)
    # A_i = [cos sin; -sin cos] (fir_qp_cvx.m:96-109): A_i' y_i = [cos*y0 - sin*y1 ; sin*y0 + cos*y1]
    r = np.concatenate([C.T @ y[:, 0] - S.T @ y[:, 1], S.T @ y[:, 0] + C.T @ y[:, 1]])
    den = np.hypot(r[:n], r[n:]).sum()
    num = (y[:, 0] * p["center"].real + y[:, 1] * p["center"].imag).sum()
    pen = (p["radius"] * np.hypot(y[:, 0], y[:, 1])).sum()
    return max(num - pen, -num - pen) / den
