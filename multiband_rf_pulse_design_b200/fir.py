"""Host-side mirror of the reference's convex FIR design step.

    [h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg)            fir_ap_cvx.m
    [h, status, n_op, f_op] = fir_ap(n, f, a, d, Peak, min_order, min_tran, min_peak, dbg)   fir_ap.m
    [h, status] = fir_linprog(n, f, a, d, h0, dbg)                  ss/fir_linprog.m
    [h, status] = fir_min_order_linprog(n, f, a, d, even_odd, dbg)  ss/fir_min_order_linprog.m
    [h, status] = fir_min_order(n, f, a, d, even_odd, a_min, dbg)   ss/fir_min_order.m (LP feasibility form)

Same names, positional arguments, defaults and status strings ('Solved' / 'Failed', h = [] on failure).
The reference hands the problem to CVX (SeDuMi/SDPT3) or MATLAB linprog; here the identical problem is
assembled on the host exactly as the .m file does (grid, band masks, bounds — O(m) scalar work) and
solved on the GPU by libmbrf's batched restarted PDHG (`mbrf_fir_pdhg_solve`): the Fourier matrix is
generated on the device, several designs that share n are solved as one batch.  No CPU solve exists.

`fir_ap_cvx_batch` is the batched entry the sweeps use (fir_ap.m's bisections, trade-off sweeps).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import PdhgBlocks, c_double_p, check, lib

c_int_p = C.POINTER(C.c_int)

# solver defaults: tolerances of BASELINE.json's north star with a safety margin
EPS_PR = 8e-7      # max constraint violation (absolute, in |H|^2 units)      (<= 1e-6 required)
EPS_DR = 1e-4      # natural residual ||z - P_X(z - c - K^T y)||_inf in column-scaled units (PDLP-style 1e-4 class
                   # dual tolerance; the objective and the violation are what the north star bounds)
EPS_GAP = 5e-5     # |primal - dual| / |primal|                              (<= 1e-4 required)
MAX_ITER = 200000
CHECK_EVERY = 64
# interior-point solver (csrc/ipm.cu): relative residuals, relative gap, absolute gap, iteration limit
IPM_FEASTOL = 1e-7
IPM_RELTOL = 2e-6
IPM_ABSTOL = 1e-12
IPM_MAX_ITER = 100
# which solver the mirrors use: "ipm" (second order, meets the 1e-4 objective tolerance at the reference's own weights
# obj = 1e4 / 1e5) or "pdhg" (first order); MBRF_FIR_METHOD overrides
import os as _os
DEFAULT_METHOD = _os.environ.get("MBRF_FIR_METHOD", "ipm")
# where fir_ap_cvx problems are assembled for the interior-point solver: "device" (mbrf_fir_ap_solve: specification in, taps
# out) or "host" (numpy below + mbrf_fir_ipm_solve); the first-order solver always takes the host path
DEFAULT_ASSEMBLE = _os.environ.get("MBRF_FIR_ASSEMBLE", "device")


_OPTION_INDEX = {"eta_factor": 0, "beta_sufficient": 1, "beta_necessary": 2, "beta_artificial": 3, "omega_smoothing": 4,
                 "min_restart_interval": 5}


def set_solver_options(halpern=None, gemm=None, tc_digits=None, **restart):
    """Process-wide knobs of the PDHG solver (include/mbrf.h: mbrf_pdhg_set_option / _set_halpern / _set_gemm /
    _set_tc_digits); host state only, no device needed.  restart: eta_factor (step = eta_factor / ||K||), the three constants
    of the PDLP restart test, omega_smoothing (primal-weight update), min_restart_interval (iterations)."""
    for name, value in restart.items():
        if name not in _OPTION_INDEX:
            raise TypeError(f"unknown solver option {name!r}; known: {sorted(_OPTION_INDEX)}")
        if lib().mbrf_pdhg_set_option(_OPTION_INDEX[name], float(value)) != 0:
            raise ValueError(f"solver option {name} must be positive, got {value!r}")
    for fn, value in (("mbrf_pdhg_set_halpern", halpern), ("mbrf_pdhg_set_gemm", gemm), ("mbrf_pdhg_set_tc_digits", tc_digits)):
        if value is not None and getattr(lib(), fn)(int(value)) != 0:
            raise ValueError(f"{fn}: value {value!r} out of range")


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


# --------------------------------------------------------------------------------------------
# problem assembly — fir_ap_cvx.m:44-142
# --------------------------------------------------------------------------------------------
def _bands(w, f, a, d):
    """fir_ap_cvx.m:51-82: band membership, linearly interpolated amplitude +- ripple, transition rows."""
    nband = len(f) // 2
    idx_band, U, L = [], [], []
    for b in range(nband):
        lo, hi = f[2 * b], f[2 * b + 1]
        idx = np.nonzero((w >= lo) & (w <= hi))[0]                       # :54
        idx_band.append(idx)
        if lo == hi:
            amp = np.full(idx.size, a[2 * b])                             # :57-58
        else:
            amp = a[2 * b] + (a[2 * b + 1] - a[2 * b]) * ((w[idx] - lo) / (hi - lo))   # :60
        U.append(amp + d[b])
        L.append(amp - d[b])
    idx_band = np.concatenate(idx_band) if idx_band else np.zeros(0, int)
    U = np.concatenate(U) if U else np.zeros(0)
    L = np.concatenate(L) if L else np.zeros(0)
    mask = np.ones(w.size, bool)
    mask[idx_band] = False
    return idx_band, np.nonzero(mask)[0], U, L


_AP_ROWS_CACHE: dict = {}


def _ap_rows(n, f, a, d, oversamp):
    """Grid rows and bounds of one band specification (everything in assemble_fir_ap that does not depend on obj / Peak).
    Cached: the designs of a trade-off sweep share it, and identical array objects let the batch code fill in blocks."""
    key = (n, f.tobytes(), a.tobytes(), d.tobytes(), oversamp)
    hit = _AP_ROWS_CACHE.get(key)
    if hit is not None:
        return hit
    m = 2 * n * oversamp                                                  # :45-46
    base = np.linspace(-np.pi, np.pi, m)
    w = np.sort(np.concatenate([base, f]))                                # :47-48
    idx_band, idx_tran, U, L = _bands(w, f, a, d)
    if idx_tran.size:                                                     # :67-75
        U_tran = np.full(idx_tran.size, U.max())
        L_tran = np.full(idx_tran.size, min(0.0, L.min()))
    else:
        U_tran = L_tran = np.zeros(0)
    w = np.concatenate([w[idx_band], w[idx_tran]])                        # :86-91
    U_b = np.concatenate([U, U_tran]) ** 2                                # :103-106
    L_b = np.concatenate([L, L_tran])
    L_b[L_b < 0] = 0                                                      # :110-112
    L_b = L_b ** 2
    L_b[L_b < 1e-20] = 1e-20                                              # :115-116 (epsilon^2)
    stop = np.sqrt(U_b) < np.sqrt(U_b).min() + 1e-2                       # :125
    for v in (w, L_b, U_b, stop):
        v.setflags(write=False)
    if len(_AP_ROWS_CACHE) > 64:
        _AP_ROWS_CACHE.clear()
    _AP_ROWS_CACHE[key] = (w, L_b, U_b, stop)
    return _AP_ROWS_CACHE[key]


def assemble_fir_ap(n, f, a, d, obj, peak, oversamp=15):
    """One design -> (w rows, lo, hi, stop mask, objective weight, radii) in the reference's row order."""
    f = np.asarray(f, float).ravel() * np.pi                              # :44
    a = np.asarray(a, float).ravel()
    d = np.asarray(d, float).ravel()
    w, L_b, U_b, stop = _ap_rows(int(n), f, a, d, int(oversamp))
    radius = (n - np.arange(1, n + 1) + 1) * float(peak)                  # :166-168
    return dict(n=n, w=w, lo=L_b, hi=U_b, stop=stop, obj=float(obj), radius=radius)


def _assemble_batch_ap(n, designs):
    """Host assembly of a batch (assemble_fir_ap dicts of one n) into the arrays of the C ABI: one matrix shared through the
    union of the designs' grids, [dim x B] per-design arrays.  Returns a dict (w_row, M, M1, srows, N, c, lo, hi, bl, bu, rho,
    upper, sw, col_type, col_kappa, col_amp, pair_i, pair_j).  `mbrf_fir_ap_assemble` / `mbrf_fir_ap_solve` do the same on
    the device (bit-identical, tests/test_fir_assemble_gpu.py); this version feeds the first-order solver."""
    B = len(designs)
    # designs of a sweep mostly share their grid (only band-edge samples differ): deduplicate before the union
    keys = [id(p["w"]) for p in designs]           # assemble_fir_ap hands out cached (shared) arrays per band specification
    distinct = {}
    for k, p in zip(keys, designs):
        distinct.setdefault(k, p["w"])
    allw = np.unique(np.concatenate(list(distinct.values())))
    M1 = allw.size
    pos_of = {}
    for k, w in distinct.items():
        ix = np.searchsorted(allw, w)
        sx = np.sort(ix)
        pos_of[k] = (ix, bool((sx[1:] == sx[:-1]).any()))
    pos = [pos_of[k][0] for k in keys]
    has_dup = [pos_of[k][1] for k in keys]
    stop_any = np.zeros(M1, bool)
    for p, ix in zip(designs, pos):
        stop_any[ix[p["stop"]]] = True
    srows = np.nonzero(stop_any)[0]
    srank = -np.ones(M1, int)
    srank[srows] = np.arange(srows.size)
    M = M1 + srows.size
    nx = 2 * n - 1
    N = nx
    w_row = np.concatenate([allw, allw[srows]])
    # per-design arrays are filled design-major (contiguous rows) and transposed once at the end: [dim x B] is what the
    # C ABI takes, but column-strided writes and ufunc.at made this loop the largest host cost of a 512-design batch
    loT = np.empty((B, M))
    hiT = np.empty((B, M))
    c = np.zeros((N, B))
    bl = np.full((N, B), -np.inf)
    bu = np.full((N, B), np.inf)
    rho = np.zeros((n - 1, B))
    upper = np.zeros(B)
    sw = np.zeros(B)
    groups = {}                                     # designs sharing rows and bounds are filled as one block
    for b, p in enumerate(designs):
        groups.setdefault((keys[b], id(p["lo"]), id(p["hi"]), id(p["stop"])), []).append(b)
    for members in groups.values():
        p, ix = designs[members[0]], pos[members[0]]
        row_lo = np.full(M, -np.inf)
        row_hi = np.full(M, np.inf)
        # a grid point may occur twice in a design (band edge coinciding with a base sample): keep the tighter
        if not has_dup[members[0]]:
            row_lo[ix] = p["lo"]
            row_hi[ix] = p["hi"]
        else:
            np.maximum.at(row_lo, ix, p["lo"])
            np.minimum.at(row_hi, ix, p["hi"])
        # `A_U(idx_stop,:)*x <= ripple_stop` with `obj*ripple_stop` in the objective (:163-165)
        #   ==  obj * max_{i in idx_stop} (A x)_i : the duplicate rows form the solver's simplex block
        row_hi[M1 + srank[ix[p["stop"]]]] = 0.0                           # membership flag of the block
        loT[members] = row_lo
        hiT[members] = row_hi
        stop_hi_max = p["hi"][p["stop"]].max()
        for b in members:
            q = designs[b]
            sw[b] = q["obj"]
            bl[0, b], bu[0, b] = -q["radius"][0], q["radius"][0]          # |x1| <= n Peak, :167 (i = 1)
            rho[:, b] = q["radius"][1:]
            upper[b] = q["radius"][0] + q["obj"] * stop_hi_max            # no feasible point has a larger objective
    c[0, :] = 1.0                                                         # minimise x(1) + ..., :163
    lo = np.ascontiguousarray(loT.T)
    hi = np.ascontiguousarray(hiT.T)
    col_type = np.concatenate([[0], np.full(n - 1, 1), np.full(n - 1, 2)]).astype(np.int32)
    k = np.arange(1, n, dtype=float)
    col_kappa = np.concatenate([[0.0], k, k])
    col_amp = np.concatenate([[1.0], np.full(2 * n - 2, 2.0)])            # A = [1, 2cos, 2sin], :100
    pair_i = np.arange(1, n, dtype=np.int32)                              # (x_i, x_{n+i-1}), :133-139
    pair_j = np.arange(n, 2 * n - 1, dtype=np.int32)
    arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in (w_row, col_kappa, col_amp, c, lo, hi, bl, bu, rho, upper, sw)]
    w_row, col_kappa, col_amp, c, lo, hi, bl, bu, rho, upper, sw = arrs
    return dict(w_row=w_row, M=M, M1=M1, srows=srows, N=N, c=c, lo=lo, hi=hi, bl=bl, bu=bu, rho=rho, upper=upper, sw=sw,
                col_type=col_type, col_kappa=col_kappa, col_amp=col_amp, pair_i=pair_i, pair_j=pair_j)


def _solve_batch_ap(n, designs, max_iter=None, check_every=CHECK_EVERY, eps_pr=EPS_PR, eps_dr=EPS_DR,
                    eps_gap=EPS_GAP, warm=None, want_dual=False, method=None, ipm_max_iter=None):
    """Solve designs (assemble_fir_ap dicts with one n) as ONE batch sharing one matrix.

    Rows of the shared matrix = union over the batch of the designs' grid points (the base grid is common,
    band-edge samples differ), followed by one duplicate of every row that is a stop-band row of at least
    one design (the `A_U(idx_stop,:)*x <= ripple_stop` block, fir_ap_cvx.m:165).  A row a design does not
    own gets the bounds (-inf, +inf) for that design.
    warm = (x0 [B, 2n-1], y0 [B, M], omega0 [B]) starts the iteration from a neighbouring design's solution (entries may
    be None); want_dual=True appends (y [B, M], omega [B]) to the result, in the row order of THIS batch's matrix (valid as a
    warm start for batches with the same grids and stop rows, e.g. the other designs of an obj x Peak sweep).
    Returns x [B, 2n-1], ripple_stop [B], info [B, 8].
    """
    B = len(designs)
    q = _assemble_batch_ap(n, designs)
    M, M1, srows, N = q["M"], q["M1"], q["srows"], q["N"]
    w_row, col_kappa, col_amp, c, lo, hi, bl, bu, rho, upper, sw = (q[k] for k in ("w_row", "col_kappa", "col_amp", "c", "lo", "hi",
                                                                                      "bl", "bu", "rho", "upper", "sw"))
    col_type, pair_i, pair_j = q["col_type"], q["pair_i"], q["pair_j"]
    z = np.zeros((N, B))
    info = np.zeros((B, 8))
    method = method or DEFAULT_METHOD
    if method == "ipm":
        if warm is not None or want_dual:
            raise ValueError("warm starts belong to the first-order solver (method='pdhg')")
        check(lib().mbrf_fir_ipm_solve(_dp(w_row), M, _ip(col_type), _dp(col_kappa), _dp(col_amp), N, _ip(pair_i), _ip(pair_j),
                                       n - 1, _dp(c), _dp(lo), _dp(hi), _dp(bl), _dp(bu), _dp(rho), B, M1, int(srows.size),
                                       _dp(sw), int(ipm_max_iter or IPM_MAX_ITER), IPM_FEASTOL, IPM_RELTOL, IPM_ABSTOL, _dp(z),
                                       _dp(info)))
        return z.T.copy(), info[:, 7].copy(), info
    if method != "pdhg":
        raise ValueError(f"unknown method {method!r}: 'ipm' or 'pdhg'")
    max_iter = max_iter or MAX_ITER
    keep = []                                   # host arrays the C side reads / writes during the solve
    if warm is not None or want_dual:
        x0, y0, om0 = warm if warm is not None else (None, None, None)
        zi = np.ascontiguousarray(np.asarray(x0, float).T) if x0 is not None else None          # [N x B]
        yi = np.ascontiguousarray(np.asarray(y0, float).T) if y0 is not None else None          # [M x B]
        oi = np.ascontiguousarray(om0, dtype=float) if om0 is not None else None
        if yi is not None and yi.shape[0] != M:
            yi = None                             # multipliers of a batch with other rows (different band edges): x only
        if (zi is not None and zi.shape != (N, B)) or (yi is not None and yi.shape != (M, B)) or (oi is not None and oi.shape != (B,)):
            raise ValueError("warm start arrays do not match this batch (x0 [B, 2n-1], y0 [B, M], omega0 [B])")
        yo = np.zeros((M, B)) if want_dual else None
        oo = np.ones(B) if want_dual else None
        keep = [zi, yi, oi, yo, oo]
        check(lib().mbrf_fir_pdhg_warm_start(*[(_dp(v) if v is not None else None) for v in keep]))
    check(lib().mbrf_fir_pdhg_solve(_dp(w_row), None, M, _ip(col_type), _dp(col_kappa), _dp(col_amp), N, -1,
                                    _ip(pair_i), _ip(pair_j), n - 1, _dp(c), _dp(lo), _dp(hi), _dp(bl), _dp(bu),
                                    _dp(rho), B, _dp(upper), M1, int(srows.size), _dp(sw), int(max_iter),
                                    int(check_every), float(eps_pr), float(eps_dr), float(eps_gap), _dp(z), _dp(info),
                                    None))
    if want_dual:
        return z.T.copy(), info[:, 7].copy(), info, keep[3].T.copy(), keep[4].copy()
    return z.T.copy(), info[:, 7].copy(), info


def _solve_batch_ap_device(n, f_list, a_list, d_list, obj_list, peak_list, ipm_max_iter=None, want_h=True, oversamp=15):
    """fir_ap_cvx for B designs of one order through ONE C call (`mbrf_fir_ap_solve`): per-design assembly on the device,
    interior-point solve, spectral factors on the device.  Returns x [B, 2n-1], h [B, n] complex (or None), info [B, 8],
    (grid rows, stop rows)."""
    B = len(f_list)
    f = np.ascontiguousarray(np.stack([np.asarray(v, float).ravel() for v in f_list]))
    a = np.ascontiguousarray(np.stack([np.asarray(v, float).ravel() for v in a_list]))
    d = np.ascontiguousarray(np.stack([np.asarray(v, float).ravel() for v in d_list]))
    nband = d.shape[1]
    if f.shape != (B, 2 * nband) or a.shape != (B, 2 * nband):
        raise ValueError("f and a need two entries per band, d one")
    obj = np.ascontiguousarray(obj_list, dtype=float)
    peak = np.ascontiguousarray(peak_list, dtype=float)
    x = np.empty((B, 2 * n - 1))
    info = np.zeros((B, 8))
    hr = np.empty((B, n)) if want_h else None
    hi = np.empty((B, n)) if want_h else None
    rows = (C.c_int * 2)()
    check(lib().mbrf_fir_ap_solve(int(n), nband, _dp(f), _dp(a), _dp(d), _dp(obj), _dp(peak), B, int(oversamp),
                                  int(ipm_max_iter or IPM_MAX_ITER), IPM_FEASTOL, IPM_RELTOL, IPM_ABSTOL, _dp(x),
                                  _dp(hr) if want_h else None, _dp(hi) if want_h else None, _dp(info), rows))
    return x, (hr + 1j * hi) if want_h else None, info, (rows[0], rows[1])


def assemble_fir_ap_device(n, f_list, a_list, d_list, obj_list, peak_list, oversamp=15):
    """The arrays `mbrf_fir_ap_solve` hands to its solver, assembled on the device (`mbrf_fir_ap_assemble`), in the layout of
    _assemble_batch_ap: dict(w_row, M, M1, ns, lo, hi, c, bl, bu, rho, sw)."""
    B = len(f_list)
    f = np.ascontiguousarray(np.stack([np.asarray(v, float).ravel() for v in f_list]))
    a = np.ascontiguousarray(np.stack([np.asarray(v, float).ravel() for v in a_list]))
    d = np.ascontiguousarray(np.stack([np.asarray(v, float).ravel() for v in d_list]))
    nband = d.shape[1]
    obj = np.ascontiguousarray(obj_list, dtype=float)
    peak = np.ascontiguousarray(peak_list, dtype=float)
    rows = (C.c_int * 2)()
    args = (int(n), nband, _dp(f), _dp(a), _dp(d), _dp(obj), _dp(peak), B, int(oversamp), rows)
    check(lib().mbrf_fir_ap_assemble(*args, None, None, None, None, None, None, None, None))
    M1, ns = rows[0], rows[1]
    M, N = M1 + ns, 2 * n - 1
    out = dict(w_row=np.empty(M), lo=np.empty((M, B)), hi=np.empty((M, B)), c=np.empty((N, B)), bl=np.empty((N, B)),
               bu=np.empty((N, B)), rho=np.empty((n - 1, B)), sw=np.empty(B))
    check(lib().mbrf_fir_ap_assemble(*args, *[_dp(out[k]) for k in ("w_row", "lo", "hi", "c", "bl", "bu", "rho", "sw")]))
    return dict(out, M=M, M1=M1, ns=ns)


# --------------------------------------------------------------------------------------------
# spectral factorisation — fir_ap_cvx.m:185-202, 253-304, batched on the GPU (csrc/fmp.cu)
# --------------------------------------------------------------------------------------------
def fmp2_batch(r):
    """Minimum-phase spectral factors of B autocorrelation sequences r [B, 2n-1] (complex or real) -> [B, n] complex:
    fmp2 / mag2mp / fftc of fir_ap_cvx.m:253-304, one CTA per sequence (libmbrf `mbrf_fmp2_batch`)."""
    r = np.atleast_2d(np.asarray(r))
    B, ln = r.shape
    if ln % 2 == 0:
        raise ValueError("filter length must be odd")                     # :265-268
    n = (ln + 1) // 2
    rr = np.ascontiguousarray(r.real, dtype=np.float64)
    ri = np.ascontiguousarray(r.imag, dtype=np.float64) if np.iscomplexobj(r) else None
    hr = np.empty((B, n)); hi = np.empty((B, n))
    check(lib().mbrf_fmp2_batch(_dp(rr), _dp(ri) if ri is not None else None, n, B, _dp(hr), _dp(hi)))
    return hr + 1j * hi


def fmp2(r):
    """Minimum-phase spectral factor of the autocorrelation r (length 2n-1) — fir_ap_cvx.m:262-283."""
    return fmp2_batch(np.asarray(r).ravel()[None, :])[0]


def _x_to_r(x, n):
    r = np.concatenate([[x[0]], x[1:n] + 1j * x[n:2 * n - 1]])            # :185
    return np.concatenate([np.conj(r[:0:-1]), r])                         # :186


def _x_to_h(x, n):
    return fmp2(_x_to_r(x, n))


def _solve_concurrently(jobs):
    """Run independent solves (different orders n => different matrices, so they cannot share a batch) from
    concurrent host threads: every thread owns a CUDA stream and scratch inside libmbrf, ctypes drops the GIL, and
    the single-design kernels are small enough to overlap on the GPU.  jobs: list of zero-argument callables."""
    if len(jobs) <= 1:
        return [j() for j in jobs]
    from concurrent.futures import ThreadPoolExecutor
    import ctypes as _C
    dev = _C.c_int(0)
    have_dev = lib().mbrf_get_device(_C.byref(dev)) == 0   # no device: the jobs themselves fail loudly when they compute

    def run(j):
        if have_dev:                                   # worker threads start on device 0: hand them the caller's device
            check(lib().mbrf_set_device(dev.value))
        return j()
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        return list(ex.map(run, jobs))


def _speculate(state, step, depth):
    """All probes a bisection can ask for within `depth` steps, whatever the outcomes.
    state: hashable search state; step(state, solved) -> next state; the probe of a state is state[-1]."""
    todo, seen, frontier = [], set(), [state]
    for _ in range(depth):
        nxt = []
        for st in frontier:
            if st is None or st[-1] is None:
                continue
            if st[-1] not in seen:
                seen.add(st[-1])
                todo.append(st[-1])
            nxt += [step(st, True), step(st, False)]
        frontier = nxt
    return todo


# --------------------------------------------------------------------------------------------
# public mirrors
# --------------------------------------------------------------------------------------------
class UndecidedProbe(UserWarning):
    """A feasibility probe of a bisection ended at the solver's iteration limit twice: neither solved nor certified
    infeasible.  The search treats it as 'Failed' (as the reference treats every non-'Solved' CVX status, fir_ap.m:86,149)
    but says so, because a wrong 'Failed' moves the bisection to a longer filter / narrower band than necessary."""


def _retry_kw(solver_kw):
    """More effort for a probe that hit the iteration limit (status 3): three times the iterations."""
    kw = dict(solver_kw)
    if (kw.get("method") or DEFAULT_METHOD) == "ipm":
        kw["ipm_max_iter"] = 3 * int(kw.get("ipm_max_iter") or IPM_MAX_ITER)
    else:
        kw["max_iter"] = 4 * int(kw.get("max_iter") or MAX_ITER)
    return kw


def fir_ap_cvx_decided(n, f_list, a, d, obj_list, peak_list, **solver_kw):
    """fir_ap_cvx_batch for the bisections: designs that end at the iteration limit (status 3 -- NOT a certificate of
    infeasibility, unlike status 2) are solved again with more iterations before their 'Failed' is believed; if the limit is
    hit again an UndecidedProbe warning is raised.  Returns (h_list, status_list, code_list)."""
    import warnings
    hs, st, ex = fir_ap_cvx_batch(n, f_list, a, d, obj_list, peak_list, return_info=True, **solver_kw)
    code = ex["info"][:, 0].astype(int)
    again = np.nonzero(code == 3)[0]
    if again.size:
        pick = lambda v: [v[i] for i in again] if isinstance(v, (list, tuple)) and np.ndim(v[0]) else v   # noqa: E731
        h2, s2, e2 = fir_ap_cvx_batch(n, [f_list[i] for i in again], pick(a), pick(d), [obj_list[i] for i in again],
                                      [peak_list[i] for i in again], return_info=True, **_retry_kw(solver_kw))
        for k, i in enumerate(again):
            hs[i], st[i], code[i] = h2[k], s2[k], int(e2["info"][k, 0])
        if (code[again] == 3).any():
            warnings.warn(f"{int((code[again] == 3).sum())} feasibility probe(s) at n = {n} stayed undecided at the iteration "
                          "limit and are treated as 'Failed'", UndecidedProbe, stacklevel=2)
    return hs, st, list(code)


def fir_ap_cvx_batch(n, f_list, a, d, obj_list, peak_list, return_info=False, **solver_kw):
    """Batched fir_ap_cvx: designs i = 0..B-1 with band edges f_list[i], trade-off obj_list[i], Peak peak_list[i]
    (a, d shared or per-design lists).  Returns (h_list, status_list[, info]); h is None where 'Failed'.
    info["info"][:, 0] keeps what the string cannot: 1 solved, 2 infeasible (certificate), 3 iteration limit (undecided).
    With the interior-point solver the whole batch is ONE C call (assemble="device", the default: specification in, taps out);
    assemble="host" builds the problem in numpy and hands the arrays over."""
    B = len(f_list)
    a_list = a if isinstance(a, (list, tuple)) and np.ndim(a[0]) else [a] * B
    d_list = d if isinstance(d, (list, tuple)) and np.ndim(d[0]) else [d] * B
    assemble = solver_kw.pop("assemble", None) or DEFAULT_ASSEMBLE
    if (solver_kw.get("method") or DEFAULT_METHOD) == "ipm" and assemble == "device":
        # (options of the first-order solver -- max_iter, eps_* -- are accepted and unused here, as on the host path)
        x, hmp, info, _ = _solve_batch_ap_device(int(n), f_list, a_list, d_list, obj_list, peak_list,
                                                 ipm_max_iter=solver_kw.get("ipm_max_iter"))
        ok = info[:, 0] == 1.0
        status = ["Solved" if o else "Failed" for o in ok]               # fir_ap_cvx.m:176-182
        hs = [hmp[b] if ok[b] else None for b in range(B)]
        if return_info:
            return hs, status, dict(x=x, ripple_stop=info[:, 7].copy(), info=info)
        return hs, status
    if assemble not in ("device", "host"):
        raise ValueError(f"unknown assemble {assemble!r}: 'device' or 'host'")
    designs = [assemble_fir_ap(n, f_list[i], a_list[i], d_list[i], obj_list[i], peak_list[i]) for i in range(B)]
    x, t, info = _solve_batch_ap(n, designs, **solver_kw)
    ok = info[:, 0] == 1.0                # 1 solved; 2 infeasible certificate; 3 iteration limit -> 'Failed'
    status = ["Solved" if o else "Failed" for o in ok]   # fir_ap_cvx.m:176-182
    hs = [None] * B
    if ok.any():                          # h = fmp2(r) for all solved designs in one launch, :185-202
        hmp = fmp2_batch(np.stack([_x_to_r(x[b], n) for b in np.nonzero(ok)[0]]))
        for k, b in enumerate(np.nonzero(ok)[0]):
            hs[b] = hmp[k]
    if return_info:
        return hs, status, dict(x=x, ripple_stop=t, info=info)
    return hs, status


def fir_ap_cvx(n, f, a, d, obj=0, Peak=1e-3, dbg=0, **solver_kw):
    """[h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg) — fir_ap_cvx.m:1-245."""
    if obj < 0:
        raise ValueError("invalid input of obj")                          # :172-174
    hs, st = fir_ap_cvx_batch(int(n), [f], a, d, [obj], [Peak], **solver_kw)
    return (hs[0] if hs[0] is not None else np.zeros(0)), st[0]


def fir_ap(n, f, a, d, Peak=1e-3, min_order=0, min_tran=0, min_peak=0, dbg=0, **solver_kw):
    """[h, status, n_op, f_op] = fir_ap(...) — fir_ap.m:1-213: bisection on transition width and/or order.

    Each bisection step of the reference is one serial fir_ap_cvx solve; here every step probes a whole
    bracket of candidates in ONE batch (k-ary search), which needs fewer rounds and fills the GPU.
    min_peak: the result goes through fir_flip_zero (fir_ap.m:199-208), all flip patterns expanded in one GPU launch.
    """
    h, status, n_op, f_op = _fir_ap_search(n, f, a, d, Peak, min_order, min_tran, dbg, **solver_kw)
    if min_peak and len(h):                                               # fir_ap.m:199-208
        from .fir_post import fir_flip_zero
        h = fir_flip_zero(h, dbg)
    return h, status, n_op, f_op


def _fir_ap_search(n, f, a, d, Peak, min_order, min_tran, dbg, **solver_kw):
    """fir_ap.m:45-176: the transition-width and order bisections."""
    f = np.asarray(f, float).ravel()
    lam, df_thre = 0.1, 0.0005                                            # fir_ap.m:45-46
    n_op, f_op = int(n), f.copy()
    def solve1(nn, ff):
        """one fir_ap_cvx probe whose 'Failed' is only believed after a retry (fir_ap_cvx_decided)"""
        hq, sq, _ = fir_ap_cvx_decided(int(nn), [ff], a, d, [lam], [Peak], **solver_kw)
        return (hq[0] if hq[0] is not None else np.zeros(0)), sq[0]

    h1, status1 = solve1(n, f)                                            # :51
    if status1 == "Failed":
        raise RuntimeError("original parameters are too tight")          # :52-54
    h, status = h1, status1
    if min_tran == 0 and min_order == 0:
        return h, status, n_op, f_op

    def widen(fa):
        fn = f.copy()
        fn[0::2] -= fa                                                    # :71-73
        fn[1::2] += fa
        return fn

    if min_tran > 0:
        df_min = (f[2:-1:2] - f[1:-2:2]).min()                            # :60-61
        bot, top = 0.0, df_min / 2                                        # :62-63
        # fir_ap.m:78-105 is a binary bisection, one serial solve per step.  The next three steps can only
        # probe the 7 nodes bot + (top-bot)*j/8 of the bisection tree, whatever the outcomes: solve them as
        # one batch, then walk the tree exactly as the reference loop would (same probes, same decisions).
        while True:
            span = top - bot
            nodes = {j: bot + span * j / 8 for j in range(1, 8)}
            hs, sts, _ = fir_ap_cvx_decided(n, [widen(nodes[j]) for j in range(1, 8)], a, d, [lam] * 7, [Peak] * 7,
                                            **solver_kw)
            res = {j: (hs[j - 1], sts[j - 1]) for j in range(1, 8)}
            lo_j, hi_j, done = 0, 8, False
            for _ in range(3):
                mid = (lo_j + hi_j) // 2                                  # f_add_mid = (bot+top)/2, :79
                h0, st0 = res[mid]
                if st0 == "Failed":
                    hi_j = mid                                            # :86-88
                else:
                    h, status, lo_j = h0, st0, mid                        # :89-93
                if span * (hi_j - lo_j) / 8 < df_thre:                    # :100-102
                    done = True
                    break
            bot, top = bot + span * lo_j / 8, bot + span * hi_j / 8
            if done:
                break
        if not (0 < min_tran <= 1):
            raise ValueError("invalid input of min_tran")                 # :131-133
        fa = bot * min_tran                                               # :110
        h0, st0 = solve1(n, widen(fa))                                    # :116
        if st0 == "Failed":
            fa = bot                                                      # :117-121
        else:
            h, status = h0, st0
        f = widen(fa)
        f_op = f.copy()
    if min_order > 0:
        n_top, n_bot = int(n), 2                                          # :140-141
        # fir_ap.m:143-162 probes n_mid = ceil((n_top+n_bot)/2) serially.  Orders differ, so the probes cannot share
        # a matrix; the next three levels of the bisection tree (<= 7 orders) are solved concurrently instead and
        # the tree is then walked with exactly the reference's decisions.
        def step(st, solved):
            bot, top, mid = st
            if solved:
                top = mid
            else:
                bot = mid
            return (bot, top, int(np.ceil((top + bot) / 2)) if top - bot > 1 else None)

        state = (n_bot, n_top, int(np.ceil((n_top + n_bot) / 2)) if n_top - n_bot > 1 else None)
        cache = {}
        while state[2] is not None:
            need = [q for q in _speculate(state, step, 3) if q not in cache]
            res = _solve_concurrently([(lambda q=q: solve1(q, f)) for q in need])
            cache.update(dict(zip(need, res)))
            for _ in range(3):
                if state[2] is None:
                    break
                h0, st0 = cache[state[2]]
                if st0 != "Failed":
                    h, status = h0, st0
                state = step(state, st0 != "Failed")
        n_bot, n_top = state[0], state[1]
        if min_order == 1:
            n_op = n_top                                                  # :164-166
        elif 0 < min_order < 1:
            n_new = int(np.ceil(n * (1 - min_order) + n_top * min_order)) # :169-173
            h, status = solve1(n_new, f)
            n_op = n_new
        else:
            raise ValueError("invalid input of min_order")                # :174-176
    return h, status, n_op, f_op


def fir_qp(n, f, a, d, min_order=0, min_tran=0, min_peak=0, dbg=0, **solver_kw):
    """[h, status] = fir_qp(n, f, a, d, min_order, min_tran, min_peak, dbg) — fir_qp.m:1-150.

    The low-pass relative of fir_ap: every probe is fir_ap_cvx(n, f, a, d, 1e5) (stop-band weight lambda = 1e5, default Peak;
    fir_qp.m:47,68,92,107,130 -- the function never calls fir_qp_cvx), the transition search moves the single transition of
    f = [-fp, fp, fs, f4] symmetrically about its centre (:57-85, threshold 1e-3), the order search is the bisection of fir_ap
    (:103-123).  As in fir_ap the next three levels of each bisection tree are solved as one batch / concurrently and the tree
    is then walked with the reference's decisions; iteration-limit probes are retried before their 'Failed' is believed."""
    f = np.asarray(f, float).ravel()
    lam, df_thre = 1e5, 0.001                                             # fir_qp.m:36-37
    peak = 1e-3                                                           # fir_ap_cvx's default (fir_ap_cvx.m:33)

    def solve1(nn, ff):
        hq, sq, _ = fir_ap_cvx_decided(int(nn), [ff], a, d, [lam], [peak], **solver_kw)
        return (hq[0] if hq[0] is not None else np.zeros(0)), sq[0]

    h, status = solve1(n, f)                                              # :47
    if status == "Failed":
        raise RuntimeError("original parameters are too tight")          # :48-50
    if min_tran > 0:
        if f.size != 4:
            raise ValueError("fir_qp's transition search is written for f = [-fp, fp, fs, f4] (fir_qp.m:66-67)")
        centre = (f[2] + f[1]) / 2                                        # :58
        top, bot = (f[2] - f[1]) / 2, 0.0                                 # :59-60
        edges = lambda dfv: np.array([-(centre - dfv), centre - dfv, centre + dfv, f[3]])   # noqa: E731  :64-67
        while True:
            # the next three bisection steps can only probe bot + (top - bot) j / 8, j = 1..7: one batch, then the reference's walk
            span = top - bot
            nodes = {j: bot + span * j / 8 for j in range(1, 8)}
            hs, sts, _ = fir_ap_cvx_decided(n, [edges(nodes[j]) for j in range(1, 8)], a, d, [lam] * 7, [peak] * 7, **solver_kw)
            lo_j, hi_j, done = 0, 8, False
            for _ in range(3):
                mid = (lo_j + hi_j) // 2                                  # df_mid = (df_top + df_bot) / 2, :63
                if sts[mid - 1] == "Failed":
                    lo_j = mid                                            # :69-71: a narrower transition cannot be designed
                else:
                    h, status, hi_j = hs[mid - 1], sts[mid - 1], mid      # :72-76
                if span * (hi_j - lo_j) / 8 < df_thre:                    # :78-80
                    done = True
                    break
            bot, top = bot + span * lo_j / 8, bot + span * hi_j / 8
            if done:
                break
        if not (0 < min_tran <= 1):
            raise ValueError("invalid input of min_tran")                 # :97-99
        df_new = ((f[2] - f[1]) / 2) * (1 - min_tran) + top * min_tran    # :90
        f = edges(df_new)                                                 # :91-96
        h, status = solve1(n, f)
    elif min_tran != 0:
        raise ValueError("invalid input of min_tran")
    if min_order > 0:
        def step(st, solved):                                             # :104-122
            b_, t_, mid = st
            if solved:
                t_ = mid
            else:
                b_ = mid
            return (b_, t_, int(np.ceil((t_ + b_) / 2)) if t_ - b_ > 1 else None)

        state = (2, int(n), int(np.ceil((int(n) + 2) / 2)) if int(n) - 2 > 1 else None)
        cache = {}
        while state[2] is not None:
            need = [q for q in _speculate(state, step, 3) if q not in cache]
            cache.update(dict(zip(need, _solve_concurrently([(lambda q=q: solve1(q, f)) for q in need]))))
            for _ in range(3):
                if state[2] is None:
                    break
                h0, st0 = cache[state[2]]
                if st0 != "Failed":
                    h, status = h0, st0
                state = step(state, st0 != "Failed")
        n_top = state[1]
        if 0 < min_order < 1:                                             # :128-131
            h, status = solve1(int(np.ceil(n * (1 - min_order) + n_top * min_order)), f)
        elif min_order != 1:
            raise ValueError("invalid input of min_order")                # :132-134
    elif min_order != 0:
        raise ValueError("invalid input of min_order")
    if min_peak and len(h):                                               # :136-146
        from .fir_post import fir_flip_zero
        h = fir_flip_zero(h, dbg)
    return h, status


def sweep_grid(f, objs, peaks, f_adds):
    """The trade-off sweep of BASELINE config 4: every combination of the stop-band weight
    `obj`, the peak bound `Peak` and the band-edge expansion `f_add` (fir_ap.m:70-83 widens every band by
    f_add on both sides).  Returns (f_list, obj_list, peak_list), f_add fastest."""
    f = np.asarray(f, float).ravel()
    fl, ol, pl = [], [], []
    for o in objs:
        for pk in peaks:
            for fa in f_adds:
                fn = f.copy()
                fn[0::2] -= fa
                fn[1::2] += fa
                fl.append(fn)
                ol.append(float(o))
                pl.append(float(pk))
    return fl, ol, pl


def fir_ap_cvx_sweep(n, f, a, d, objs, peaks, f_adds, rank=0, world=1, batch=512, seed_stride=0, **solver_kw):
    """Solve this rank's share of the sweep grid in batches of at most `batch` designs.  Design instance i goes
    to rank i mod world: neighbouring instances differ in one parameter and cost about the same, so
    the interleaving balances the ranks; nothing is exchanged until the caller gathers the results.
    seed_stride > 1: two passes per batch.  Designs that share band edges and Peak form a chain ordered by the stop-band
    weight; every seed_stride-th design of a chain is solved cold, the others start from the solution (x, multipliers,
    primal weight) of the nearest solved chain member -- what fir_ap.m-style searches do serially with their previous answer.
    seed_stride = "auto" spaces the seeds about 0.1 decades of the weight apart and runs coarser sweeps cold.
    (Splitting a batch into sub-batches solved concurrently from several host threads was measured and is slower:
    the persistent product kernels own every SM, so the streams serialise: 5.6 s -> 7.1 s / 13 s with 2 / 4 streams.)
    Returns dict(index, x, ripple_stop, info) for the local designs."""
    fl, ol, pl = sweep_grid(f, objs, peaks, f_adds)
    mine = np.arange(rank, len(fl), world)
    concurrent = int(solver_kw.pop("concurrent_batches", 0) or 0)

    assemble = solver_kw.pop("assemble", None) or DEFAULT_ASSEMBLE
    on_device = (solver_kw.get("method") or DEFAULT_METHOD) == "ipm" and assemble == "device"

    def solve_batch(ids):
        if on_device:                                   # specification in, solutions out: no host assembly, no [M x B] upload
            x, _, info, _ = _solve_batch_ap_device(n, [fl[i] for i in ids], [a] * len(ids), [d] * len(ids), [ol[i] for i in ids],
                                                   [pl[i] for i in ids], ipm_max_iter=solver_kw.get("ipm_max_iter"), want_h=False)
            return x, info[:, 7].copy(), info
        designs = [assemble_fir_ap(n, fl[i], a, d, ol[i], pl[i]) for i in ids]
        stride = seed_stride
        if stride == "auto":                            # seeds about 0.1 decades of the weight apart; coarser sweeps run cold
            lo = np.sort(np.log10(np.unique([ol[i] for i in ids])))
            step = float(np.median(np.diff(lo))) if lo.size > 1 else np.inf
            stride = int(0.1 / step) if step > 0 else 0
            if stride < 8:
                stride = 0
        if stride and stride > 1 and len(designs) > 2 * stride:
            return _solve_seeded(n, designs, [np.asarray(fl[i]).tobytes() for i in ids], [pl[i] for i in ids],
                                 [ol[i] for i in ids], int(stride), **solver_kw)
        return _solve_batch_ap(n, designs, **solver_kw)

    # Batch composition.  A batch runs in lockstep until its slowest design is done, and the designs of this grid fall into
    # groups with very different iteration counts (infeasible: certificate after ~17 iterations; solved: ~45; edge of
    # feasibility: 70-100).  What decides the group is mostly Peak and the band edges, not the weight, so the local instances
    # are ordered by (Peak, band edges, weight) before they are cut into batches: batches of hopeless Peak values end early
    # instead of waiting for the stragglers of a mixed batch.  Results are returned in the order of `index` either way.
    # The batches with the largest Peak (the ones with solvable designs and stragglers) go first, so that the short batches fill
    # the end of the schedule when several batches are in flight.
    order = solver_kw.pop("order", "grouped")
    if order in ("grouped", "grouped_ascending") and mine.size > batch:
        sign = -1.0 if order == "grouped" else 1.0
        key = sorted(range(mine.size), key=lambda k: (sign * pl[mine[k]], tuple(fl[mine[k]]), ol[mine[k]]))
        mine = mine[np.array(key, dtype=int)]
    chunks = [mine[b0:b0 + batch] for b0 in range(0, mine.size, batch)]
    if concurrent > 1 and len(chunks) > 1 and (solver_kw.get("method") or DEFAULT_METHOD) == "ipm":
        # interior point: the per-design kernels (Cholesky, triangular solves) of a batch leave SMs idle once most of its designs
        # have finished; a second batch on another host thread / stream fills them
        from concurrent.futures import ThreadPoolExecutor
        import ctypes as _C
        dev = _C.c_int(0)
        have_dev = lib().mbrf_get_device(_C.byref(dev)) == 0   # no device: the solves themselves fail loudly

        def run(ids):
            if have_dev:                                   # worker threads start on device 0: hand them the caller's device
                check(lib().mbrf_set_device(dev.value))
            return solve_batch(ids)
        with ThreadPoolExecutor(max_workers=concurrent) as ex:
            results = list(ex.map(run, chunks))
    else:
        results = [solve_batch(ids) for ids in chunks]
    xs, ts, infos = [], [], []
    for x, t, info in results:
        xs.append(x); ts.append(t); infos.append(info)
    cat = lambda v, w: np.concatenate(v) if v else np.zeros((0, w))   # noqa: E731
    back = np.argsort(mine, kind="stable")                # results in ascending instance order, whatever the batch composition
    return dict(index=mine[back], x=cat(xs, 2 * n - 1)[back], ripple_stop=(np.concatenate(ts) if ts else np.zeros(0))[back],
                info=cat(infos, 8)[back])


def _solve_seeded(n, designs, fkeys, peaks, objs, stride, **solver_kw):
    """Two-pass solve of one batch (see fir_ap_cvx_sweep).  Designs sharing band edges and Peak form a chain ordered by the
    stop-band weight; every `stride`-th chain member (and the largest weight, the slowest to converge) is a seed.  Pass 1
    solves the seeds cold, pass 2 all other designs, each started from its nearest seed (x, multipliers, primal weight).
    All designs of a chain share grid rows, so the multipliers carry over row by row.
    Measured (512 designs, one B200): only SHORT steps in the weight help -- pass 2 needs ~1/8 of the cold iterations when
    the seed is within 0.05-0.4 decades, while a start 1.5 decades away is slower than a cold start, and chaining warm
    starts from seed to seed was erratic (some links ran into the iteration limit).  Hence cold seeds, one warm hop."""
    B = len(designs)
    chains = {}
    for b in range(B):
        chains.setdefault((fkeys[b], peaks[b]), []).append(b)
    seeds, src = [], np.full(B, -1)
    for members in chains.values():
        members.sort(key=lambda b: objs[b])
        pick = members[stride // 2::stride] or [members[len(members) // 2]]
        if members[-1] not in pick:
            pick = pick + [members[-1]]
        seeds += pick
        lo = np.log(np.array([objs[b] for b in pick]))
        for b in members:
            src[b] = pick[int(np.argmin(np.abs(lo - np.log(objs[b]))))]
    seeds = sorted(set(seeds))
    sset = set(seeds)
    rest = [b for b in range(B) if b not in sset]
    x = np.zeros((B, 2 * n - 1)); t = np.zeros(B); info = np.zeros((B, 8))
    xs_, ts_, is_, ys_, os_ = _solve_batch_ap(n, [designs[b] for b in seeds], want_dual=True, **solver_kw)
    pos = {b: k for k, b in enumerate(seeds)}
    x[seeds], t[seeds], info[seeds] = xs_, ts_, is_
    if rest:
        k = np.array([pos[src[b]] for b in rest])
        good = is_[k, 0] == 1                       # a seed that did not converge is no starting point
        x0 = np.where(good[:, None], xs_[k], 0.0)
        y0 = np.where(good[:, None], ys_[k], 0.0)
        om = np.where(good, os_[k], 1.0)
        xr, tr, ir = _solve_batch_ap(n, [designs[b] for b in rest], warm=(x0, y0, om), **solver_kw)
        x[rest], t[rest], info[rest] = xr, tr, ir
    return x, t, info


# --------------------------------------------------------------------------------------------
# ss/fir_linprog.m — linear-phase (real or complex-Hermitian) FIR by LP
# --------------------------------------------------------------------------------------------
def assemble_fir_linprog(n, f, a, d, a_min=None):
    """ss/fir_linprog.m:46-240 -> rows w, bounds, column description, objective.  Returns None for the
    'n even and amplitude 1 at fs/2' case the reference rejects up front (:63-75).
    a_min (ss/fir_min_order.m:14, ss/fir_pm.m:42-43,102): lower bound of the response in the transition regions; its default
    min(0, min(a - d)) is what fir_linprog.m:165-170 uses."""
    f = np.asarray(f, float).ravel() * np.pi                              # :46
    a = np.asarray(a, float).ravel()
    d = np.asarray(d, float).ravel()
    real_filter = not (f.min() < 0)                                       # :48-52
    odd = (n & 1) == 1                                                    # :56-60
    if not odd:
        idx = np.nonzero(np.abs(f) == np.pi)[0]                           # :66-75
        if np.any(a[idx] == 1):
            return None
    nhalf = int(np.ceil(n / 2))                                           # :79
    oversamp = 15                                                         # :92
    if real_filter:
        w = np.linspace(0, np.pi, oversamp * n)                           # :96-98
    else:
        w = np.linspace(-np.pi, np.pi, 2 * oversamp * n)                  # :99-101
    w = np.sort(np.concatenate([w, f]))                                   # :107
    idx_band, idx_tran, U, L = _bands(w, f, a, d)
    if idx_tran.size:                                                     # :163-171
        U_tran = np.full(idx_tran.size, U.max())
        L_tran = np.full(idx_tran.size, min(0.0, L.min()) if a_min is None or np.size(a_min) == 0 else float(a_min))
    else:
        U_tran = L_tran = np.zeros(0)
    w = np.concatenate([w[idx_band], w[idx_tran]])                        # :175-180
    hi = np.concatenate([U, U_tran])                                      # :221-226 (amplitude, not power)
    lo = np.concatenate([L, L_tran])
    ntran = idx_tran.size
    if odd:                                                               # :195-217
        kc = np.arange(1, nhalf, dtype=float)
        types = [0] + [1] * (nhalf - 1)
        kappa = [0.0] + list(kc)
        amp = [1.0] + [2.0] * (nhalf - 1)
        if not real_filter:
            types += [2] * (nhalf - 1)
            kappa += list(kc)
            amp += [2.0] * (nhalf - 1)
    else:
        kc = np.arange(0, nhalf, dtype=float) + 0.5
        types = [1] * nhalf
        kappa = list(kc)
        amp = [2.0] * nhalf
        if not real_filter:
            types += [2] * nhalf
            kappa += list(kc)
            amp += [2.0] * nhalf
    return dict(n=n, nhalf=nhalf, real=real_filter, odd=odd, w=w, lo=lo, hi=hi, nband_rows=idx_band.size, ntran=ntran,
                col_type=np.array(types, np.int32), col_kappa=np.array(kappa, float), col_amp=np.array(amp, float))


def _lp_objective(p):
    """fmin = sum(A(idx_tran,:), 1)  (fir_linprog.m:231) evaluated from the column description."""
    wt = p["w"][p["nband_rows"]:]
    c = np.zeros(p["col_type"].size)
    for j, (t, k, am) in enumerate(zip(p["col_type"], p["col_kappa"], p["col_amp"])):
        c[j] = am * (wt.size if t == 0 else (np.cos(wt * k).sum() if t == 1 else np.sin(wt * k).sum()))
    return c


def _fill_h(x, p):
    """fill_h, fir_linprog.m:274-296."""
    nh = p["nhalf"]
    if p["real"]:
        return np.concatenate([x[:0:-1], x]) if p["odd"] else np.concatenate([x[::-1], x])
    if p["odd"]:
        h = x[:nh] + 1j * np.concatenate([[0.0], x[nh:]])
        return np.concatenate([np.conj(h[:0:-1]), h])
    h = x[:nh] + 1j * x[nh:]
    return np.concatenate([np.conj(h[::-1]), h])


def _fill_opt_param(h0, p):
    """fill_opt_param, ss/fir_linprog.m:298-373, for a previous filter h0 of the same parity: the taps that fit are copied
    into the optimisation vector, the rest stays zero.  Returns None where the reference falls back to its FFT initialisation
    (no h0, or the other parity, :322-349) -- the solver then starts from its own point."""
    if h0 is None or len(h0) == 0:
        return None
    h0 = np.asarray(h0).ravel()
    nh = h0.size
    if bool(nh & 1) != bool(p["odd"]):                                    # :332-336
        return None
    nh_half = int(np.ceil(nh / 2))
    nx = p["col_type"].size
    nx_half = nx if p["real"] else ((nx + 1) // 2 if p["odd"] else nx // 2)   # :305-313
    k = min(nx_half, nh_half)
    x0 = np.zeros(nx)
    if p["odd"]:                                                          # :353-361 (1-based h0(nh_half : nh_half+k-1))
        x0[:k] = h0[nh_half - 1:nh_half - 1 + k].real
        if not p["real"]:
            x0[nx_half:nx_half + k - 1] = h0[nh_half:nh_half + k - 1].imag
    else:                                                                 # :362-371 (h0(nh_half+1 : nh_half+k))
        x0[:k] = h0[nh_half:nh_half + k].real
        if not p["real"]:
            x0[nx_half:nx_half + k] = h0[nh_half:nh_half + k].imag
    return x0


def fir_linprog(n, f, a, d, h0=None, dbg=0, return_info=False, **solver_kw):
    """[h, status] = fir_linprog(n, f, a, d, h0, dbg) — ss/fir_linprog.m:2-272.

    The LP `min fmin*x s.t. [A;-A]x <= [U,-L]` (:246-252) is solved on the GPU, by default with the interior-point
    solver (method="ipm"; infeasible specifications end with a Farkas certificate -> 'Failed').  h0 is the reference's
    starting point for linprog's medium-scale algorithm (:157); an interior-point method starts from its own centred point
    (MATLAB's linprog ignores x0 for its interior-point algorithm too), so with method="ipm" h0 does not change the result and
    is not used.  With method="pdhg" a previous filter of the same parity is the starting iterate, mapped as fill_opt_param
    does (:298-373) -- what the reference's order searches do with `hbest`.
    method="pdhg" (first order) adds the redundant box |x_j| <= 2*max|bounds| so that its dual bound certifies
    infeasibility (|H| <= max U on a 15x oversampled grid bounds every Fourier coefficient by it)."""
    n = int(n)
    p = assemble_fir_linprog(n, f, a, d, a_min=solver_kw.pop("a_min", None))
    if p is None:
        return (np.zeros(0), "Failed", dict(info=None)) if return_info else (np.zeros(0), "Failed")   # :68-72
    M, N = p["w"].size, p["col_type"].size
    c = _lp_objective(p)
    arr = lambda v: np.ascontiguousarray(v, dtype=np.float64)            # noqa: E731
    lo, hi, cc = arr(p["lo"].reshape(M, 1)), arr(p["hi"].reshape(M, 1)), arr(c.reshape(N, 1))
    z, info = np.zeros((N, 1)), np.zeros((1, 8))
    w_row, kap, amp = arr(p["w"]), arr(p["col_kappa"]), arr(p["col_amp"])
    method = solver_kw.pop("method", None) or DEFAULT_METHOD
    ipm_max_iter = solver_kw.pop("ipm_max_iter", None)
    if method == "ipm":
        # interior point (csrc/ipm.cu): the LP exactly as posed, no auxiliary box
        check(lib().mbrf_fir_ipm_solve(_dp(w_row), M, _ip(p["col_type"]), _dp(kap), _dp(amp), N, None, None, 0, _dp(cc), _dp(lo),
                                       _dp(hi), None, None, None, 1, 0, 0, None, int(ipm_max_iter or IPM_MAX_ITER),
                                       IPM_FEASTOL, IPM_RELTOL, IPM_ABSTOL, _dp(z), _dp(info)))
    elif method == "pdhg":
        big = 2.0 * max(np.abs(p["hi"]).max(), np.abs(p["lo"]).max())
        bl, bu = arr(np.full((N, 1), -big)), arr(np.full((N, 1), big))
        upper = arr([p["ntran"] * p["hi"].max() + 1e-9])                  # fmin*x = sum_tran H <= ntran*max U
        kw = dict(max_iter=MAX_ITER, check_every=CHECK_EVERY, eps_pr=EPS_PR, eps_dr=EPS_DR, eps_gap=EPS_GAP)
        kw.update(solver_kw)
        x0 = _fill_opt_param(h0, p)
        if x0 is not None:                                                # x0 of fir_linprog.m:157 -> starting iterate
            z0 = arr(x0.reshape(N, 1))
            check(lib().mbrf_fir_pdhg_warm_start(_dp(z0), None, None, None, None))
        check(lib().mbrf_fir_pdhg_solve(_dp(w_row), None, M, _ip(p["col_type"]), _dp(kap), _dp(amp), N, -1, None, None, 0,
                                        _dp(cc), _dp(lo), _dp(hi), _dp(bl), _dp(bu), None, 1, _dp(upper), 0, 0, None,
                                        int(kw["max_iter"] or MAX_ITER), int(kw["check_every"]), float(kw["eps_pr"]),
                                        float(kw["eps_dr"]), float(kw["eps_gap"]), _dp(z), _dp(info), None))
    else:
        raise ValueError(f"unknown method {method!r}: 'ipm' or 'pdhg'")
    ok = info[0, 0] == 1.0                                                # exitflag == 1, :265
    h = _fill_h(z[:, 0], p) if ok else np.zeros(0)
    st = "Solved" if ok else "Failed"
    if return_info:
        return h, st, dict(x=z[:, 0].copy(), info=info[0].copy(), problem=p, c=c)
    return h, st


def _fir_linprog_decided(n, f, a, d, h0=None, dbg=0, **solver_kw):
    """fir_linprog as a bisection probe: an iteration-limit ending (status 3) is retried with more iterations, see
    fir_ap_cvx_decided."""
    import warnings
    h, st, ex = fir_linprog(n, f, a, d, h0, dbg, return_info=True, **solver_kw)
    if isinstance(ex, dict) and ex.get("info") is not None and int(ex["info"][0]) == 3:
        h, st, ex = fir_linprog(n, f, a, d, h0, dbg, return_info=True, **_retry_kw(solver_kw))
        if int(ex["info"][0]) == 3:
            warnings.warn(f"fir_linprog probe at n = {n} stayed undecided at the iteration limit and is treated as 'Failed'",
                          UndecidedProbe, stacklevel=2)
    return h, st


def _min_order_search(n, f, a, d, even_odd, solve, pick_longer):
    """The bisection of ss/fir_min_order_linprog.m:84-232 (and ss/fir_min_order.m:84-226): odd lengths, then
    even lengths capped by the best odd one; `solve(n_tap, h_warm)` is one feasibility probe."""
    if even_odd not in (1, 2):
        even_odd = 0                                                      # :65-69
    hbest_odd, hbest_even = None, None
    n_odd_max = 2 * ((n - 1) // 2) + 1                                    # :78-79
    n_even_max = 2 * (n // 2)

    def bisect(n_top, tap_of, hbest):
        # state (n_bot, n_top, n_cur): the reference's loop, :91-147 / :152-211, one probe per step; the next three
        # levels of its decision tree are solved concurrently (orders differ -> separate matrices), then walked.
        def step(st, solved):
            bot, top, cur = st
            if solved:
                top = cur
                cur = bot if top == bot + 1 else int(np.ceil((top + bot) / 2))   # :131-136
            else:
                bot = cur
                cur = int(np.ceil((bot + top) / 2))                              # :141-142
            return (bot, top, cur if top - bot > 1 else None)

        state = (1, n_top, n_top if n_top - 1 > 1 else None)
        cache = {}
        while state[2] is not None:
            need = [q for q in _speculate(state, step, 3) if q not in cache]
            res = _solve_concurrently([(lambda q=q: solve(tap_of(q), None)) for q in need])
            cache.update(dict(zip(need, res)))
            for _ in range(3):
                if state[2] is None:
                    break
                h, st = cache[state[2]]
                if st == "Solved":
                    hbest = h
                state = step(state, st == "Solved")
        return hbest

    if even_odd != 2:
        hbest_odd = bisect((n_odd_max + 1) // 2, lambda k: 2 * k - 1, None)
    if even_odd != 1:
        n_top = n_even_max // 2 if hbest_odd is None else min(n_even_max // 2, (len(hbest_odd) + 1) // 2)
        hbest_even = bisect(n_top, lambda k: 2 * k, None)
    if hbest_odd is None and hbest_even is None:
        return np.zeros(0), "Failed"
    if pick_longer:                                                       # fir_min_order.m:222-226 (its quirk)
        lo_, le_ = (0 if hbest_odd is None else len(hbest_odd)), (0 if hbest_even is None else len(hbest_even))
        return (hbest_odd if lo_ > le_ else hbest_even), "Solved"
    if hbest_odd is None:
        return hbest_even, "Solved"
    if hbest_even is None or len(hbest_odd) < len(hbest_even):            # fir_min_order_linprog.m:220-228
        return hbest_odd, "Solved"
    return hbest_even, "Solved"


def fir_min_order_linprog(n, f, a, d, even_odd=0, dbg=0, **solver_kw):
    """[h, status] = fir_min_order_linprog(n, f, a, d, even_odd, dbg) — ss/fir_min_order_linprog.m:54-234."""
    return _min_order_search(int(n), f, a, d, even_odd, lambda nt, hw: _fir_linprog_decided(nt, f, a, d, hw, dbg, **solver_kw),
                             pick_longer=False)


def fir_min_order(n, f, a, d, even_odd=0, a_min=None, dbg=0, **solver_kw):
    """[h, status] = fir_min_order(n, f, a, d, even_odd, a_min, dbg) — ss/fir_min_order.m:55-230.

    The reference probes with fir_pm -> cfirpm (closed-source Parks-McClellan, ss/fir_pm.m:175), which cannot
    be reproduced; per BASELINE.json's north star the search is re-expressed as LP feasibility: the same
    bisection (and the same 'longer of odd/even' selection, :222-226) with fir_linprog probes.  a_min is fir_pm's lower
    bound of the response in the transition regions (ss/fir_pm.m:42-43,102: default min(0, min(a - d)), which is also the
    LP's default, fir_linprog.m:165-170); a given value replaces that bound in every probe."""
    kw = dict(solver_kw)
    if a_min is not None and np.size(a_min):
        kw["a_min"] = float(a_min)
    return _min_order_search(int(n), f, a, d, even_odd, lambda nt, hw: _fir_linprog_decided(nt, f, a, d, hw, dbg, **kw),
                             pick_longer=True)


# --------------------------------------------------------------------------------------------
# fir_qp_cvx.m — min-energy / min-peak FIR with quadratic phase target (SOCP)
# --------------------------------------------------------------------------------------------
def assemble_fir_qp(n, f, a, d, k=100.0, oversamp=10):
    """fir_qp_cvx.m:34-121: grid, bands, desired response Hd, radii.  Rows ordered band then transition."""
    f = np.asarray(f, float).ravel() * np.pi                              # :34
    a = np.asarray(a, float).ravel()
    d = np.asarray(d, float).ravel()
    m = n * oversamp                                                      # :35-36
    w = np.sort(np.concatenate([np.linspace(-np.pi, np.pi, m), f]))       # :37-38
    nband = len(f) // 2
    idx_band, M_band, D_band = [], [], []
    for b in range(nband):                                                # :46-61
        lo_, hi_ = f[2 * b], f[2 * b + 1]
        idx = np.nonzero((w >= lo_) & (w <= hi_))[0]
        idx_band.append(idx)
        amp = np.full(idx.size, a[2 * b]) if lo_ == hi_ else \
            a[2 * b] + (a[2 * b + 1] - a[2 * b]) * ((w[idx] - lo_) / (hi_ - lo_))
        M_band.append(amp)
        D_band.append(np.full(idx.size, d[b]))
    idx_band = np.concatenate(idx_band)
    M_band, D_band = np.concatenate(M_band), np.concatenate(D_band)
    mask = np.ones(w.size, bool)
    mask[idx_band] = False
    wband, wtran = w[idx_band], w[mask]                                   # :75-76
    Hd = M_band * np.exp(1j * (k * wband ** 2 - wband * (n - 1) / 2))     # :113-121
    return dict(n=n, w=np.concatenate([wband, wtran]), center=np.concatenate([Hd, np.zeros(wtran.size, complex)]),
                radius=np.concatenate([D_band, np.full(wtran.size, 1 + d.max() * 5)]), nband=wband.size)   # :151,156


def fir_qp_cvx(n, f, a, d, k=100, obj=0, dbg=0, return_info=False, oversamp=10, **solver_kw):
    """[h, status] = fir_qp_cvx(n, f, a, d, k, obj, dbg) — fir_qp_cvx.m:1-243.

    Scalar `obj` (:145-166):
        minimise E_total + obj*Peak   s.t.  ||A_i x - Hd_i|| <= D_i  (bands),  ||A_i x|| <= 1+5 max(d)  (transitions),
                                            ||(x_i, x_{n+i})|| <= Peak,  ||x|| <= E_total.
    Two-element `obj` (:170-191, minimax on the frequency response):
        minimise delta + obj(1)*E_total + obj(2)*Peak   s.t.  ||A_i x - Hd_i|| <= D_i*delta  (bands),
                                            ||A_i x|| <= 1.1  (transitions), Peak and E_total as above.

    E_total, Peak and delta are eliminated (E_total = ||x||, Peak = max_i ||(x_i, x_{n+i})||,
    delta = max_i ||A_i x - Hd_i|| / D_i at the optimum): the solver sees the norm term, a group block over identity
    rows, one disk per constrained grid point and — for the minimax form — a centred group block over the band rows
    scaled by 1/D_i.  No epigraph variables."""
    minimax = np.size(obj) == 2
    if np.size(obj) not in (1, 2):
        raise ValueError("invalid input of obj")                          # :193-195
    n = int(n)
    ob = [float(v) for v in np.ravel(obj)]
    p = assemble_fir_qp(n, f, a, d, float(k), oversamp=int(oversamp))    # oversamp = 10: fir_qp_cvx.m:35 (16 = the 4096-point grid)
    m, nb = p["w"].size, p["nband"]
    N, M = 2 * n, 2 * m + 2 * n
    w_row = np.concatenate([np.repeat(p["w"], 2), np.zeros(2 * n)])
    row_phase = np.concatenate([np.tile([0.0, np.pi / 2], m), np.zeros(2 * n)])   # [cos sin; -sin cos], :96-109
    row_scale = np.concatenate([np.ones(2 * m), np.zeros(2 * n)])
    kk = np.arange(n, dtype=float)
    col_type = np.concatenate([np.full(n, 1), np.full(n, 2)]).astype(np.int32)
    col_kappa = np.concatenate([kk, kk])
    col_amp = np.ones(N)
    ti = (2 * m + np.arange(2 * n)).astype(np.int32)                      # identity rows, pair i = (x_i, x_{n+i}), :126-139
    tj = np.empty(2 * n, np.int32)
    tj[0::2] = np.arange(n)
    tj[1::2] = n + np.arange(n)
    tv = np.ones(2 * n)
    lo = np.full((M, 1), -np.inf)
    hi = np.full((M, 1), np.inf)
    radius = p["radius"].copy()
    blocks = PdhgBlocks()
    one = np.array([1.0])
    if minimax:
        inv_d = 1.0 / p["radius"][:nb]                                    # band rows / D_i: ||.|| <= delta, :176
        row_scale[0:2 * nb:2] = inv_d
        row_scale[1:2 * nb:2] = inv_d
        lo[0:2 * nb:2, 0] = p["center"][:nb].real * inv_d
        lo[1:2 * nb:2, 0] = p["center"][:nb].imag * inv_d
        lo[2 * nb:2 * m, 0] = 0.0                                         # transition disks: centre 0, radius 1 + 0.1, :180
        hi[2 * nb:2 * m:2, 0] = 1.0 + 0.1
        radius[nb:] = 1.0 + 0.1
        blocks.group2_row0, blocks.group2_pairs, blocks.group2_w = 0, nb, _dp(one)
        blocks.disk_row0, blocks.disk_pairs = 2 * nb, m - nb
        gw, lam = np.array([ob[1]]), np.array([ob[0]])
    else:
        lo[0:2 * m:2, 0] = p["center"].real
        lo[1:2 * m:2, 0] = p["center"].imag
        hi[0:2 * m:2, 0] = p["radius"]
        blocks.disk_row0, blocks.disk_pairs = 0, m
        gw, lam = np.array([ob[0]]), np.array([1.0])
    # the solver sees the objective divided by its largest weight: at the reference's obj = 1e6 (dzrf_mb.m:211-213) the
    # problem is "minimise Peak" with the energy as a 1e-6 tie-break, and a first-order method needs O(1) weights
    oscale = float(solver_kw.pop("objective_scale", 0.0)) or max(1.0, float(gw[0]), float(lam[0]))
    gw, lam = gw / oscale, lam / oscale
    if minimax:
        one /= oscale                                                     # in place: blocks.group2_w already points at it
    big = 2.0 * p["radius"].max() + 2.0 * np.abs(p["center"]).max() + (2.0 if minimax else 0.0)
    c = np.zeros((N, 1))
    bl, bu = np.full((N, 1), -big), np.full((N, 1), big)
    blocks.group_row0, blocks.group_pairs, blocks.group_w = 2 * m, n, _dp(gw)
    blocks.norm_coords, blocks.norm_w = N, _dp(lam)
    kw = dict(max_iter=MAX_ITER, check_every=CHECK_EVERY, eps_pr=EPS_PR, eps_dr=EPS_DR, eps_gap=EPS_GAP)
    kw.update({k: v for k, v in solver_kw.items() if k != "want_dual"})
    z, info = np.zeros((N, 1)), np.zeros((1, 8))
    arr = lambda v: np.ascontiguousarray(v, dtype=np.float64)            # noqa: E731
    w_row, row_phase, row_scale, col_kappa, col_amp, tv, lo, hi = map(arr, (w_row, row_phase, row_scale, col_kappa,
                                                                             col_amp, tv, lo, hi))
    y_out = om_out = None
    if solver_kw.get("want_dual"):                                       # final multipliers [M] (rows as above), for certificates
        y_out, om_out = np.zeros((M, 1)), np.ones(1)
        check(lib().mbrf_fir_pdhg_warm_start(None, None, None, _dp(y_out), _dp(om_out)))
    check(lib().mbrf_fir_pdhg_solve2(_dp(w_row), _dp(row_phase), _dp(row_scale), M, _ip(col_type), _dp(col_kappa),
                                     _dp(col_amp), N, 2 * n, _ip(ti), _ip(tj), _dp(tv), None, None, 0, _dp(c), _dp(lo),
                                     _dp(hi), _dp(bl), _dp(bu), None, 1, None, C.byref(blocks), int(kw["max_iter"]),
                                     int(kw["check_every"]), float(kw["eps_pr"]), float(kw["eps_dr"]),
                                     float(kw["eps_gap"]), _dp(z), _dp(info), None))
    info[0, 2] *= oscale                                                  # objective / dual value / bound in the caller's units
    info[0, 3] *= oscale
    if abs(info[0, 6]) < 1e300:
        info[0, 6] *= oscale
    ok = info[0, 0] == 1.0
    x = z[:, 0]
    h = x[:n] + 1j * x[n:] if ok else np.zeros(0)                         # :209
    st = "Solved" if ok else "Failed"                                     # :200-206
    if return_info:
        p = dict(p, radius=radius)
        ex = dict(x=x.copy(), info=info[0].copy(), problem=p)
        if y_out is not None:
            ex["y"] = y_out[:, 0].copy()                                  # rows 2i, 2i+1: the disk of grid point i; then the identity rows
        return h, st, ex
    return h, st
