"""TEST INFRASTRUCTURE ONLY — numpy twin of the batched restarted-PDHG the GPU runs (csrc/fir_pdhg.cu),
plus the assembly of the fir_ap_cvx problem into the solver's canonical form.

Canonical form (one dense matrix K shared by a batch of B designs, everything else per design):

        minimise   c^T z      subject to   lo <= K z <= hi ,   z in X
        X = product of  boxes  z_j in [bl_j, bu_j]  and 2-D disks  ||(z_i, z_j)|| <= rho

fir_ap_cvx (fir_ap_cvx.m:160-169) maps onto it with z = x (2n-1):
  rows 0..m-1      A x            in [L_b, U_b]                      (:103-120)
  rows m..m+ns-1   A(stop) x      the "simplex block": obj * max_i (A x)_i is added to the objective, which is
                                  `obj*ripple_stop` s.t. `A_U(idx_stop,:)*x <= ripple_stop` (:163-165) with
                                  ripple_stop eliminated; its multipliers live on {y >= 0, sum y = obj}
  x1 in [-n Peak, n Peak],  ||(x_i, x_{n+i-1})|| <= (n-i+1) Peak     (:166-168)
Columns are scaled to unit norm (a cos/sin pair shares one scale so disks stay disks).

Dual function (exact, X compact):  q(y) = -sum(hi*y+ - lo*y-) + min_{z in X} (c + K^T y)^T z,
a rigorous lower bound on the optimum for every y the iteration produces; p - q is the stopping gap and
q > (upper bound on any feasible objective) certifies infeasibility ('Failed').

Parity status: PARITY UNPINNED by the reference (see fir_problems.py); pinned against HiGHS on the
cone-free LP and on polygonal inner/outer approximations of the cones (tests/test_oracle_fir.py).
"""
from __future__ import annotations

import numpy as np

from .fir_problems import matrix_fir_ap


def assemble_fir_ap(problems):
    """problems: list of build_fir_ap dicts sharing n and the grid w (bounds/objective/radii may differ).
    Returns the canonical batch: K (M,N), and per-design arrays shaped (dim, B)."""
    p0 = problems[0]
    n, w = p0["n"], p0["w"]
    for p in problems:
        assert p["n"] == n and p["w"].shape == w.shape and np.array_equal(p["w"], w), "batch must share the grid"
        assert np.array_equal(p["stop"], p0["stop"]), "batch must share the stop rows"
    A = matrix_fir_ap(w, n)
    m, nx = A.shape
    st = p0["stop"]
    ns = st.size
    B = len(problems)
    N = nx
    K = np.zeros((m + ns, N))
    K[:m] = A
    K[m:] = A[st]                      # simplex block: obj * max_i (A x)_i over the stop rows
    cs = np.sqrt((K ** 2).sum(0))
    pair_i = np.arange(1, n)
    pair_j = np.arange(n, 2 * n - 1)
    pm = np.sqrt(0.5 * (cs[pair_i] ** 2 + cs[pair_j] ** 2))
    cs[pair_i] = pm
    cs[pair_j] = pm
    K = K / cs
    lo = np.full((m + ns, B), -np.inf)
    hi = np.zeros((m + ns, B))
    c = np.zeros((N, B))
    bl = np.full((N, B), -np.inf)
    bu = np.full((N, B), np.inf)
    rho = np.zeros((n - 1, B))
    sw = np.zeros(B)
    for b, p in enumerate(problems):
        lo[:m, b] = p["lo"]
        hi[:m, b] = p["hi"]
        c[0, b] = p["c"][0] / cs[0]
        sw[b] = p["c"][-1]
        bl[0, b], bu[0, b] = -p["radius"][0] * cs[0], p["radius"][0] * cs[0]
        rho[:, b] = p["radius"][1:] * pm
    return dict(K=K, lo=lo, hi=hi, c=c, bl=bl, bu=bu, pair_i=pair_i, pair_j=pair_j, rho=rho, colscale=cs,
                n=n, m=m, ns=ns, srow0=m, sw=sw)


def proj_simplex(v, w):
    """Columns of v projected onto {y >= 0, sum y = w[b]} (sort-based; the GPU uses Michelot's iteration)."""
    ns, B = v.shape
    u = -np.sort(-v, axis=0)
    css = np.cumsum(u, axis=0) - w
    k = np.arange(1, ns + 1)[:, None]
    cond = u - css / k > 0
    r = ns - 1 - np.argmax(cond[::-1], axis=0)
    theta = css[r, np.arange(B)] / (r + 1)
    return np.where(w > 0, np.maximum(v - theta, 0.0), 0.0)


def proj_X(z, q):
    z = np.clip(z, q["bl"], q["bu"])
    zi, zj = z[q["pair_i"]], z[q["pair_j"]]
    r = np.hypot(zi, zj)
    s = np.where(r > q["rho"], q["rho"] / np.maximum(r, 1e-300), 1.0)
    z[q["pair_i"]] = zi * s
    z[q["pair_j"]] = zj * s
    return z


def dual_value_x(g, q):
    """min_{z in X} g^T z  (per design), g = c + K^T y."""
    inpair = np.zeros(g.shape[0], bool)
    inpair[q["pair_i"]] = True
    inpair[q["pair_j"]] = True
    with np.errstate(invalid="ignore"):
        box = np.where(g > 0, g * q["bl"], np.where(g < 0, g * q["bu"], 0.0))
    box = np.where(inpair[:, None], 0.0, box)
    disk = -(q["rho"] * np.hypot(g[q["pair_i"]], g[q["pair_j"]])).sum(0)
    return box.sum(0) + disk


def solve(q, max_iter=60000, check_every=64, eps_pr=8e-7, eps_dr=1e-4, eps_gap=5e-5, obj_upper=None, verbose=False):
    """Batched restarted PDHG.  Returns dict(z (N,B) in ORIGINAL units, obj, dual, pr, status, iters)."""
    K = q["K"]
    M, N = K.shape
    B = q["c"].shape[1]
    rng = np.random.default_rng(0)
    v = rng.normal(size=N)
    for _ in range(60):
        v = K.T @ (K @ v)
        nk = np.linalg.norm(v)
        v /= nk
    eta = 0.9 / np.sqrt(nk)
    omega = np.ones(B)
    z = proj_X(np.zeros((N, B)), q)
    y = np.zeros((M, B))
    zs, ys = np.zeros_like(z), np.zeros_like(y)
    z0, y0 = z.copy(), y.copy()
    cnt = np.zeros(B)
    last_err = np.full(B, np.inf)          # error at the last restart
    prev_err = np.full(B, np.inf)          # candidate error at the previous check
    status = np.zeros(B, int)              # 0 running, 1 solved, 2 infeasible
    done_iter = np.zeros(B, int)
    zbest, ybest = z.copy(), y.copy()
    since = np.zeros(B, int)               # iterations since this design's last restart
    tot = 0

    def metrics(zz, yy):
        """pr: max row violation; dr: natural residual ||z - P_X(z - g)||_inf; pobj; dobj = -h*(y) + g^T z
        (equals c^T z exactly when the rows are complementary); rig: rigorous lower bound q(y)."""
        Kz = K @ zz
        s0, ns_ = q.get("srow0", K.shape[0]), q.get("ns", 0)
        Kr, lo_r, hi_r, y_r = Kz[:s0], q["lo"][:s0], q["hi"][:s0], yy[:s0]
        pr = np.maximum(np.maximum(Kr - hi_r, lo_r - Kr), 0).max(0)
        g = q["c"] + K.T @ yy
        dr = np.abs(zz - proj_X(zz - g, q)).max(0)
        tmax = np.maximum(Kz[s0:s0 + ns_].max(0), 0.0) if ns_ else 0.0
        pobj = (q["c"] * zz).sum(0) + (q["sw"] * tmax if ns_ else 0.0)
        yp, ym = np.maximum(y_r, 0), np.maximum(-y_r, 0)
        hs = np.where(yp > 0, hi_r * yp, 0.0).sum(0) - np.where(ym > 0, lo_r * ym, 0.0).sum(0)
        dobj = -hs + (g * zz).sum(0)
        rig = -hs + dual_value_x(g, q)
        return pr, dr, pobj, dobj, rig

    for it in range(1, max_iter + 1):
        tau, sig = eta / omega, eta * omega
        zn = proj_X(z - tau * (q["c"] + K.T @ y), q)
        vv = y + sig * (K @ (2 * zn - z))
        wv = vv / sig
        with np.errstate(invalid="ignore"):
            y = np.where(wv > q["hi"], vv - sig * q["hi"], np.where(wv < q["lo"], vv - sig * q["lo"], 0.0))
        if q.get("ns", 0):
            s0_, ns_ = q["srow0"], q["ns"]
            y[s0_:s0_ + ns_] = proj_simplex(vv[s0_:s0_ + ns_], q["sw"])
        z = zn
        zs += z
        ys += y
        cnt += 1
        since += 1
        tot += 1
        if it % check_every:
            continue
        za, ya = zs / np.maximum(cnt, 1), ys / np.maximum(cnt, 1)
        pa, ra, oa, da, qa = metrics(za, ya)
        pc, rc, oc, dc, qc = metrics(z, y)
        ea = np.maximum(np.maximum(pa, ra), np.abs(oa - da))
        ec = np.maximum(np.maximum(pc, rc), np.abs(oc - dc))
        use_avg = ea < ec
        cz = np.where(use_avg, za, z)
        cy = np.where(use_avg, ya, y)
        ce = np.where(use_avg, ea, ec)
        cp = np.where(use_avg, pa, pc)
        cr = np.where(use_avg, ra, rc)
        co = np.where(use_avg, oa, oc)
        cd = np.where(use_avg, da, dc)
        solved = ((cp <= eps_pr) & (cr <= eps_dr) & (np.abs(co - cd) <= eps_gap * np.maximum(np.abs(co), 1e-12))
                  & (status == 0))
        if obj_upper is not None:
            infeas = (np.maximum(qa, qc) > obj_upper) & (status == 0) & ~solved
        else:
            infeas = np.zeros(B, bool)
        newly = solved | infeas
        zbest[:, newly] = cz[:, newly]
        ybest[:, newly] = cy[:, newly]
        status[solved] = 1
        status[infeas] = 2
        done_iter[newly] = it
        if verbose and it % (check_every * 16) == 0:
            print(it, " | ".join("pr %.1e dr %.1e obj %.7f d %.7f rig %.5f om %.1e" % (cp[b], cr[b], co[b], cd[b], max(qa[b], qc[b]), omega[b]) for b in range(min(B, 3))), status[:8])
        if (status != 0).all():
            break
        # restart rules (PDLP): sufficient decay, necessary decay + stall, or long since last restart
        do = (ce <= 0.2 * last_err) | ((ce <= 0.8 * last_err) & (ce > prev_err)) | (since >= 0.36 * tot)
        prev_err = ce
        if do.any():
            dz = np.linalg.norm(cz - z0, axis=0)
            dy = np.linalg.norm(cy - y0, axis=0)
            ok = do & (dz > 1e-12) & (dy > 1e-12)
            omega = np.where(ok, np.exp(0.5 * np.log(np.maximum(dy, 1e-300) / np.maximum(dz, 1e-300)) + 0.5 * np.log(omega)), omega)
            z = np.where(do, cz, z)
            y = np.where(do, cy, y)
            z0 = np.where(do, z, z0)
            y0 = np.where(do, y, y0)
            last_err = np.where(do, ce, last_err)
            since = np.where(do, 0, since)
            zs = np.where(do, 0.0, zs)
            ys = np.where(do, 0.0, ys)
            cnt = np.where(do, 0.0, cnt)
    run = status == 0
    zbest[:, run] = z[:, run]
    ybest[:, run] = y[:, run]
    done_iter[run] = tot
    pr, dr, pobj, dobj, rig = metrics(zbest, ybest)
    ripple = (K @ zbest)[q["srow0"]:q["srow0"] + q["ns"]].max(0) if q.get("ns", 0) else np.zeros(B)
    return dict(z=np.vstack([zbest / q["colscale"][:, None], ripple[None, :]]), obj=pobj, dual=dobj, pr=pr, dr=dr, rigorous_lower=rig, status=status, iters=done_iter, y=ybest)
