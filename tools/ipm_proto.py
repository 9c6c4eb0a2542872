#!/usr/bin/env python
"""Development prototype (numpy, dense): homogeneous self-dual primal-dual interior-point method for
    minimise c'x  s.t.  G x + s = h,  s in K = R+^nl x Q^q1 x ... x Q^qk
with Nesterov-Todd scaling and Mehrotra's corrector.  It is the numpy twin of csrc/ipm.cu (same iteration, same
stopping rules) used to debug the CUDA solver and to study iteration counts; nothing in the product imports it.

    python tools/ipm_proto.py 256 1e4        # fir_ap_cvx on the dual-band H-1 spec, N = 256, obj = 1e4, cone-free
"""
from __future__ import annotations

import sys
import time

import numpy as np
import scipy.linalg as sla

REG = 0.0
STEP = 0.99       # fraction of the way to the boundary
SIGPOW = 3        # sigma = (1 - alpha_aff)^SIGPOW
INIT = "e"        # "e": x = 0, s = z = e;  "ls": least-squares start shifted into the cone (CVXOPT)
NCORR = 0         # extra Gondzio centrality correctors


def _soc_step(u, du):
    """largest alpha with u + alpha*du in Q (u interior); blocks in rows: u [k, q]"""
    if u.shape[0] == 0:
        return np.inf
    a = du[:, 0] ** 2 - (du[:, 1:] ** 2).sum(1)
    b = 2 * (u[:, 0] * du[:, 0] - (u[:, 1:] * du[:, 1:]).sum(1))
    c = u[:, 0] ** 2 - (u[:, 1:] ** 2).sum(1)
    best = np.full(u.shape[0], np.inf)
    # u0 + alpha du0 >= 0
    neg = du[:, 0] < 0
    best[neg] = -u[neg, 0] / du[neg, 0]
    disc = b * b - 4 * a * c
    for k in range(u.shape[0]):
        if abs(a[k]) < 1e-300:
            if b[k] < 0:
                best[k] = min(best[k], -c[k] / b[k])
            continue
        if disc[k] < 0:
            continue
        sq = np.sqrt(disc[k])
        q = -0.5 * (b[k] + (sq if b[k] >= 0 else -sq))
        roots = [q / a[k]]
        if q != 0:
            roots.append(c[k] / q)
        for r in roots:
            if r > 0:
                best[k] = min(best[k], r)
    return best.min()


class Cone:
    def __init__(self, nl, qdims):
        self.nl = nl
        self.q = list(qdims)
        self.off = np.concatenate([[nl], nl + np.cumsum(self.q)]).astype(int)
        self.dim = int(self.off[-1])
        self.deg = nl + len(self.q)

    def blocks(self, v):
        return [v[self.off[i]:self.off[i + 1]] for i in range(len(self.q))]

    def e(self):
        v = np.zeros(self.dim)
        v[:self.nl] = 1
        for i in range(len(self.q)):
            v[self.off[i]] = 1
        return v

    def prod(self, u, v):
        o = np.empty(self.dim)
        o[:self.nl] = u[:self.nl] * v[:self.nl]
        for i in range(len(self.q)):
            a, b = self.off[i], self.off[i + 1]
            o[a] = u[a:b] @ v[a:b]
            o[a + 1:b] = u[a] * v[a + 1:b] + v[a] * u[a + 1:b]
        return o

    def div(self, u, d):
        """solve u o x = d"""
        o = np.empty(self.dim)
        o[:self.nl] = d[:self.nl] / u[:self.nl]
        for i in range(len(self.q)):
            a, b = self.off[i], self.off[i + 1]
            det = u[a] ** 2 - u[a + 1:b] @ u[a + 1:b]
            x0 = (u[a] * d[a] - u[a + 1:b] @ d[a + 1:b]) / det
            o[a] = x0
            o[a + 1:b] = (d[a + 1:b] - x0 * u[a + 1:b]) / u[a]
        return o

    def max_step(self, u, du):
        al = np.inf
        nl = self.nl
        neg = du[:nl] < 0
        if neg.any():
            al = min(al, (-u[:nl][neg] / du[:nl][neg]).min())
        for i in range(len(self.q)):
            a, b = self.off[i], self.off[i + 1]
            al = min(al, _soc_step(u[None, a:b], du[None, a:b]))
        return al

    def scaling(self, s, z):
        nl = self.nl
        W = dict(d=z[:nl] / s[:nl], wl=np.sqrt(s[:nl] / z[:nl]), eta=[], wbar=[])
        lam = np.empty(self.dim)
        lam[:nl] = np.sqrt(s[:nl] * z[:nl])
        for i in range(len(self.q)):
            a, b = self.off[i], self.off[i + 1]
            ss, zz = s[a:b], z[a:b]
            sn = np.sqrt(ss[0] ** 2 - ss[1:] @ ss[1:])
            zn = np.sqrt(zz[0] ** 2 - zz[1:] @ zz[1:])
            sb, zb = ss / sn, zz / zn
            gam = np.sqrt((1 + sb @ zb) / 2)
            wb = (sb + np.concatenate([[zb[0]], -zb[1:]])) / (2 * gam)
            W["eta"].append(np.sqrt(sn / zn))
            W["wbar"].append(wb)
        self.W = W
        lam[nl:] = self.Wmul(z)[nl:]
        return lam

    def Wmul(self, v, inv=False):
        """W v (or W^-1 v); W symmetric"""
        W = self.W
        o = np.empty(self.dim)
        nl = self.nl
        o[:nl] = v[:nl] / W["wl"] if inv else v[:nl] * W["wl"]
        for i in range(len(self.q)):
            a, b = self.off[i], self.off[i + 1]
            wb, eta = W["wbar"][i], W["eta"][i]
            w0, w1 = wb[0], (-wb[1:] if inv else wb[1:])
            u0 = w0 * v[a] + w1 @ v[a + 1:b]
            u1 = v[a] * w1 + v[a + 1:b] + w1 * (w1 @ v[a + 1:b]) / (1 + w0)
            sc = 1 / eta if inv else eta
            o[a] = sc * u0
            o[a + 1:b] = sc * u1
        return o

    def hessian(self, G):
        """G' W^-2 G"""
        W = self.W
        nl = self.nl
        H = G[:nl].T @ (W["d"][:, None] * G[:nl])
        for i in range(len(self.q)):
            a, b = self.off[i], self.off[i + 1]
            Gk = G[a:b]
            wb, eta = W["wbar"][i], W["eta"][i]
            v = np.concatenate([[wb[0]], -wb[1:]])
            u = Gk.T @ v
            J = np.ones(b - a)
            J[1:] = -1
            H += (2 * np.outer(u, u) - Gk.T @ (J[:, None] * Gk)) / eta ** 2
        return H

    def W2inv_mul(self, v):
        return self.Wmul(self.Wmul(v, inv=True), inv=True)


def conelp(c, G, h, cone: Cone, feastol=1e-9, abstol=1e-10, reltol=1e-9, maxit=100, verbose=False):
    nv = c.size
    x = np.zeros(nv)
    s = cone.e()
    z = cone.e()
    tau = kap = 1.0
    e = cone.e()
    if INIT == "ls":
        # CVXOPT's start: x = argmin ||G x - h||, s = h - G x shifted into the cone; z = argmin ||z|| s.t. G'z + c = 0, shifted
        GtG = G.T @ G
        cf0 = sla.cho_factor(GtG + 1e-12 * np.trace(GtG) / nv * np.eye(nv))
        x = sla.cho_solve(cf0, G.T @ h)
        s = h - G @ x
        z = -G @ sla.cho_solve(cf0, c)
        def shift(v):
            m_ = v[:cone.nl].min() if cone.nl else np.inf
            for i in range(len(cone.q)):
                a_, b_ = cone.off[i], cone.off[i + 1]
                m_ = min(m_, v[a_] - np.linalg.norm(v[a_ + 1:b_]))
            return v + (1.0 - m_) * e if m_ < 0 else (v + e if m_ < 1e-8 else v)
        s, z = shift(s), shift(z)
    nrm_h, nrm_c = max(1.0, np.linalg.norm(h)), max(1.0, np.linalg.norm(c))
    status = "maxit"
    hist = []
    for it in range(maxit):
        rx = -G.T @ z - c * tau
        rz = s + G @ x - h * tau
        rt = kap + c @ x + h @ z
        mu = (s @ z + kap * tau) / (cone.deg + 1)
        pcost, dcost = c @ x / tau, -(h @ z) / tau
        pres = np.linalg.norm(rz) / tau / nrm_h
        dres = np.linalg.norm(rx) / tau / nrm_c
        gap = s @ z / tau ** 2
        relgap = gap / max(abs(pcost), abs(dcost), 1e-300)
        if verbose:
            print(f"{it:3d} pcost {pcost:+.10e} dcost {dcost:+.10e} gap {gap:.2e} pres {pres:.1e} dres {dres:.1e} "
                  f"tau {tau:.2e} kap {kap:.2e} mu {mu:.2e}")
        hist.append((pcost, dcost, gap, pres, dres))
        if pres <= feastol and dres <= 100 * feastol and (gap <= abstol or relgap <= reltol):
            status = "optimal"
            break
        hz = h @ z
        if hz < 0 and np.linalg.norm(G.T @ z) / (-hz) <= feastol:
            status = "primal infeasible"
            break
        cx = c @ x
        if cx < 0 and np.linalg.norm(G @ x + s) / (-cx) <= feastol:
            status = "dual infeasible"
            break
        lam = cone.scaling(s, z)
        H = cone.hessian(G)
        reg = REG * np.abs(np.diag(H)).max()
        try:
            cf = sla.cho_factor(H + reg * np.eye(nv))
        except np.linalg.LinAlgError:
            cf = sla.cho_factor(H + 1e-9 * np.abs(np.diag(H)).max() * np.eye(nv))

        def ksolve(bx, bz, refine=2):
            rhs = bx + G.T @ cone.W2inv_mul(bz)
            ux = sla.cho_solve(cf, rhs)
            uz = cone.W2inv_mul(G @ ux - bz)
            for _ in range(refine):
                # refinement on the un-reduced system: the residual of G'uz = bx is evaluated with uz (size of the multipliers),
                # not through H (size 1/slack^2)
                ex = bx - G.T @ uz
                dx = sla.cho_solve(cf, ex)
                ux += dx
                uz += cone.W2inv_mul(G @ dx)
            return ux, uz

        x1, z1 = ksolve(-c, h)
        den = c @ x1 + h @ z1 - kap / tau

        def direction(dx_, dz_, dt_, ds_, dk_):
            lds = cone.div(lam, ds_)
            dzp = dz_ - cone.Wmul(lds)
            x2, z2 = ksolve(-dx_, dzp)
            dtau = (dt_ - dk_ / tau - c @ x2 - h @ z2) / den
            Dx = x2 + dtau * x1
            Dz = z2 + dtau * z1
            Ds = cone.Wmul(lds - cone.Wmul(Dz))
            Dk = (dk_ - kap * dtau) / tau
            return Dx, Dz, dtau, Ds, Dk

        def maxstep(Ds, Dz, dtau, Dk):
            al = min(cone.max_step(s, Ds), cone.max_step(z, Dz))
            if dtau < 0:
                al = min(al, -tau / dtau)
            if Dk < 0:
                al = min(al, -kap / Dk)
            return al

        Dx, Dz, dtau, Ds, Dk = direction(-rx, -rz, -rt, -cone.prod(lam, lam), -kap * tau)
        al = min(1.0, maxstep(Ds, Dz, dtau, Dk))
        sig = (1 - al) ** SIGPOW
        ds_c = -cone.prod(lam, lam) - cone.prod(cone.Wmul(Ds, inv=True), cone.Wmul(Dz)) + sig * mu * e
        dk_c = -kap * tau - Dk * dtau + sig * mu
        Dx, Dz, dtau, Ds, Dk = direction(-(1 - sig) * rx, -(1 - sig) * rz, -(1 - sig) * rt, ds_c, dk_c)
        al = min(1.0, STEP * maxstep(Ds, Dz, dtau, Dk))
        x += al * Dx
        s += al * Ds
        z += al * Dz
        tau += al * dtau
        kap += al * Dk
    return dict(status=status, x=x / tau, s=s / tau, z=z / tau, iters=it, pcost=c @ x / tau, dcost=-(h @ z) / tau, hist=hist)


def fir_ap_conic(p, A, cones=True):
    """fir_ap_cvx problem (oracle.fir_problems.build_fir_ap) as c, G, h, cone; variables [x (2n-1); ripple_stop]"""
    n = p["n"]
    m = A.shape[0]
    nv = 2 * n
    st = p["stop"]
    rows = [np.hstack([A, np.zeros((m, 1))]), np.hstack([-A, np.zeros((m, 1))]),
            np.hstack([A[st], -np.ones((st.size, 1))])]
    rhs = [p["hi"], -p["lo"], np.zeros(st.size)]
    qd = []
    if cones:
        b = np.zeros((2, nv)); b[0, 0] = 1; b[1, 0] = -1
        rows.append(b); rhs.append(np.full(2, p["radius"][0]))
    nl = sum(r.shape[0] for r in rows)
    if cones:
        for i in range(1, n):
            blk = np.zeros((3, nv))
            blk[1, i] = -1
            blk[2, n + i - 1] = -1
            rows.append(blk); rhs.append(np.array([p["radius"][i], 0, 0]))
            qd.append(3)
    return p["c"].copy(), np.vstack(rows), np.concatenate(rhs), Cone(nl, qd)


if __name__ == "__main__":
    sys.path.insert(0, ".")
    from oracle.fir_problems import H1_DUALBAND, build_fir_ap, matrix_fir_ap, violation_fir_ap
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    obj = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
    peak = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-2
    cones = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
    spec = H1_DUALBAND
    scale = 256 / n
    f = np.clip(np.array(spec["f"]) * scale, -1, 1)
    p = build_fir_ap(n, f, spec["a"], spec["d"], obj, peak)
    A = matrix_fir_ap(p["w"], n)
    c, G, h, cone = fir_ap_conic(p, A, cones)
    t0 = time.time()
    r = conelp(c, G, h, cone, verbose=True)
    print(r["status"], r["iters"], "pcost", r["pcost"], "dcost", r["dcost"], "viol", violation_fir_ap(p, r["x"]) if cones else None,
          f"{time.time() - t0:.1f}s")
