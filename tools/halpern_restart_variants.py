"""CPU prototype (numpy twin of the solver): reflected Halpern PDHG with different restart parameters / restart tests on hard
(large stop-band weight) designs.  Developer experiment for DESIGN.md 7b item 1; nothing here ships.
usage: python tools/halpern_restart_variants.py [n [number of variants to run]]"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.fir_problems import build_fir_ap
from oracle import pdhg_reference as R

def solve_h(q, max_iter=80000, check_every=64, eps_pr=8e-7, eps_dr=1e-4, eps_gap=5e-5, crit="kkt", b_suff=0.2, b_nec=0.8, b_art=0.36, omega_theta=0.5, min_since=1, want_z=False):
    K = q["K"]; M, N = K.shape; B = q["c"].shape[1]
    rng = np.random.default_rng(0); v = rng.normal(size=N)
    for _ in range(60):
        v = K.T @ (K @ v); nk = np.linalg.norm(v); v /= nk
    eta = 0.9 / np.sqrt(nk)
    omega = np.ones(B)
    z = R.proj_X(np.zeros((N, B)), q); y = np.zeros((M, B))
    z0, y0 = z.copy(), y.copy()
    last_err = np.full(B, np.inf); prev_err = np.full(B, np.inf)
    status = np.zeros(B, int); done = np.zeros(B, int); since = np.zeros(B); tot = 0
    zsol = np.zeros((N, B)); ysol = np.zeros((M, B))
    s0, ns = q["srow0"], q["ns"]
    def metrics(zz, yy):
        Kz = K @ zz
        Kr, lo_r, hi_r, y_r = Kz[:s0], q["lo"][:s0], q["hi"][:s0], yy[:s0]
        pr = np.maximum(np.maximum(Kr - hi_r, lo_r - Kr), 0).max(0)
        g = q["c"] + K.T @ yy
        dr = np.abs(zz - R.proj_X(zz - g, q)).max(0)
        tmax = np.maximum(Kz[s0:s0 + ns].max(0), 0.0)
        pobj = (q["c"] * zz).sum(0) + q["sw"] * tmax
        yp, ym = np.maximum(y_r, 0), np.maximum(-y_r, 0)
        with np.errstate(invalid="ignore"):
            hs = np.where(yp > 0, hi_r * yp, 0.0).sum(0) - np.where(ym > 0, lo_r * ym, 0.0).sum(0)
        return pr, dr, pobj, -hs + (g * zz).sum(0)
    for it in range(1, max_iter + 1):
        tau, sig = eta / omega, eta * omega
        zh = R.proj_X(z - tau * (q["c"] + K.T @ y), q)
        vv = y + sig * (K @ (2 * zh - z)); wv = vv / sig
        with np.errstate(invalid="ignore"):
            yh = np.where(wv > q["hi"], vv - sig * q["hi"], np.where(wv < q["lo"], vv - sig * q["lo"], 0.0))
        yh[s0:s0 + ns] = R.proj_simplex(vv[s0:s0 + ns], q["sw"])
        fp = np.sqrt(omega * ((zh - z) ** 2).sum(0) + ((yh - y) ** 2).sum(0) / omega)   # fixed-point residual in the omega norm
        rho = (since + 1) / (since + 2)
        z = rho * (2 * zh - z) + (1 - rho) * z0
        y = rho * (2 * yh - y) + (1 - rho) * y0
        since = since + 1; tot += 1
        if it % check_every:
            continue
        pa, ra, oa, da = metrics(zh, yh)
        pc, rc, oc, dc = metrics(z, y)
        ea = np.maximum(np.maximum(pa, ra), np.abs(oa - da)); ec = np.maximum(np.maximum(pc, rc), np.abs(oc - dc))
        ec = np.where(np.isfinite(ec), ec, np.inf)
        use = ea < ec
        cz, cy = np.where(use, zh, z), np.where(use, yh, y)
        ce = np.where(use, ea, ec); cp = np.where(use, pa, pc); cr = np.where(use, ra, rc); co = np.where(use, oa, oc); cd = np.where(use, da, dc)
        solved = (cp <= eps_pr) & (cr <= eps_dr) & (np.abs(co - cd) <= eps_gap * np.maximum(np.abs(co), 1e-12)) & (status == 0)
        status[solved] = 1; done[solved] = it; zsol[:, solved] = cz[:, solved]; ysol[:, solved] = cy[:, solved]
        if (status != 0).all():
            break
        me = fp if crit == "fp" else ce
        do = ((me <= b_suff * last_err) | ((me <= b_nec * last_err) & (me > prev_err)) | (since >= b_art * tot)) & (since >= min_since)
        prev_err = me
        if do.any():
            dz = np.linalg.norm(cz - z0, axis=0); dy = np.linalg.norm(cy - y0, axis=0)
            ok = do & (dz > 1e-12) & (dy > 1e-12)
            omega = np.where(ok, np.exp(omega_theta * np.log(np.maximum(dy, 1e-300) / np.maximum(dz, 1e-300)) + (1 - omega_theta) * np.log(omega)), omega)
            z = np.where(do, cz, z); y = np.where(do, cy, y)
            z0 = np.where(do, z, z0); y0 = np.where(do, y, y0)
            # the fixed-point residual at the restart point is what later residuals are compared with
            last_err = np.where(do, me, last_err); since = np.where(do, 0.0, since)
    done[status == 0] = tot
    if want_z:
        return (done, status, zsol, ysol) if want_z == 2 else (done, status, zsol)
    return done, status

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    f = [-0.6, -0.35, -0.2, 0.18, 0.38, 0.6]; a = [0.866, 0.866, 0, 0, 0.707, 0.707]; d = [0.02, 0.03, 0.025]
    objs = [1.0, 3.0, 6.0, 10.0, 10.0, 10.0]; peaks = [10**-1.5]*4 + [10**-1.8, 10**-2.0]
    q = R.assemble_fir_ap([build_fir_ap(n, f, a, d, o, pk) for o, pk in zip(objs, peaks)])
    for name, kw in [("kkt default", {}), ("nec 0.9", dict(b_nec=0.9)), ("nec 0.9 min 256", dict(b_nec=0.9, min_since=256)),
                     ("nec 0.95 min 256", dict(b_nec=0.95, min_since=256)), ("min 256", dict(min_since=256)), ("fp criterion", dict(crit="fp")), ("kkt b_art 0.2", dict(b_art=0.2)), ("kkt b_art 0.6", dict(b_art=0.6)),
                     ("kkt suff 0.1 nec 0.9", dict(b_suff=0.1, b_nec=0.9)), ("kkt check 32", dict(check_every=32)), ("kkt theta 0.2", dict(omega_theta=0.2))][:int(sys.argv[2]) if len(sys.argv) > 2 else None]:
        t = time.time(); it, st = solve_h(q, **kw)
        print(f"{name:22s}", it, st, f"{time.time()-t:.0f} s", flush=True)


if __name__ == "__main__":
    main()
