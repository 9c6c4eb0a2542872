"""CPU prototype (numpy, on top of the solver's numpy twin oracle/pdhg_reference.py): restarted PDHG with running averages
(what the GPU runs) vs reflected Halpern PDHG (r2HPDHG, Lu & Yang 2024) with the same restart test and primal-weight rule.
Prints iterations to the solver's tolerances per design.  Developer experiment for DESIGN.md 7b item 1; nothing here ships."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.fir_problems import build_fir_ap
from oracle import pdhg_reference as R


def solve_halpern(q, max_iter=60000, check_every=64, eps_pr=8e-7, eps_dr=1e-4, eps_gap=5e-5, reflect=1.0):
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
        hs = np.where(yp > 0, hi_r * yp, 0.0).sum(0) - np.where(ym > 0, lo_r * ym, 0.0).sum(0)
        return pr, dr, pobj, -hs + (g * zz).sum(0)

    for it in range(1, max_iter + 1):
        tau, sig = eta / omega, eta * omega
        zh = R.proj_X(z - tau * (q["c"] + K.T @ y), q)                     # T(z, y)
        vv = y + sig * (K @ (2 * zh - z)); wv = vv / sig
        with np.errstate(invalid="ignore"):
            yh = np.where(wv > q["hi"], vv - sig * q["hi"], np.where(wv < q["lo"], vv - sig * q["lo"], 0.0))
        yh[s0:s0 + ns] = R.proj_simplex(vv[s0:s0 + ns], q["sw"])
        k = since
        rho = (k + 1) / (k + 2)
        z = rho * ((1 + reflect) * zh - reflect * z) + (1 - rho) * z0      # reflected Halpern step
        y = rho * ((1 + reflect) * yh - reflect * y) + (1 - rho) * y0
        since = since + 1; tot += 1
        if it % check_every:
            continue
        pa, ra, oa, da = metrics(zh, yh)                                   # candidate 0: the PDHG output
        pc, rc, oc, dc = metrics(R.proj_X(z.copy(), q), y)                 # candidate 1: the Halpern iterate
        ea = np.maximum(np.maximum(pa, ra), np.abs(oa - da)); ec = np.maximum(np.maximum(pc, rc), np.abs(oc - dc))
        use = ea < ec
        cz, cy = np.where(use, zh, z), np.where(use, yh, y)
        ce = np.where(use, ea, ec); cp = np.where(use, pa, pc); cr = np.where(use, ra, rc); co = np.where(use, oa, oc); cd = np.where(use, da, dc)
        solved = (cp <= eps_pr) & (cr <= eps_dr) & (np.abs(co - cd) <= eps_gap * np.maximum(np.abs(co), 1e-12)) & (status == 0)
        status[solved] = 1; done[solved] = it
        if (status != 0).all():
            break
        do = (ce <= 0.2 * last_err) | ((ce <= 0.8 * last_err) & (ce > prev_err)) | (since >= 0.36 * tot)
        prev_err = ce
        if do.any():
            dz = np.linalg.norm(cz - z0, axis=0); dy = np.linalg.norm(cy - y0, axis=0)
            ok = do & (dz > 1e-12) & (dy > 1e-12)
            omega = np.where(ok, np.exp(0.5 * np.log(np.maximum(dy, 1e-300) / np.maximum(dz, 1e-300)) + 0.5 * np.log(omega)), omega)
            z = np.where(do, cz, z); y = np.where(do, cy, y)
            z0 = np.where(do, z, z0); y0 = np.where(do, y, y0)
            last_err = np.where(do, ce, last_err); since = np.where(do, 0.0, since)
    done[status == 0] = tot
    return dict(status=status, iters=done)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    f = [-0.6, -0.35, -0.2, 0.18, 0.38, 0.6]; a = [0.866, 0.866, 0, 0, 0.707, 0.707]; d = [0.02, 0.03, 0.025]
    objs = [0.03, 0.3, 1.0, 3.0, 6.0, 10.0]
    probs = [build_fir_ap(n, f, a, d, o, 10 ** -1.5) for o in objs]
    q = R.assemble_fir_ap(probs)
    t = time.time(); r0 = R.solve(q, max_iter=60000); t0 = time.time() - t
    print("averaged restarts :", r0["iters"], r0["status"], f"{t0:.0f} s", flush=True)
    for refl in (1.0, 0.0):
        t = time.time(); r1 = solve_halpern(q, max_iter=60000, reflect=refl); t1 = time.time() - t
        print(f"Halpern (reflect {refl}):", r1["iters"], r1["status"], f"{t1:.0f} s", flush=True)
