"""numpy prototype for DESIGN.md 7b item 1: large stop-band weights through fixed-ripple LPs.
F(t) = x0*(t) + w t with x0*(t) = min x0 s.t. stop rows <= t is convex and F'(t) = w - lambda(t) (lambda = sum of the stop-row
multipliers), so the weighted problem is a 1-D root search over weight-free LPs, which the solver converges quickly and with a
tight test.  Phase 1: the weighted problem at loose tolerances -> t_hat.  Phase 2: LPs at t_hat * (1 + s), s in a small
stencil (one batch: same matrix), interpolate lambda(t) = w, one more LP at that t.  Compared with HiGHS on the weighted problem.
RESULT (n = 64, H-1 dual-band spec scaled to the shorter filter): negative.  The fixed-ripple LPs near the smallest feasible
ripple are as slow as the weighted problem (40-60 k iterations against 10-135 k), i.e. the difficulty is the near-degenerate
geometry at small ripple, not the weight in the objective; w = 10 ends 7e-5 from HiGHS, w = 30 does not get its stencil solved.
Also measured with HiGHS: the ripple has to be within ~1-2 % of the optimal one for a 1e-4 objective (F is that curved).
Developer experiment; nothing here ships.   usage: python tools/two_phase_weight_proto.py [n [weights]]"""
import copy
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from oracle.fir_problems import build_fir_ap, solve_fir_ap_highs, matrix_fir_ap
from oracle import pdhg_reference as R
from halpern_restart_variants import solve_h


def fixed_ripple(p, t):
    p2 = copy.copy(p)
    p2["hi"] = p["hi"].copy()
    p2["hi"][p["stop"]] = np.minimum(p["hi"][p["stop"]], t)
    p2["c"] = p["c"].copy()
    p2["c"][-1] = 0.0
    return p2


def lp_batch(p, ts, n):
    q = R.assemble_fir_ap([fixed_ripple(p, t) for t in ts])
    its, st, z, y = solve_h(q, max_iter=60000, want_z=2)
    x0 = z[0] / q["colscale"][0]
    lam = np.array([np.maximum(y[:q["srow0"], j][p["stop"]], 0.0).sum() for j in range(len(ts))])
    return its, st, x0, lam


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    weights = [float(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4.0, 10.0, 30.0]
    sc = 256.0 / n                                       # bench.py:H1_DUALBAND scaled to the shorter filter
    f = [sc * v for v in [-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006]]
    a = [0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886]; d = [0.014436, 0.022361, 0.017683]
    for w in weights:
        p = build_fir_ap(n, f, a, d, w, 0.01 if n == 256 else 0.1)
        ref = solve_fir_ap_highs(p)[0]
        t0 = time.time()
        q = R.assemble_fir_ap([p])
        it1, st1, z = solve_h(q, max_iter=60000, eps_pr=8e-6, eps_dr=1e-3, eps_gap=5e-4, want_z=True)
        x = z[:2 * n - 1, 0] / q["colscale"][:2 * n - 1]
        t_hat = max((matrix_fir_ap(p["w"], n) @ x)[p["stop"]].max(), 1e-12)
        stencil = t_hat * (1 + np.array([-0.06, -0.02, 0.02, 0.06]))
        it2, st2, x0s, lam = lp_batch(p, stencil, n)
        order = np.argsort(lam)                          # lambda decreases with t: interpolate t at lambda = w
        t_new = float(np.interp(w, lam[order], stencil[order]))
        it3, st3, x03, lam3 = lp_batch(p, [t_new], n)
        F = x03[0] + w * t_new
        print(f"w {w:5.1f}: HiGHS F* {ref.fun:.9f} t* {ref.x[-1]:.5e} | phase 1 {it1[0]} its -> t_hat {t_hat:.5e} | stencil its {it2.max()} "
              f"lambda {np.array2string(lam, precision=2)} -> t {t_new:.5e} | final LP {it3[0]} its lambda {lam3[0]:.2f} F {F:.9f} "
              f"rel excess {(F - ref.fun) / ref.fun:.2e}  ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main()
