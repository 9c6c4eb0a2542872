"""How far from the optimum is the solver's answer when its termination test fires, as a function of the stop-band weight?
numpy twin (tools/halpern_restart_variants.py:solve_h, same test as the GPU solver) against HiGHS on the same fir_ap_cvx
problem (oracle/fir_problems.py:solve_fir_ap_highs; Peak chosen so that the peak cones are inactive, which makes the cone-free
LP exact).  Developer experiment for DESIGN.md 6 "known weak spot"; nothing here ships.
usage: python tools/weight_certificate_study.py [n [h1]]      (h1: the bench's dual-band H-1 spec, meant for n = 256)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from oracle.fir_problems import build_fir_ap, solve_fir_ap_highs, violation_fir_ap, matrix_fir_ap
from oracle import pdhg_reference as R
from halpern_restart_variants import solve_h


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    f = [-0.6, -0.35, -0.2, 0.18, 0.38, 0.6]; a = [0.866, 0.866, 0, 0, 0.707, 0.707]; d = [0.02, 0.03, 0.025]
    weights = [0.1, 1.0, 4.0, 10.0, 30.0, 100.0]
    peak = 10 ** -1.0
    variants = [("default tolerances", {}), ("eps_dr 1e-5", dict(eps_dr=1e-5)), ("eps_dr 1e-6", dict(eps_dr=1e-6)),
                ("eps_gap 5e-6", dict(eps_gap=5e-6))]
    if len(sys.argv) > 2 and sys.argv[2] == "h1":       # bench.py:H1_DUALBAND (specsat_H1_dualband.m after dzrf_mb's shift)
        f = [-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006]
        a = [0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886]; d = [0.014436, 0.022361, 0.017683]
        weights = [1.0, 4.0, 10.0]
        peak = 10 ** -2.0
        variants = variants[:1]
    probs = [build_fir_ap(n, f, a, d, w, peak) for w in weights]
    q = R.assemble_fir_ap(probs)
    for label, kw in variants:
        its, st, z = solve_h(q, max_iter=120000, want_z=True, **kw)
        print(label)
        for j, (w, p) in enumerate(zip(weights, probs)):
            x = z[:2 * n - 1, j] / q["colscale"][:2 * n - 1]      # the solver works on unit-norm columns
            S = matrix_fir_ap(p["w"], n) @ x
            t = max(S[p["stop"]].max(), 0.0)
            obj = x[0] + w * t
            ref, _ = solve_fir_ap_highs(p)
            viol = violation_fir_ap(p, np.concatenate([x, [t]]))
            slack = (p["radius"][1:] - np.hypot(x[1:n], x[n:2 * n - 1])).min()      # > 0: every peak cone inactive, the LP is exact
            print(f"  weight {w:6.1f}: {its[j]:6d} iterations, status {st[j]}, objective {obj:.8f}  HiGHS {ref.fun:.8f}  "
                  f"rel diff {abs(obj - ref.fun) / abs(ref.fun):.2e}  violation {viol:.1e}  min cone slack {slack:.1e}", flush=True)


if __name__ == "__main__":
    main()
