#!/usr/bin/env python
"""Known answers for the minimax form of fir_qp_cvx (two-element obj, fir_qp_cvx.m:170-191) from an independent CPU
solver: SciPy trust-constr on the epigraph form with squared-norm constraints (oracle/fir_problems.py)."""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.fir_problems import (build_fir_qp, objective_fir_qp_minimax, solve_fir_qp_minimax_reference,  # noqa: E402
                                 violation_fir_qp_minimax)

SPEC = dict(f=[-0.6, -0.25, 0.1, 0.45], a=[1, 1, 0.5, 0.5], d=[0.05, 0.05])
CASES = {"mm_n16_a": (16, 2.0, [0.5, 1.0]), "mm_n16_b": (16, 2.0, [0.1, 0.0]), "mm_n12_c": (12, 1.5, [1.0, 2.0])}
out = {}
for name, (n, k, obj2) in CASES.items():
    t0 = time.time()
    p = build_fir_qp(n, SPEC["f"], SPEC["a"], SPEC["d"], k, 0.0)
    r = solve_fir_qp_minimax_reference(p, obj2)
    x = r.x[:2 * n]
    assert r.status in (1, 2) and violation_fir_qp_minimax(p, x) < 1e-7, (name, r.status)
    out[name] = dict(n=n, k=k, obj=obj2, f=SPEC["f"], a=SPEC["a"], d=SPEC["d"], rows=int(p["w"].size), status=int(r.status),
                     objective=float(objective_fir_qp_minimax(p, x, obj2)), epigraph_objective=float(r.fun),
                     violation=float(violation_fir_qp_minimax(p, x)), seconds=round(time.time() - t0, 1))
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(HERE, "fir_qp_minimax_known.json"), "w"), indent=1)
