#!/usr/bin/env python
"""Known answers for fir_qp_cvx (scalar-obj SOCP, fir_qp_cvx.m:145-166) from an independent CPU solver:
SciPy trust-constr on the epigraph form with squared-norm constraints (oracle/fir_problems.py).  ~1 min per case."""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.fir_problems import build_fir_qp, objective_fir_qp, solve_fir_qp_reference, violation_fir_qp  # noqa: E402

SPEC = dict(f=[-0.6, -0.25, 0.1, 0.45], a=[1, 1, 0.5, 0.5], d=[0.05, 0.05])
CASES = {"qp_n16_obj1": (16, 2.0, 1.0), "qp_n16_obj0": (16, 2.0, 0.0), "qp_n12_obj3": (12, 1.5, 3.0)}
out = {}
for name, (n, k, obj) in CASES.items():
    t0 = time.time()
    p = build_fir_qp(n, SPEC["f"], SPEC["a"], SPEC["d"], k, obj)
    r = solve_fir_qp_reference(p)
    x = r.x[:2 * n]
    assert r.status in (1, 2) and violation_fir_qp(p, x) < 1e-7, (name, r.status)   # converged and feasible, or it is no anchor
    out[name] = dict(n=n, k=k, obj=obj, f=SPEC["f"], a=SPEC["a"], d=SPEC["d"], rows=int(p["w"].size), status=int(r.status),
                     objective=float(objective_fir_qp(p, x)), epigraph_objective=float(r.fun),
                     violation=float(violation_fir_qp(p, x)), seconds=round(time.time() - t0, 1))
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(HERE, "fir_qp_known.json"), "w"), indent=1)
