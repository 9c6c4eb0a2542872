#!/usr/bin/env python
"""Known answer for BASELINE config 3: fir_qp_cvx(256, f, a, d, 120, 1e6) on the dual-band H-1 spec, reference grid
(oversamp 10: 2566 points), from an INDEPENDENT CPU solver: SciPy trust-constr on the smooth squared-norm form with epigraph
variables (oracle/fir_problems.py: solve_fir_qp_reference), objective divided by obj for conditioning.  Takes tens of minutes.

    python tests/golden/make_golden_fir_qp_n256.py      -> tests/golden/fir_qp_n256_known.json
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.fir_problems import H1_DUALBAND as S, build_fir_qp, objective_fir_qp, violation_fir_qp  # noqa: E402
from oracle import fir_problems as F  # noqa: E402

n, k, obj = 256, 120.0, 1e6
p = build_fir_qp(n, S["f"], S["a"], S["d"], k, obj)
ps = dict(p, obj=1.0)                       # trust-constr sees  E/obj + Peak  (same minimiser)
t0 = time.time()


def scaled_solve(p1, scale, maxiter):
    """min E*scale + P  by temporarily re-weighting: objective_fir_qp is E + obj*P, so solve with obj = 1/scale and rescale"""
    q = dict(p1, obj=1.0 / scale)           # E + (1/scale) P  ==  (1/scale) (scale E + P)
    return F.solve_fir_qp_reference(q, maxiter=maxiter)


res = scaled_solve(p, 1.0 / obj, int(sys.argv[1]) if len(sys.argv) > 1 else 4000)
x = res.x[:2 * n]
rec = dict(n=n, k=k, obj=obj, f=list(map(float, S["f"])), a=list(map(float, S["a"])), d=list(map(float, S["d"])), rows=int(p["w"].size),
           objective=float(objective_fir_qp(p, x)), violation=float(violation_fir_qp(p, x)), energy=float(np.linalg.norm(x)),
           peak=float(np.hypot(x[:n], x[n:]).max()), status=int(res.status), message=str(res.message), iterations=int(res.nit),
           seconds=round(time.time() - t0, 1))
print(rec, flush=True)
json.dump({"qp_n256_obj1e6": rec}, open(os.path.join(HERE, "fir_qp_n256_known.json"), "w"), indent=1)
