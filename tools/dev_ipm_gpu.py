#!/usr/bin/env python
"""GPU bring-up of the interior-point solver: golden cases of tests/golden/fir_ap_known.json and the N = 256 weights."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import H1_DUALBAND, build_fir_ap, violation_fir_ap

K = json.load(open("tests/golden/fir_ap_known.json"))
lib = m.lib()
args = sys.argv[1:]
prec = int(args[0]) if args else 2
verbose = int(args[1]) if len(args) > 1 else 0
big = (args[2] != "0") if len(args) > 2 else True
lib.mbrf_ipm_set_option(0, float(prec))
lib.mbrf_ipm_set_option(3, float(verbose))
for name, k in K.items():
    if "n" not in k or "obj" not in k or k["n"] > 64:
        continue
    t0 = time.time()
    hs, st, ex = m.fir_ap_cvx_batch(k["n"], [k["f"]], k["a"], k["d"], [k["obj"]], [k["peak"]], return_info=True, method="ipm")
    dt = time.time() - t0
    info = ex["info"][0]
    want = k.get("outer_obj") if k.get("outer_status", 1) == 0 else None
    p = build_fir_ap(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"])
    viol = violation_fir_ap(p, np.concatenate([ex["x"][0], [ex["ripple_stop"][0]]])) if st[0] == "Solved" else float("nan")
    print(f"{name:32s} {st[0]:7s} it {int(info[1]):3d} obj {info[2]:.10f} want {want} inner {k.get('inner_obj')} viol {viol:.1e} ({info[4]:.1e}) {dt:.2f}s", flush=True)
if big:
    known = {0.1: 0.01603297372525376, 4.0: 0.017046095660090497, 10.0: 0.018290282454801167, 100.0: 0.02649552520674127,
             1e4: 0.40001113382795833, 1e5: 3.0110470878389908}
    objs = [float(v) for v in os.environ["OBJS"].split(",")] if os.environ.get("OBJS") else list(known)
    peak = float(os.environ.get("PEAK", "1e-2"))
    t0 = time.time()
    hs, st, ex = m.fir_ap_cvx_batch(256, [H1_DUALBAND["f"]] * len(objs), H1_DUALBAND["a"], H1_DUALBAND["d"], objs, [peak] * len(objs),
                                    return_info=True, method="ipm")
    dt = time.time() - t0
    for i, o in enumerate(objs):
        info = ex["info"][i]
        print(f"N=256 obj {o:8g} {st[i]} it {int(info[1])} obj {info[2]:.10f} highs {known.get(o, float("nan")):.10f} rel {(info[2]-known.get(o, float("nan")))/known.get(o, float("nan")):+.2e} viol {info[4]:.1e} dres {info[5]:.1e}")
    print(f"batch of {len(objs)}: {dt:.2f}s")
