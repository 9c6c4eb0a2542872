import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from multiband_rf_pulse_design_b200 import fir
H1 = dict(f=[-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006], a=[0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886], d=[0.014436, 0.022361, 0.017683])
objs = np.logspace(-2, 1, 64); peaks = np.logspace(-3.2, -2, 8)
if len(sys.argv) > 1:    # single design trace: obj index, peak index
    o, pk = objs[int(sys.argv[1])], peaks[int(sys.argv[2])]
    os.environ["MBRF_PDHG_TRACE"] = "1"
    hs, st, ex = fir.fir_ap_cvx_batch(256, [H1["f"]], H1["a"], H1["d"], [o], [pk], return_info=True, max_iter=60000)
    print(o, pk, st, ex["info"])
elif os.environ.get("BIG"):
    objs = np.logspace(-2, 1, 512)
    r = fir.fir_ap_cvx_sweep(256, H1["f"], H1["a"], H1["d"], objs, peaks[:1], [0.0], max_iter=60000)
    info = r["info"]
    bad = np.nonzero(info[:, 0] != 1)[0]
    print("unsolved obj idx:", bad.tolist(), "obj values", objs[bad].tolist())
    print("iters: mean", info[:, 1].mean(), "p90", np.percentile(info[:, 1], 90), "max", info[:, 1].max())
    for b in bad[:2]:
        os.environ["MBRF_PDHG_TRACE"] = "1"
        hs, st, ex = fir.fir_ap_cvx_batch(256, [H1["f"]], H1["a"], H1["d"], [objs[b]], [peaks[0]], return_info=True, max_iter=60000)
        print("single", objs[b], st, ex["info"])
else:
    r = fir.fir_ap_cvx_sweep(256, H1["f"], H1["a"], H1["d"], objs, peaks, [0.0], max_iter=60000)
    info = r["info"].reshape(64, 8, 8)
    bad = np.argwhere(info[:, :, 0] != 1)
    print("unsolved (obj idx, peak idx):", bad.tolist())
    it = info[:, :, 1]
    print("iters/1000 by obj idx (max over peaks):", np.round(it.max(1) / 1000, 1).tolist())
