"""Effect of the initial primal weight on the cold sweep: iterations and time for omega0 in argv."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND
n = 256
objs = np.logspace(-2, 1, 64); peaks = np.logspace(-3.2, -2, 8)
fl, ol, pl = fir.sweep_grid(H1_DUALBAND["f"], objs, peaks, [0.0])
designs = [fir.assemble_fir_ap(n, fl[i], H1_DUALBAND["a"], H1_DUALBAND["d"], ol[i], pl[i]) for i in range(512)]
fir._solve_batch_ap(n, designs[:64], max_iter=256)
ref = None
for om in [float(v) for v in sys.argv[1:]] or [1.0, 10.0, 100.0]:
    t = time.perf_counter()
    x, ts, info = fir._solve_batch_ap(n, designs, max_iter=int(os.environ.get('MAXIT', '60000')), warm=(None, None, np.full(512, om)), eps_dr=float(os.environ.get('EPS_DR', '1e-4')))
    dt = time.perf_counter() - t
    if ref is None:
        ref = info
    print(f"omega0 {om:7.1f}: {dt:.2f} s, solved {(info[:,0]==1).sum()}, iters mean {info[:,1].mean():.0f} max {info[:,1].max():.0f}, "
          f"per-design rel obj diff: max {(np.abs(info[:,2]-ref[:,2])/np.abs(ref[:,2])).max():.1e} median {np.median(np.abs(info[:,2]-ref[:,2])/np.abs(ref[:,2])):.1e}"
          f" argmax {int(np.argmax(np.abs(info[:,2]-ref[:,2])/np.abs(ref[:,2])))} obj {ref[int(np.argmax(np.abs(info[:,2]-ref[:,2])/np.abs(ref[:,2]))),2]:.6f} vs {info[int(np.argmax(np.abs(info[:,2]-ref[:,2])/np.abs(ref[:,2]))),2]:.6f}"
          f" duals {ref[int(np.argmax(np.abs(info[:,2]-ref[:,2])/np.abs(ref[:,2]))),3]:.6f} {info[int(np.argmax(np.abs(info[:,2]-ref[:,2])/np.abs(ref[:,2]))),3]:.6f}", flush=True)
