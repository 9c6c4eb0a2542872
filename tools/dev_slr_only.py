import os, sys, numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
lib = m.lib()
rf = np.load("tests/golden/pulses.npz")["rf512_rad"]
ns, nx = rf.size, 1_000_000
T = lambda a: torch.tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda")
d_rfr, d_rfi, d_g, d_x = T(rf.real), T(rf.imag), T(np.full(ns, 2 * np.pi / ns)), T(np.linspace(-40, 40, nx))
ab = torch.empty((4, nx), dtype=torch.float64, device="cuda")
ws = torch.empty(int(lib.mbrf_abr_workspace_bytes(ns)), dtype=torch.uint8, device="cuda")
for _ in range(3):
    m._lib.check(lib.mbrf_abr_device(d_rfr.data_ptr(), d_rfi.data_ptr(), d_g.data_ptr(), None, ns, d_x.data_ptr(), nx, None, 1, 0, 0, nx,
                                     ab[0].data_ptr(), ab[1].data_ptr(), ab[2].data_ptr(), ab[3].data_ptr(), ws.data_ptr(), None))
torch.cuda.synchronize()
print("ok")
