import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiband_rf_pulse_design_b200 as m
lib = m.lib(); ms = C.c_float()
lib.mbrf_ipm_cholesky_bench(512, 512, 1, 2, C.byref(ms)); print("cholesky dd 512 x 512:", ms.value, "ms")
lib.mbrf_ipm_cholesky_bench(512, 512, 0, 2, C.byref(ms)); print("cholesky fp64 512 x 512:", ms.value, "ms")
