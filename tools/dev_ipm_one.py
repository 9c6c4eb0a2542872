import sys
import numpy as np
sys.path.insert(0, ".")
from multiband_rf_pulse_design_b200 import fir
from oracle.fir_problems import H1_DUALBAND as S
fir.fir_ap_cvx_batch(256, [S["f"]], S["a"], S["d"], [0.1], [1.0])
