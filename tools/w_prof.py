import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiband_rf_pulse_design_b200 import fir
from bench import H1_DUALBAND
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
designs = [fir.assemble_fir_ap(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 5.0 + 0.1 * b, 10 ** -2.5) for b in range(B)]
fir._solve_batch_ap(256, designs, max_iter=128)
fir._solve_batch_ap(256, designs, max_iter=256, eps_pr=1e-30)
