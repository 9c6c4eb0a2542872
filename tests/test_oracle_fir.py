"""CPU tests of the solver-path oracle: problem construction restated from fir_ap_cvx.m, the numpy twin of the
GPU's PDHG, and their agreement with the independent HiGHS known answers (tests/golden/fir_ap_known.json).
Parity with the reference's own solver (CVX) is UNPINNED: it does not exist offline."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

KNOWN = json.load(open(os.path.join(GOLDEN, "fir_ap_known.json")))


def test_product_assembly_equals_oracle_restatement():
    """Host assembly in the product (fir.py) and the oracle's independent restatement build the same problem."""
    from multiband_rf_pulse_design_b200 import fir
    from oracle.fir_problems import H1_DUALBAND, build_fir_ap
    for n, obj, peak in ((256, 0.1, 1e-3), (64, 3.0, 5e-3)):
        a = fir.assemble_fir_ap(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], obj, peak)
        b = build_fir_ap(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], obj, peak)
        assert np.array_equal(a["w"], b["w"]) and np.array_equal(a["lo"], b["lo"]) and np.array_equal(a["hi"], b["hi"])
        assert np.array_equal(np.nonzero(a["stop"])[0], b["stop"]) and np.array_equal(a["radius"], b["radius"])
    k = KNOWN["h1_dualband_n256"]
    assert k["rows"] == 7686 and k["stop_rows"] == 118            # 264 band + 7422 transition rows


def test_known_answer_anchor():
    """the cone-free N=256 LP on the dual-band H-1 spec has objective 0.01603 (HiGHS)."""
    k = KNOWN["h1_dualband_n256"]
    assert k["cone_free_status"] == 0 and abs(k["cone_free_obj"] - 0.0160330) < 2e-7
    assert KNOWN["h1_dualband_n128_infeasible"]["cone_free_status"] != 0
    assert KNOWN["lowpass_n10_infeasible"]["cone_free_status"] == 2


@pytest.mark.parametrize("case", ["lowpass_n24", "lowpass_n24_tightpeak"])
def test_numpy_pdhg_twin_matches_highs(case):
    from oracle import pdhg_reference as P
    from oracle.fir_problems import build_fir_ap, violation_fir_ap
    k = KNOWN[case]
    p = build_fir_ap(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"])
    r = P.solve(P.assemble_fir_ap([p]), max_iter=60000)
    assert r["status"][0] == 1
    z = r["z"][:, 0]
    lo, hi = k["outer_obj"], k["inner_obj"]                        # bounds on the true SOCP optimum
    assert lo * (1 - 1e-4) <= p["c"] @ z <= hi * (1 + 1e-4)
    assert violation_fir_ap(p, z) <= 1e-6


def test_numpy_pdhg_twin_certifies_infeasibility():
    from oracle import pdhg_reference as P
    from oracle.fir_problems import build_fir_ap
    k = KNOWN["lowpass_n24_peak_infeasible"]
    p = build_fir_ap(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"])
    upper = np.array([p["radius"][0] + p["c"][-1] * p["hi"][p["stop"]].max()])
    r = P.solve(P.assemble_fir_ap([p]), max_iter=5000, obj_upper=upper)
    assert r["status"][0] == 2


def test_numpy_qp_twin_matches_scipy_reference():
    """fir_qp_cvx SOCP: the numpy twin of the GPU algorithm against SciPy trust-constr (small design)."""
    from oracle import pdhg_qp_reference as Q
    from oracle.fir_problems import build_fir_qp, objective_fir_qp, violation_fir_qp
    k = json.load(open(os.path.join(GOLDEN, "fir_qp_known.json")))["qp_n16_obj1"]
    p = build_fir_qp(k["n"], k["f"], k["a"], k["d"], k["k"], k["obj"])
    x, _, iters = Q.solve_qp_twin(p)
    assert iters < 100000
    assert abs(objective_fir_qp(p, x) - k["objective"]) <= 1e-4 * k["objective"]
    assert violation_fir_qp(p, x) <= 1e-6


def _minimum_phase_case(n, rad, seed):
    """A filter with all zeros inside |z| <= rad < 1 (angles spread around the circle, so the spectrum has no deep null and
    the cepstrum does not alias on the 8x padded grid) and its autocorrelation: the spectral factor of r is h0 itself."""
    from oracle.fir_post import poly_reference
    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * (np.arange(n - 1) + rng.uniform(-0.3, 0.3, n - 1)) / (n - 1)
    z = rng.uniform(0.2, rad, n - 1) * np.exp(1j * ang)
    h0 = poly_reference(z[rng.permutation(n - 1)])
    h0 = h0 / np.abs(h0).max()
    return h0, np.correlate(h0, h0, mode="full")


MINPHASE_CASES = ((16, 0.7, 1e-13), (33, 0.8, 1e-12), (64, 0.85, 1e-10), (128, 0.85, 1e-9), (256, 0.9, 1e-6))


def test_fmp2_restatement_recovers_known_minimum_phase_filters():
    """Known answer for fmp2 (fir_ap_cvx.m:262-283): the minimum-phase spectral factor is unique, so for a filter whose zeros
    all lie inside the unit circle fmp2(autocorrelation) must return the filter itself -- to rounding where the log-spectrum is
    smooth (measured 1e-16 ... 4e-12 up to n = 128, 4e-9 at n = 256 with zeros out to 0.9)."""
    from oracle.fir_problems import fmp2_reference
    for n, rad, tol in MINPHASE_CASES:
        for k in range(4):
            h0, r = _minimum_phase_case(n, rad, seed=n + 100 * k)
            assert np.abs(fmp2_reference(r) - h0).max() < tol


def test_mag2mp_of_fmp2_is_the_function_pinned_through_b2a():
    """fir_ap_cvx.m:292-303 (local mag2mp) and rf_tools/mag2mp.m are the same text; the restatement used by b2a_m is pinned to
    the reference's compiled b2rf (tests/test_oracle.py), and the one used by fmp2_reference is bit-identical to it."""
    from oracle import ref
    from oracle.fir_problems import mag2mp_reference
    x = np.abs(np.random.default_rng(1).standard_normal(4096)) + 0.05
    assert np.array_equal(mag2mp_reference(x), ref.mag2mp_m(x))


def test_peak_lower_bound_is_valid_for_any_multipliers():
    """oracle.fir_problems.peak_lower_bound_fir_qp is a weak-duality bound: whatever the multipliers, it cannot exceed the Peak
    of a feasible point.  Feasibility by construction: the radii are set so that an arbitrary x0 satisfies every disk."""
    from oracle.fir_problems import H1_DUALBAND as S, build_fir_qp, peak_lower_bound_fir_qp, response_fir_qp, violation_fir_qp
    rng = np.random.default_rng(3)
    n = 24
    p = build_fir_qp(n, np.array(S["f"]) * 8, S["a"], S["d"], 20.0, 1.0)
    x0 = rng.standard_normal(2 * n) * 0.05
    p = dict(p, radius=np.abs(response_fir_qp(p["w"], n, x0) - p["center"]) + rng.uniform(0.0, 0.05, p["w"].size))
    assert violation_fir_qp(p, x0) == 0.0
    peak0 = np.hypot(x0[:n], x0[n:]).max()
    best = -np.inf
    for trial in range(200):
        y = rng.standard_normal(2 * p["w"].size) * rng.uniform(0, 1, 2 * p["w"].size) ** 8
        lb = peak_lower_bound_fir_qp(p, y)
        assert lb <= peak0 + 1e-12
        best = max(best, lb)
    # multipliers pointing from the disk centres towards the response of x0 on the tightest disks give a positive bound
    H = response_fir_qp(p["w"], n, x0)
    tight = np.argsort(p["radius"] - np.abs(H - p["center"]))[:5]
    y = np.zeros((p["w"].size, 2))
    d = (p["center"] - H)[tight]
    y[tight, 0], y[tight, 1] = d.real, d.imag
    assert peak_lower_bound_fir_qp(p, y.ravel()) <= peak0 + 1e-12
