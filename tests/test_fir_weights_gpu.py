"""Parity of the convex FIR design step at the stop-band weights the reference's own callers use (obj = 1e4: dzrf_mb.m:167-170;
1e5: fir_qp.m:47) and in WIDE batches (>= 128 designs at N = 256, ragged band edges: the union-grid rows and the batched
kernels -- tcgen05 tiles for the first-order solver, moment contraction + batched Cholesky for the interior-point solver).

Known answers: HiGHS on the restated LP (tests/golden/make_golden_fir_weights.py -> fir_ap_weights_known.json; 1-13 minutes
per case on one core), peak cones inactive (Peak = 1).  Tolerances of the north star: objective 1e-4 relative, violation 1e-6
recomputed on the CPU from the returned point."""
import json
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
W = json.load(open(os.path.join(GOLDEN, "fir_ap_weights_known.json")))
BASE = json.load(open(os.path.join(GOLDEN, "fir_ap_known.json")))["h1_dualband_n256"]
TOL_OBJ, TOL_VIOL = 1e-4, 1e-6
FEASIBLE = [k for k, v in W.items() if v["status"] == 0]
INFEASIBLE = [k for k, v in W.items() if v["status"] != 0]


def _check(k, x, t):
    from oracle.fir_problems import build_fir_ap, violation_fir_ap
    p = build_fir_ap(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"])
    z = np.concatenate([x, [t]])
    obj = p["c"] @ z
    assert abs(obj - k["cone_free_obj"]) <= TOL_OBJ * k["cone_free_obj"], (obj, k["cone_free_obj"])
    assert violation_fir_ap(p, z) <= TOL_VIOL
    return abs(obj - k["cone_free_obj"]) / k["cone_free_obj"]


@pytest.mark.parametrize("case", FEASIBLE)
def test_reference_weights_single_design(mbrf, case):
    """one design at a time (thin path: matrix-vector products), interior-point solver: obj from 0.01 to 1e5"""
    from multiband_rf_pulse_design_b200 import fir
    k = W[case]
    hs, st, ex = fir.fir_ap_cvx_batch(k["n"], [k["f"]], k["a"], k["d"], [k["obj"]], [k["peak"]], return_info=True, method="ipm")
    assert st == ["Solved"] and hs[0].size == k["n"]
    rel = _check(k, ex["x"][0], ex["ripple_stop"][0])
    assert rel <= 1e-5            # an interior-point method does much better than the north star asks
    assert ex["info"][0, 1] <= 100


def _wide_batch(cases, filler_objs, filler_peaks, filler_fadds):
    f0 = np.array(BASE["f"])
    fl = [W[c]["f"] for c in cases] + [BASE["f"]]
    ol = [W[c]["obj"] for c in cases] + [BASE["obj"]]
    pl = [W[c]["peak"] for c in cases] + [1.0]
    for o in filler_objs:
        for pk in filler_peaks:
            for fa in filler_fadds:
                f = f0.copy()
                f[0::2] -= fa
                f[1::2] += fa
                fl.append(list(f)); ol.append(float(o)); pl.append(float(pk))
    return fl, ol, pl


def test_wide_batch_interior_point_vs_highs(mbrf):
    """>= 128 designs in ONE batch (mixed obj / Peak, non-zero f_add -> ragged union grid); all 11 feasible goldens, the
    base design and the 3 infeasible goldens are checked."""
    from multiband_rf_pulse_design_b200 import fir
    cases = FEASIBLE + INFEASIBLE
    fl, ol, pl = _wide_batch(cases, np.logspace(-2, 4, 7), [1e-2, 2e-3, 1.0, 5e-4], [0.0, 1e-4, 3e-4, 6e-4, 2e-4])
    assert len(fl) >= 128
    hs, st, ex = fir.fir_ap_cvx_batch(256, fl, BASE["a"], BASE["d"], ol, pl, return_info=True, method="ipm")
    for i, c in enumerate(cases):
        if W[c]["status"] == 0:
            assert st[i] == "Solved", c
            _check(W[c], ex["x"][i], ex["ripple_stop"][i])
        else:
            assert st[i] == "Failed" and int(ex["info"][i, 0]) == 2, c        # HiGHS status 4: a Farkas certificate here
    i = len(cases)
    assert st[i] == "Solved"
    _check(dict(BASE, peak=1.0), ex["x"][i], ex["ripple_stop"][i])
    codes = ex["info"][:, 0]
    assert set(np.unique(codes)) <= {1.0, 2.0, 3.0}
    assert (codes == 3).sum() <= 2           # designs at the feasibility boundary may stay undecided; none of the checked ones


def test_wide_batch_first_order_vs_highs(mbrf):
    """The tcgen05 path of the first-order solver (batch width >= 64) against the oracle, not against itself: 128 designs in
    one batch, the goldens with weights <= 1 checked to the north star's tolerances."""
    from multiband_rf_pulse_design_b200 import fir
    cases = [c for c in FEASIBLE if W[c]["obj"] <= 1.0]
    assert len(cases) >= 3
    fl, ol, pl = _wide_batch(cases, np.logspace(-2, 0, 5), [1e-2, 3e-3, 1.0, 2e-3, 1.5e-3], [0.0, 1e-4, 2e-4, 3e-4, 5e-4])
    assert len(fl) >= 128
    hs, st, ex = fir.fir_ap_cvx_batch(256, fl, BASE["a"], BASE["d"], ol, pl, return_info=True, method="pdhg", max_iter=120000)
    for i, c in enumerate(cases):
        assert st[i] == "Solved", c
        _check(W[c], ex["x"][i], ex["ripple_stop"][i])
    i = len(cases)
    assert st[i] == "Solved"
    _check(dict(BASE, peak=1.0), ex["x"][i], ex["ripple_stop"][i])


def test_both_solvers_agree_on_active_cones(mbrf):
    """Peak cones in the problem at the reference's default Peak = 1e-3 scale (no LP golden exists with cones; the n = 24
    'tightpeak' case of test_fir_gpu.py brackets binding cones with HiGHS): the two solvers -- different algorithms,
    different code -- agree at N = 256."""
    from multiband_rf_pulse_design_b200 import fir
    args = (256, [BASE["f"]] * 2, BASE["a"], BASE["d"], [0.1, 1.0], [1.1e-3, 2e-3])
    _, s1, e1 = fir.fir_ap_cvx_batch(*args, return_info=True, method="ipm")
    _, s2, e2 = fir.fir_ap_cvx_batch(*args, return_info=True, method="pdhg", max_iter=120000)
    assert s1 == s2 == ["Solved", "Solved"]
    for b in range(2):
        assert abs(e1["info"][b, 2] - e2["info"][b, 2]) <= TOL_OBJ * abs(e1["info"][b, 2])
        assert e1["info"][b, 4] <= TOL_VIOL and e2["info"][b, 4] <= TOL_VIOL
        assert e1["info"][b, 2] >= BASE["cone_free_obj"] * (1 - 1e-6)      # never below the cone-free optimum


def test_iteration_limit_is_not_infeasibility(mbrf, monkeypatch):
    """fir_ap.m:86,149 branch on 'Failed'.  A probe that merely ran out of iterations (status 3) must not steer the bisection:
    it is solved again with more iterations, and an UndecidedProbe warning is raised if that does not decide it either."""
    from multiband_rf_pulse_design_b200 import fir
    k = json.load(open(os.path.join(GOLDEN, "fir_ap_known.json")))["lowpass_minorder_from40"]
    monkeypatch.setattr(fir, "IPM_MAX_ITER", 12)          # every feasible probe now hits the limit at the first attempt
    hs, st, ex = fir.fir_ap_cvx_batch(24, [k["f"]], k["a"], k["d"], [0.1], [k["peak"]], return_info=True, method="ipm")
    assert st == ["Failed"] and int(ex["info"][0, 0]) == 3          # the string cannot tell; the code can
    with warnings.catch_warnings():
        warnings.simplefilter("error", fir.UndecidedProbe)
        h, st, n_op, _ = fir.fir_ap(k["n"], k["f"], k["a"], k["d"], k["peak"], 1, 0, 0, 0, method="ipm")
    assert st == "Solved" and n_op == k["n_op"]                     # the retries (36 iterations) decide every probe
    monkeypatch.setattr(fir, "IPM_MAX_ITER", 3)           # now even the retry (9 iterations) cannot decide
    with pytest.warns(fir.UndecidedProbe):
        hs, st, codes = fir.fir_ap_cvx_decided(24, [k["f"]], k["a"], k["d"], [0.1], [k["peak"]], method="ipm")
    assert st == ["Failed"] and codes == [3]


def test_c13_bssfp_order_search_finds_the_reference_order(mbrf):
    """BASELINE config 5 on a reduced search range: the spec restated from bSSFP_pulse_lp_ap.m / spectrum_C13.m /
    rf_ripple_GFA.m (bench.py c13_bssfp_spec).  The reference script hard-codes the outcome of this very search as
    dzrf_mb(..., 58) (bSSFP_pulse_lp_ap.m:79).  The interior-point solver here decides 58 and 60 feasible, 50 infeasible
    (Farkas certificate) and finds n = 57 feasible as well -- violation 1e-12 recomputed on the CPU, objective 0.0277 against
    0.0147 at n = 58 -- so the search may end one tap below the reference's CVX-driven answer; 52..56 sit at the edge of
    feasibility (no decision within 300 iterations: UndecidedProbe, treated as 'Failed')."""
    import bench
    from multiband_rf_pulse_design_b200 import fir
    from oracle.fir_problems import build_fir_ap, violation_fir_ap
    f, a, d, dt = bench.c13_bssfp_spec()
    assert abs(dt - 0.02) < 1e-15 and len(f) == 10 and abs(a[0] - np.sin(np.pi / 6)) < 1e-3 and abs(d[1] - 0.0025) < 1e-6
    for n, want in ((58, 1), (60, 1), (50, 2)):
        hs, st, ex = fir.fir_ap_cvx_batch(n, [f], a, d, [0.1], [1e-3], return_info=True, method="ipm")
        assert int(ex["info"][0, 0]) == want, (n, ex["info"][0])
        if want == 1:
            z = np.concatenate([ex["x"][0], [ex["ripple_stop"][0]]])
            assert violation_fir_ap(build_fir_ap(n, f, a, d, 0.1, 1e-3), z) <= TOL_VIOL
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", fir.UndecidedProbe)
        h, st, n_op, _ = mbrf.fir_ap(96, f, a, d, 1e-3, 1, 0, 0, 0, method="ipm")
    assert st == "Solved" and n_op in (57, 58) and h.size == n_op


def test_fir_qp_searches_walk_the_reference_bisections(mbrf):
    """fir_qp.m (every probe is fir_ap_cvx(n, f, a, d, 1e5), :47-130): the batched / speculative searches of the mirror make
    exactly the decisions of the reference's serial loops (re-run here one probe at a time), and the returned filter meets the
    specification of the band edges the search ended with."""
    from multiband_rf_pulse_design_b200 import fir
    n, f, a, d = 40, np.array([-0.12, 0.12, 0.4, 1.0]), [0.2, 0.2, 0, 0], [0.01, 0.005]
    lam, peak = 1e5, 1e-3

    def probe(nn, ff):
        hs, st, _ = fir.fir_ap_cvx_decided(int(nn), [np.asarray(ff, float)], a, d, [lam], [peak])
        return hs[0], st[0]

    # serial transition search, fir_qp.m:57-96
    centre, top, bot = (f[2] + f[1]) / 2, (f[2] - f[1]) / 2, 0.0
    edges = lambda dfv: np.array([-(centre - dfv), centre - dfv, centre + dfv, f[3]])   # noqa: E731
    assert probe(n, f)[1] == "Solved"
    while True:
        mid = (top + bot) / 2
        if probe(n, edges(mid))[1] == "Failed":
            bot = mid
        else:
            top = mid
        if top - bot < 0.001:
            break
    f_want = edges(((f[2] - f[1]) / 2) * (1 - 0.9) + top * 0.9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", fir.UndecidedProbe)
        h, st = mbrf.fir_qp(n, f, a, d, 0, 0.9, 0, 0)
    assert st == "Solved" and h.size == n
    w = np.linspace(-np.pi, np.pi, 4001)
    H = np.abs(np.exp(-1j * np.outer(w, np.arange(n))) @ h)
    pb = (w >= f_want[0] * np.pi) & (w <= f_want[1] * np.pi)
    sb = (w >= f_want[2] * np.pi) & (w <= f_want[3] * np.pi)
    assert H[pb].max() <= 0.2 + 0.01 + 2e-3 and H[pb].min() >= 0.2 - 0.01 - 2e-3 and H[sb].max() <= 0.005 + 2e-3
    # serial order search, fir_qp.m:103-123
    n_top, n_bot = n, 2
    while True:
        n_mid = int(np.ceil((n_top + n_bot) / 2))
        if probe(n_mid, f)[1] == "Failed":
            n_bot = n_mid
        else:
            n_top = n_mid
        if n_top - n_bot == 1:
            break
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", fir.UndecidedProbe)
        h2, st2 = mbrf.fir_qp(n, f, a, d, 1, 0, 0, 0)
    assert st2 == "Solved" and h2.size == n_top
