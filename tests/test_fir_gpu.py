"""GPU tests of the convex FIR design step, through the C ABI, on BOTH solvers: the interior-point method
(mbrf_fir_ipm_solve, the default of the mirrors) and the batched restarted PDHG (mbrf_fir_pdhg_solve).

Tolerances of BASELINE.json's north star: optimal objective within 1e-4 relative of the reference solve,
constraint violation <= 1e-6.  The reference solve is HiGHS on the restated problem (cone-free LP, and
outer/inner 32-gon brackets of the peak cones) — tests/golden/fir_ap_known.json."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
KNOWN = json.load(open(os.path.join(GOLDEN, "fir_ap_known.json")))
TOL_OBJ = 1e-4     # relative
TOL_VIOL = 1e-6    # absolute, in the problem's own units (|H|^2, |r_k|)


@pytest.fixture(params=["ipm", "pdhg"], autouse=True)
def solver_method(request):
    """every test of this file runs once per solver (fir_qp_cvx has the first-order solver only)"""
    from multiband_rf_pulse_design_b200 import fir
    old, fir.DEFAULT_METHOD = fir.DEFAULT_METHOD, request.param
    yield request.param
    fir.DEFAULT_METHOD = old


def _solve(mbrf, k, **kw):
    from multiband_rf_pulse_design_b200 import fir
    hs, st, ex = fir.fir_ap_cvx_batch(k["n"], [k["f"]], k["a"], k["d"], [k["obj"]], [k["peak"]], return_info=True, **kw)
    return hs[0], st[0], ex


@pytest.mark.parametrize("case", ["lowpass_n24", "lowpass_n24_obj10", "lowpass_n24_tightpeak", "twoband_n40"])
def test_small_designs_vs_highs(mbrf, case):
    from oracle.fir_problems import build_fir_ap, violation_fir_ap
    k = KNOWN[case]
    h, st, ex = _solve(mbrf, k)
    assert st == "Solved" and h.size == k["n"]
    p = build_fir_ap(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"])
    z = np.concatenate([ex["x"][0], [ex["ripple_stop"][0]]])
    obj = p["c"] @ z
    assert k["outer_obj"] * (1 - TOL_OBJ) <= obj <= k["inner_obj"] * (1 + TOL_OBJ), (obj, k["outer_obj"], k["inner_obj"])
    assert violation_fir_ap(p, z) <= TOL_VIOL


def test_n256_dualband_known_answer(mbrf):
    """BASELINE config 4's spec at N=256: 7686 grid rows, objective 0.0160330 (HiGHS, cones inactive)."""
    from oracle.fir_problems import build_fir_ap, violation_fir_ap
    k = KNOWN["h1_dualband_n256"]
    h, st, ex = _solve(mbrf, k)
    assert st == "Solved"
    p = build_fir_ap(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"])
    z = np.concatenate([ex["x"][0], [ex["ripple_stop"][0]]])
    assert abs(p["c"] @ z - k["cone_free_obj"]) <= TOL_OBJ * k["cone_free_obj"]
    assert violation_fir_ap(p, z) <= TOL_VIOL
    # the returned taps are the minimum-phase spectral factor: |H(w)|^2 reproduces the solved spectrum A x
    from oracle.fir_problems import matrix_fir_ap
    w = np.linspace(-np.pi, np.pi, 512, endpoint=False)
    S = matrix_fir_ap(w, k["n"]) @ ex["x"][0]
    H = np.array([np.sum(h * np.exp(-1j * wi * np.arange(k["n"]))) for wi in w])
    assert np.abs(np.abs(H) ** 2 - S).max() < 2e-2 * S.max()      # fmp2 is a 4096-point FFT/Hilbert approximation (~1 %)


@pytest.mark.parametrize("case", ["lowpass_n24_peak_infeasible", "lowpass_n10_infeasible", "h1_dualband_n128_infeasible"])
def test_infeasible_designs_fail_like_cvx(mbrf, case):
    """fir_ap_cvx.m:176-182: anything but 'Solved' is 'Failed' with h = []."""
    k = KNOWN[case]
    h, st = mbrf.fir_ap_cvx(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"], max_iter=40000)
    assert st == "Failed" and h.size == 0


def test_batch_equals_single_and_ragged_edges(mbrf):
    """Designs with different band edges share one matrix (union of grid rows); results equal single solves."""
    from multiband_rf_pulse_design_b200 import fir
    k = KNOWN["lowpass_n24"]
    f0 = np.array(k["f"], float)
    fl = []
    for fa in (0.0, 0.01, 0.03):
        f = f0.copy()
        f[0::2] -= fa
        f[1::2] += fa
        f = np.clip(f, -1, 1)
        fl.append(f)
    objs, peaks = [0.1, 1.0, 0.1], [0.02, 0.02, 0.03]
    hs, st, ex = fir.fir_ap_cvx_batch(k["n"], fl, k["a"], k["d"], objs, peaks, return_info=True)
    assert st == ["Solved"] * 3
    for b in range(3):
        h1, s1, e1 = fir.fir_ap_cvx_batch(k["n"], [fl[b]], k["a"], k["d"], [objs[b]], [peaks[b]], return_info=True)
        assert s1 == ["Solved"]
        o_b, o_1 = ex["info"][b, 2], e1["info"][0, 2]
        assert abs(o_b - o_1) <= 2 * TOL_OBJ * abs(o_1)


def test_fir_ap_min_transition_search(mbrf):
    """fir_ap.m:57-134: widen the bands until the design fails; the returned spec is feasible, a wider one is not."""
    k = KNOWN["lowpass_n24"]
    h, st, n_op, f_op = mbrf.fir_ap(k["n"], k["f"], k["a"], k["d"], 0.02, 0, 1.0, 0, 0, max_iter=60000)
    assert st == "Solved" and n_op == k["n"] and h.size == k["n"]
    widened = f_op[1] - np.array(k["f"])[1]
    assert widened > 0.0
    # judge the answer with the independent CPU solver: the returned spec is feasible, and one widened by a few
    # bisection thresholds (df_thre = 5e-4, fir_ap.m:45) is not
    from oracle.fir_problems import build_fir_ap, solve_fir_ap_highs
    r, _ = solve_fir_ap_highs(build_fir_ap(k["n"], f_op, k["a"], k["d"], 0.1, 0.02))
    assert r.status == 0
    f_more = np.array(f_op, float)
    f_more[0::2] -= 4e-3
    f_more[1::2] += 4e-3
    r, _ = solve_fir_ap_highs(build_fir_ap(k["n"], np.clip(f_more, -1, 1), k["a"], k["d"], 0.1, 0.02))
    assert r.status != 0          # HiGHS: 2 infeasible (4 = numerically at the boundary)

# ---- ss/fir_linprog.m family ---------------------------------------------------------------------
LPK = json.load(open(os.path.join(GOLDEN, "fir_lp_known.json")))


@pytest.mark.parametrize("case", ["lp_real_odd_n31", "lp_real_even_n30", "lp_cplx_odd_n41", "lp_cplx_even_n40"])
def test_fir_linprog_vs_highs(mbrf, case):
    from oracle.fir_problems import build_fir_lp
    k = LPK[case]
    h, st, ex = mbrf.fir_linprog(k["n"], k["f"], k["a"], k["d"], return_info=True)
    assert st == "Solved" and h.size == k["n"]
    p = build_fir_lp(k["n"], k["f"], k["a"], k["d"])
    x = ex["x"]
    H = p["A"] @ x
    assert max((H - p["hi"]).max(), (p["lo"] - H).max()) <= TOL_VIOL
    assert abs(p["c"] @ x - k["obj"]) <= TOL_OBJ * abs(k["obj"])
    # linear phase: Hermitian taps (fill_h, fir_linprog.m:274-296)
    assert np.abs(h - np.conj(h[::-1])).max() < 1e-12


def test_fir_linprog_h0_is_a_warm_start_for_the_first_order_solver(mbrf):
    """h0 (ss/fir_linprog.m:157, fill_opt_param :298-373): with method="pdhg" the previous filter is the starting iterate --
    same optimum, fewer iterations; with the interior-point solver it is accepted and unused."""
    k = LPK["lp_cplx_even_n40"]
    h, st, ex = mbrf.fir_linprog(k["n"], k["f"], k["a"], k["d"], return_info=True, method="pdhg")
    assert st == "Solved"
    h2, st2, ex2 = mbrf.fir_linprog(k["n"], k["f"], k["a"], k["d"], h, return_info=True, method="pdhg")
    assert st2 == "Solved" and abs(ex2["info"][2] - ex["info"][2]) <= 2 * TOL_OBJ * abs(ex["info"][2])
    assert ex2["info"][1] < ex["info"][1]
    h3, st3, ex3 = mbrf.fir_linprog(k["n"], k["f"], k["a"], k["d"], h, return_info=True, method="ipm")
    assert st3 == "Solved" and abs(ex3["info"][2] - k["obj"]) <= TOL_OBJ * abs(k["obj"])


def test_fir_min_order_a_min_bounds_the_transition_response(mbrf):
    """a_min of ss/fir_min_order.m:14 (ss/fir_pm.m:42-43,102): the response in the transition regions stays above it.  Default
    is min(0, min(a - d)) = -0.01 here; a_min = 0 is honoured by the returned filter and cannot make it shorter."""
    k = LPK["minorder_real_n40"]
    h0, st0 = mbrf.fir_min_order(k["n"], k["f"], k["a"], k["d"], 1, None, 0)
    h1, st1 = mbrf.fir_min_order(k["n"], k["f"], k["a"], k["d"], 1, 0.0, 0)
    assert st0 == st1 == "Solved" and h1.size >= h0.size == k["linprog_len"]
    wt = np.linspace(k["f"][1], k["f"][2], 400)[1:-1] * np.pi
    resp = lambda h: np.real(np.exp(-1j * np.outer(wt, np.arange(h.size) - (h.size - 1) / 2)) @ h)   # noqa: E731
    assert resp(h1).min() >= -1e-4
    h2, st2 = mbrf.fir_linprog(h0.size, k["f"], k["a"], k["d"], a_min=0.5)      # the transition cannot stay above 0.5 down to the stop band
    assert st2 == "Failed"


def test_fir_linprog_failures(mbrf):
    k = LPK["lp_real_odd_n11_infeasible"]
    h, st = mbrf.fir_linprog(k["n"], k["f"], k["a"], k["d"], max_iter=40000)
    assert st == "Failed" and h.size == 0
    # even length with amplitude 1 at fs/2 is refused up front (fir_linprog.m:63-75)
    h, st = mbrf.fir_linprog(20, [0, 0.3, 0.5, 1], [0, 0, 1, 1], [0.01, 0.01])
    assert st == "Failed" and h.size == 0


@pytest.mark.parametrize("case", ["minorder_real_n40", "minorder_cplx_n48"])
def test_min_order_searches(mbrf, case):
    """ss/fir_min_order_linprog.m and ss/fir_min_order.m (LP-feasibility form): same bisection, same answer as
    the search driven by HiGHS probes; fir_min_order keeps the reference's 'longer of odd/even' selection."""
    k = LPK[case]
    h, st = mbrf.fir_min_order_linprog(k["n"], k["f"], k["a"], k["d"], 0, 0, max_iter=60000)
    assert st == k["linprog_status"] and h.size == k["linprog_len"]
    h, st = mbrf.fir_min_order(k["n"], k["f"], k["a"], k["d"], 0, None, 0, max_iter=60000)
    assert st == k["minorder_status"] and h.size == k["minorder_len"]


# ---- fir_qp_cvx.m (scalar-obj SOCP) ------------------------------------------------------------------
def _qp_known():
    return json.load(open(os.path.join(GOLDEN, "fir_qp_known.json")))


@pytest.mark.parametrize("case", ["qp_n16_obj1", "qp_n16_obj0", "qp_n12_obj3"])
def test_fir_qp_cvx_vs_scipy_reference(mbrf, case):
    """Objective E_total + obj*Peak within 1e-4 relative of SciPy trust-constr, every disk satisfied to 1e-6."""
    from oracle.fir_problems import build_fir_qp, objective_fir_qp, violation_fir_qp
    k = _qp_known()[case]
    h, st, ex = mbrf.fir_qp_cvx(k["n"], k["f"], k["a"], k["d"], k["k"], k["obj"], return_info=True)
    assert st == "Solved" and h.size == k["n"]
    p = build_fir_qp(k["n"], k["f"], k["a"], k["d"], k["k"], k["obj"])
    x = ex["x"]
    assert np.array_equal(h, x[:k["n"]] + 1j * x[k["n"]:])                  # fir_qp_cvx.m:209
    assert abs(objective_fir_qp(p, x) - k["objective"]) <= TOL_OBJ * k["objective"]
    assert violation_fir_qp(p, x) <= TOL_VIOL


def test_fir_qp_cvx_reference_call_is_certified_on_the_cpu(mbrf, solver_method):
    """BASELINE config 3 on the reference grid: fir_qp_cvx(256, f, a, d, 120, 1e6) (dzrf_mb.m:211-213).  No CPU solver finishes
    this SOCP in test time, so the answer is bracketed instead of compared: the returned x gives an upper bound on the optimum
    (objective and every disk recomputed on the CPU), and the returned multipliers of the disk rows give a RIGOROUS lower bound
    by weak duality, also evaluated on the CPU (oracle.fir_problems.peak_lower_bound_fir_qp: valid for any multipliers, so
    nothing of the solver's own bookkeeping is trusted).  The optimum lies in between; the bracket is the test."""
    if solver_method != "pdhg":
        pytest.skip("fir_qp_cvx runs on the first-order solver only")
    from oracle.fir_problems import H1_DUALBAND as S, build_fir_qp, objective_fir_qp, peak_lower_bound_fir_qp, violation_fir_qp
    n, k, obj = 256, 120.0, 1e6
    h, st, ex = mbrf.fir_qp_cvx(n, S["f"], S["a"], S["d"], k, obj, return_info=True, max_iter=1500000, want_dual=True)
    assert st == "Solved"
    p = build_fir_qp(n, S["f"], S["a"], S["d"], k, obj)
    x = ex["x"]
    assert violation_fir_qp(p, x) <= TOL_VIOL
    upper = objective_fir_qp(p, x)
    lower = obj * peak_lower_bound_fir_qp(p, ex["y"][:2 * p["w"].size])
    print(f"cfg3 bracket: {lower:.6f} <= optimum <= {upper:.6f}  (relative width {(upper - lower) / upper:.2e})")
    # x satisfies the disks to 1e-6 absolute (7e-5 of the smallest radius), so its objective may undercut the optimum by about
    # that much: the two bounds may cross by the tolerance, not more
    assert lower <= upper * (1 + TOL_OBJ)
    assert abs(upper - lower) <= TOL_OBJ * upper                               # measured: 8.7e-6 (14864.5427 <= optimum <= 14864.6724)


def test_fir_qp_cvx_config3(mbrf):
    """BASELINE config 3: min-energy multiband FIR, N=256, dual-band H-1 spec, k=120 (dzrf_mb.m:211-213)."""
    from oracle.fir_problems import H1_DUALBAND, build_fir_qp, objective_fir_qp, violation_fir_qp
    h, st, ex = mbrf.fir_qp_cvx(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 120, 1.0, return_info=True)
    assert st == "Solved" and h.size == 256
    p = build_fir_qp(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 120, 1.0)
    assert p["w"].size == 2560 + 6                                           # m = 2560 + 2*nband
    assert violation_fir_qp(p, ex["x"]) <= TOL_VIOL
    # the solver's own objective and its dual value bracket the optimum to 1e-4
    assert abs(ex["info"][2] - objective_fir_qp(p, ex["x"])) <= 1e-9
    assert abs(ex["info"][2] - ex["info"][3]) <= TOL_OBJ * ex["info"][2]
    with pytest.raises(ValueError):
        mbrf.fir_qp_cvx(16, [-0.5, 0.5], [1, 1], [0.1], 2, [1.0, 1.0, 1.0])     # fir_qp_cvx.m:193-195 "invalid input of obj"


@pytest.mark.parametrize("case", ["mm_n16_a", "mm_n16_b", "mm_n12_c"])
def test_fir_qp_cvx_minimax_vs_scipy_reference(mbrf, case):
    """Two-element obj (fir_qp_cvx.m:170-191): delta + obj(1)*E_total + obj(2)*Peak within 1e-4 relative of SciPy
    trust-constr, transition rows within 1.1 to 1e-6."""
    from oracle.fir_problems import build_fir_qp, objective_fir_qp_minimax, violation_fir_qp_minimax
    k = json.load(open(os.path.join(GOLDEN, "fir_qp_minimax_known.json")))[case]
    h, st, ex = mbrf.fir_qp_cvx(k["n"], k["f"], k["a"], k["d"], k["k"], k["obj"], return_info=True)
    assert st == "Solved" and h.size == k["n"]
    p = build_fir_qp(k["n"], k["f"], k["a"], k["d"], k["k"], 0.0)
    x = ex["x"]
    assert abs(ex["info"][2] - objective_fir_qp_minimax(p, x, k["obj"])) <= 1e-9 * max(1.0, ex["info"][2])
    assert abs(objective_fir_qp_minimax(p, x, k["obj"]) - k["objective"]) <= TOL_OBJ * k["objective"]
    assert violation_fir_qp_minimax(p, x) <= TOL_VIOL


def test_fir_ap_min_order_search(mbrf):
    """fir_ap.m:137-176 (BASELINE config 5's search, here on a small spec): minimum order by bisection, probes of one
    round solved concurrently; same answer as the serial bisection driven by HiGHS (Peak loose: cones inactive)."""
    k = KNOWN["lowpass_minorder_from40"]
    h, st, n_op, f_op = mbrf.fir_ap(k["n"], k["f"], k["a"], k["d"], k["peak"], 1, 0, 0, 0, max_iter=60000)
    assert st == "Solved" and n_op == k["n_op"] and h.size == k["n_op"]
    assert np.array_equal(f_op, np.array(k["f"], float))
