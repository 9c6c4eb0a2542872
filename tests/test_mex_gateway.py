"""The MEX gateways (multiband_rf_pulse_design_b200/matlab/*.c), compiled against the stub mex.h and
driven with stub mxArrays by the SAME harness that drives the reference's own mexFunction
(oracle/ref.py:mex_call).  MATLAB/Octave are absent here, so this is how the boundary is exercised."""
import os
import subprocess

import numpy as np
import pytest

from conftest import TOL_BLOCH, TOL_SLR, golden, golden_files

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
MEX = os.path.join(ROOT, "multiband_rf_pulse_design_b200", "matlab")


@pytest.fixture(scope="module")
def gateways(mbrf):
    subprocess.check_call(["make", "-s", "-C", MEX, "check"])
    return {n: os.path.join(MEX, n + "_stub.so") for n in ("blochC", "blochH", "abrx", "fir_pdhg", "fir_solve", "b2a", "ab2rf", "fmp2", "flip_zero", "fir_ap", "bloch_sweep")}


def test_gateways_link_and_report_errors_like_the_reference(gateways, oracle, mbrf):
    # abrx usage / length errors: same text as abrx.c:41,45
    out, err = oracle.mex_call(gateways["abrx"], 2, np.ones(8), np.ones(7), np.zeros(3))
    assert out is None and err == "rf and gradient vectors are of different lengths"
    out, err = oracle.mex_call(gateways["abrx"], 1, np.ones(8), np.ones(8), np.zeros(3))
    assert out is None and err == "Usage: [alpha, beta] = abrx(rf, g, x {, y})"
    if mbrf.lib().mbrf_device_count() == 0:
        # no device: the gateway must fail loudly through mexErrMsgTxt, never compute on the CPU
        out, err = oracle.mex_call(gateways["blochC"], 3, np.ones(4), np.zeros(4), 1e-5, 1.0, 1.0, np.zeros(3), 0.0)
        assert out is None and "no CPU path" in err
        out, err = oracle.mex_call(gateways["abrx"], 2, np.ones(8) * 0.1, np.ones(8), np.zeros(3))
        assert out is None and "no CPU path" in err


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_files("bloch_rand_*.npz")[::3] + golden_files("bloch_time_*.npz"))
def test_bloch_gateway_vs_reference_outputs(gateways, oracle, name):
    g = golden(name)
    if "tp" in g:
        args, nuc = [g["b1"], g["gr"], g["tp"], float(g["t1"]), float(g["t2"]), g["df"], g["dp"], 0], "C-13"
    else:
        args = [g["b1"].reshape(-1, 1), g["gr"], float(g["dt"]), float(g["t1"]), float(g["t2"]), g["df"], g["dp"],
                int(g["mode"])]
        if int(g["use_m0"]):
            args += [g["m0"][0], g["m0"][1], g["m0"][2]]
        nuc = str(g["nucleus"])
    outs, err = oracle.mex_call(gateways["blochC" if nuc == "C-13" else "blochH"], 3, *args)
    assert err is None, err
    for o, k in zip(outs, ("mx", "my", "mz")):
        assert o.shape == g[k].shape
        assert np.abs(o - g[k]).max() < TOL_BLOCH


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_files("abrx_*.npz"))
def test_abrx_gateway_vs_reference_outputs(gateways, oracle, name):
    g = golden(name)
    args = [g["rf"], g["g"], g["x"]]
    if "use_y" in g and int(g["use_y"]):
        args.append(g["y"])
    outs, err = oracle.mex_call(gateways["abrx"], 2, *args)
    assert err is None, err
    assert outs[0].shape == g["alpha"].shape
    assert np.abs(outs[0] - g["alpha"]).max() < TOL_SLR and np.abs(outs[1] - g["beta"]).max() < TOL_SLR


@pytest.mark.gpu
def test_fir_pdhg_gateway_solves_like_the_python_mirror(gateways, oracle, mbrf):
    """fir_pdhg_mex driven with the arguments matlab/fir_ap_cvx.m builds (MATLAB dim-by-B matrices, 1-based
    indices): same optimum as the Python mirror, which goes through the same C entry point."""
    import json
    from multiband_rf_pulse_design_b200 import fir
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "fir_ap_known.json")))["lowpass_n24"]
    n, obj, peak = k["n"], k["obj"], k["peak"]
    p = fir.assemble_fir_ap(n, k["f"], k["a"], k["d"], obj, peak)
    m, st = p["w"].size, np.nonzero(p["stop"])[0]
    N = 2 * n - 1
    w_row = np.concatenate([p["w"], p["w"][st]])
    col_type = np.concatenate([[0], np.ones(n - 1), 2 * np.ones(n - 1)])
    col_kappa = np.concatenate([[0], np.arange(1, n), np.arange(1, n)])
    col_amp = np.concatenate([[1], 2 * np.ones(2 * n - 2)])
    c = np.zeros((N, 1)); c[0] = 1
    lo = np.concatenate([p["lo"], np.full(st.size, -np.inf)]).reshape(-1, 1)
    hi = np.concatenate([p["hi"], np.zeros(st.size)]).reshape(-1, 1)
    bl = np.full((N, 1), -np.inf); bu = np.full((N, 1), np.inf)
    tmax = p["hi"][st].max()
    bl[0], bu[0] = -n * peak, n * peak
    rho = p["radius"][1:].reshape(-1, 1)
    outs, err = oracle.mex_call(gateways["fir_pdhg"], 2, w_row, np.zeros((0, 0)), col_type, col_kappa, col_amp, 0.0,
                                np.arange(2, n + 1, dtype=float), np.arange(n + 1, 2 * n, dtype=float), c, lo, hi, bl, bu,
                                rho, np.array([n * peak + obj * tmax]), np.array([200000, 64, 8e-7, 1e-4, 5e-5]),
                                np.array([m + 1, st.size, obj], dtype=float))
    assert err is None, err
    z, info = outs
    assert z.shape == (N, 1) and info.shape == (8, 1) and info[0, 0] == 1.0
    assert k["outer_obj"] * (1 - 1e-4) <= info[2, 0] <= k["inner_obj"] * (1 + 1e-4)   # x1 + obj*ripple_stop
    out, err = oracle.mex_call(gateways["fir_pdhg"], 2, w_row)
    assert out is None and err.startswith("Usage:")


E = np.zeros((0, 0))        # MATLAB []


@pytest.mark.gpu
def test_fir_solve_gateway_interior_point_like_fir_ap_cvx_batch_m(gateways, oracle, mbrf):
    """fir_solve_mex (method 1 = interior point) with the arguments matlab/fir_ap_cvx_batch.m builds for ONE design
    (dim-by-B matrices, 1-based rows, the 9-entry blocks vector): the optimum of the HiGHS-bracketed golden."""
    import json
    from multiband_rf_pulse_design_b200 import fir
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "fir_ap_known.json")))["lowpass_n24_obj10"]
    n, obj, peak = k["n"], k["obj"], k["peak"]
    p = fir.assemble_fir_ap(n, k["f"], k["a"], k["d"], obj, peak)
    allw = np.unique(p["w"])                                        # fir_ap_cvx_batch.m: union grid, ismember positions
    pos = np.searchsorted(allw, p["w"])
    M1 = allw.size
    srows = np.unique(pos[p["stop"]])
    M, N = M1 + srows.size, 2 * n - 1
    lo = np.full((M, 1), -np.inf); hi = np.full((M, 1), np.inf)
    np.maximum.at(lo[:, 0], pos, p["lo"]); np.minimum.at(hi[:, 0], pos, p["hi"])
    hi[M1:, 0] = 0.0
    w_row = np.concatenate([allw, allw[srows]])
    col_type = np.concatenate([[0], np.ones(n - 1), 2 * np.ones(n - 1)])
    col_kappa = np.concatenate([[0], np.arange(1, n), np.arange(1, n)])
    col_amp = np.concatenate([[1], 2 * np.ones(2 * n - 2)])
    c = np.zeros((N, 1)); c[0] = 1
    bl = np.full((N, 1), -np.inf); bu = np.full((N, 1), np.inf)
    bl[0], bu[0] = -n * peak, n * peak
    rho = ((n - np.arange(2, n + 1) + 1) * peak).reshape(-1, 1)
    blocks = np.array([M1 + 1, srows.size, 0, 0, 0, 0, 0, 0, 0], float)
    block_w = np.array([[obj], [0], [0], [0]], float)
    args = [1.0, w_row, E, E, col_type, col_kappa, col_amp, E, np.arange(2, n + 1, dtype=float), np.arange(n + 1, 2 * n, dtype=float),
            c, lo, hi, bl, bu, rho, E, np.array([100, 1e-7, 2e-6, 1e-12]), blocks, block_w]
    outs, err = oracle.mex_call(gateways["fir_solve"], 2, *args)
    assert err is None, err
    z, info = outs
    assert z.shape == (N, 1) and info.shape == (8, 1) and info[0, 0] == 1.0
    assert k["outer_obj"] * (1 - 1e-4) <= info[2, 0] <= k["inner_obj"] * (1 + 1e-4)
    # the first-order solver through the same gateway and the same arguments
    args[0] = 0.0
    args[16] = np.array([n * peak + obj * p["hi"][p["stop"]].max()])
    args[17] = np.array([200000, 64, 8e-7, 1e-4, 5e-5])
    (z2, info2), err = oracle.mex_call(gateways["fir_solve"], 2, *args)
    assert err is None and info2[0, 0] == 1.0 and abs(info2[2, 0] - info[2, 0]) <= 2e-4 * info[2, 0]
    # misuse is reported through mexErrMsgTxt
    bad = list(args); bad[0] = 1.0; bad[2] = np.zeros(M)           # a row phase with the interior-point solver
    out, err = oracle.mex_call(gateways["fir_solve"], 2, *bad)
    assert out is None and "interior-point solver takes" in err
    out, err = oracle.mex_call(gateways["fir_solve"], 2, *args[:5])
    assert out is None and err.startswith("Usage:")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["qp_n16_obj1", "mm_n16_a"])
def test_fir_solve_gateway_like_fir_qp_cvx_m(gateways, oracle, case):
    """fir_solve_mex (method 0) with the arguments matlab/fir_qp_cvx.m builds -- row phases / scales, explicit identity
    entries, the blocks vector, 4-by-B weights, objective scaling -- for the scalar-obj form and the minimax form:
    the SciPy trust-constr known answers of tests/golden (the same the Python mirror is held to)."""
    import json
    minimax = case.startswith("mm")
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "fir_qp_minimax_known.json" if minimax else "fir_qp_known.json")))[case]
    n, kk, obj = k["n"], k["k"], np.atleast_1d(np.asarray(k["obj"], float))
    f = np.asarray(k["f"], float) * np.pi; a = np.asarray(k["a"], float); d = np.asarray(k["d"], float)
    w = np.sort(np.concatenate([np.linspace(-np.pi, np.pi, n * 10), f]))                    # fir_qp_cvx.m:35-38
    inband = np.zeros(w.size, bool); Mb, Db, bidx = [], [], []
    for b in range(len(f) // 2):
        e0, e1 = f[2 * b], f[2 * b + 1]
        sel = np.nonzero((w >= e0) & (w <= e1))[0]
        amp = np.full(sel.size, a[2 * b]) if e0 == e1 else a[2 * b] + (a[2 * b + 1] - a[2 * b]) * (w[sel] - e0) / (e1 - e0)
        bidx.append(sel); Mb.append(amp); Db.append(np.full(sel.size, d[b])); inband[sel] = True
    bidx, Mb, Db = map(np.concatenate, (bidx, Mb, Db))
    wband, wtran = w[bidx], w[~inband]
    Hd = Mb * np.exp(1j * (kk * wband ** 2 - wband * (n - 1) / 2))
    wall = np.concatenate([wband, wtran]); m, nb = wall.size, wband.size
    centre = np.concatenate([Hd, np.zeros(wtran.size)]); radius = np.concatenate([Db, np.full(wtran.size, 1 + 5 * d.max())])
    N, M = 2 * n, 2 * m + 2 * n
    w_row = np.concatenate([np.repeat(wall, 2), np.zeros(2 * n)])
    row_phase = np.concatenate([np.tile([0, np.pi / 2], m), np.zeros(2 * n)])
    row_scale = np.concatenate([np.ones(2 * m), np.zeros(2 * n)])
    col_type = np.concatenate([np.ones(n), 2 * np.ones(n)]); col_kappa = np.concatenate([np.arange(n), np.arange(n)]).astype(float)
    entries = np.column_stack([2 * m + np.arange(1, 2 * n + 1), np.column_stack([np.arange(1, n + 1), n + np.arange(1, n + 1)]).ravel(),
                               np.ones(2 * n)]).astype(float)
    lo = np.full((M, 1), -np.inf); hi = np.full((M, 1), np.inf)
    blocks = np.zeros(9)
    if minimax:
        row_scale[0:2 * nb:2] = 1 / Db; row_scale[1:2 * nb:2] = 1 / Db
        lo[0:2 * nb:2, 0] = Hd.real / Db; lo[1:2 * nb:2, 0] = Hd.imag / Db
        lo[2 * nb:2 * m, 0] = 0; hi[2 * nb:2 * m:2, 0] = 1.1
        blocks[[7, 8]] = [1, nb]; blocks[[2, 3]] = [2 * nb + 1, m - nb]
        gw, lam, g2 = obj[1], obj[0], 1.0
    else:
        lo[0:2 * m:2, 0] = centre.real; lo[1:2 * m:2, 0] = centre.imag; hi[0:2 * m:2, 0] = radius
        blocks[[2, 3]] = [1, m]
        gw, lam, g2 = obj[0], 1.0, 0.0
    big = 2 * radius.max() + 2 * np.abs(centre).max() + 2 * minimax
    blocks[[4, 5]] = [2 * m + 1, n]; blocks[6] = N
    oscale = max(1.0, gw, lam)
    block_w = np.array([[0.0], [gw], [lam], [g2]]) / oscale
    outs, err = oracle.mex_call(gateways["fir_solve"], 2, 0.0, w_row, row_phase, row_scale, col_type, col_kappa, np.ones(N), entries,
                                E, E, np.zeros((N, 1)), lo, hi, np.full((N, 1), -big), np.full((N, 1), big), E, E,
                                np.array([400000, 64, 8e-7, 1e-4, 5e-5]), blocks, block_w)
    assert err is None, err
    z, info = outs
    assert info[0, 0] == 1.0
    assert abs(info[2, 0] * oscale - k["objective"]) <= 1e-4 * k["objective"]


@pytest.mark.gpu
def test_fir_solve_gateway_like_fir_linprog_m(gateways, oracle):
    """fir_solve_mex (method 1) with the arguments matlab/fir_linprog.m builds (no bounds, no pairs, no blocks: all [])
    for a complex even-length linear-phase design: the HiGHS optimum."""
    import json
    from multiband_rf_pulse_design_b200 import fir
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "fir_lp_known.json")))["lp_cplx_even_n40"]
    p = fir.assemble_fir_linprog(k["n"], k["f"], k["a"], k["d"])
    c = fir._lp_objective(p).reshape(-1, 1)
    outs, err = oracle.mex_call(gateways["fir_solve"], 2, 1.0, p["w"], E, E, p["col_type"].astype(float), p["col_kappa"], p["col_amp"], E,
                                E, E, c, p["lo"].reshape(-1, 1), p["hi"].reshape(-1, 1), E, E, E, E, np.array([100, 1e-7, 2e-6, 1e-12]), E, E)
    assert err is None, err
    z, info = outs
    assert info[0, 0] == 1.0 and abs(info[2, 0] - k["obj"]) <= 1e-4 * abs(k["obj"])


def test_postprocessing_gateways_usage_errors(gateways, oracle):
    """b2a / ab2rf / fmp2 gateways (matlab/islr_mex.c): arity and shape errors through mexErrMsgTxt, before any device work."""
    out, err = oracle.mex_call(gateways["b2a"], 1, np.ones(8), np.ones(8))
    assert out is None and err == "Usage: aca = b2a(bc)"
    out, err = oracle.mex_call(gateways["ab2rf"], 1, np.ones(8), np.ones(9))
    assert out is None and err == "ab2rf: ac and bc must have the same size"
    out, err = oracle.mex_call(gateways["fmp2"], 1, np.ones(8))
    assert out is None and err == "filter length must be odd"            # fir_ap_cvx.m:265-268


@pytest.mark.gpu
def test_postprocessing_gateways_vs_restatements(gateways, oracle):
    """One polynomial (vector) and a batch (n-by-B matrix) through the b2a / ab2rf / fmp2 gateways vs oracle/ restatements."""
    from oracle.fir_problems import fmp2_reference
    n, B = 64, 3
    k = np.arange(n) - (n - 1) / 2
    bc = np.stack([np.sinc(k * (3.0 + q) / n) * np.hamming(n) * 0.3 / (n / 8) * np.exp(1j * 0.01 * q * k) for q in range(B)], axis=1)
    (a1,), err = oracle.mex_call(gateways["b2a"], 1, bc[:, 0])
    assert err is None, err
    assert np.abs(a1.ravel() - oracle.b2a_m(bc[:, 0])).max() < 1e-11
    (ab,), err = oracle.mex_call(gateways["b2a"], 1, bc)
    assert err is None and ab.shape == (n, B)
    (rf,), err = oracle.mex_call(gateways["ab2rf"], 1, ab, bc)
    assert err is None and rf.shape == (n, B)
    for q in range(B):
        a_ref = oracle.b2a_m(bc[:, q])
        assert np.abs(ab[:, q] - a_ref).max() < 1e-11
        assert np.abs(rf[:, q] - oracle.ab2rf_m(a_ref, bc[:, q])).max() < 1e-9
    h = np.random.default_rng(0).standard_normal(33) * np.exp(-np.arange(33) / 9.0)
    r = np.correlate(h, h, mode="full"); r[32] *= 1.000001
    (hm,), err = oracle.mex_call(gateways["fmp2"], 1, r)
    assert err is None and hm.size == 33
    assert np.abs(hm.ravel() - fmp2_reference(r)).max() < 1e-10


def test_flip_zero_gateway_usage_errors(gateways, oracle):
    out, err = oracle.mex_call(gateways["flip_zero"], 1, np.ones(4), np.ones(2))
    assert out is None and err.startswith("Usage: [h_new, best, peak, power] = flip_zero_mex")
    out, err = oracle.mex_call(gateways["flip_zero"], 1, np.ones(4), np.array([1.0, 2.0]), np.zeros((3, 4)), 1.0)
    assert out is None and err == "flip_zero_mex: mask must have one row per passband zero"


@pytest.mark.gpu
def test_flip_zero_gateway_like_fir_flip_zero_m(gateways, oracle):
    """flip_zero_mex with the arguments matlab/fir_flip_zero.m builds (complex Z, 1-based idx_pb, N_z-by-Num mask) against the
    restatement of fir_flip_zero.m:66-99."""
    from oracle import fir_post as O
    from test_fir_post import _filter, flip_zero_check
    h = _filter(18, 7, seed=9)
    o = O.flip_zero_reference(h)
    outs, err = oracle.mex_call(gateways["flip_zero"], 4, o["Z"], (o["idx_pb"] + 1).astype(float), o["mask"].astype(float), h.sum())
    assert err is None, err
    h_new, best, peak, power = outs
    assert h_new.shape == (h.size, 1) and peak.shape == (1, 128)
    flip_zero_check(h, o["Z"], o["idx_pb"], o["mask"].T, dict(h_new=h_new.ravel(), best=int(best.ravel()[0]) - 1, peak=peak.ravel(),
                                                             power=power.ravel()))


@pytest.mark.gpu
def test_fir_solve_gateway_like_fir_qprog_phs_m(gateways, oracle):
    """fir_solve_mex (method 0) with the arguments matlab/fir_qprog_phs.m builds (a phase per row, no scales / entries / pairs,
    the norm term over all of x, obj_upper = amax): the minimiser of the CPU QP (oracle/fir_post.py, SciPy SLSQP)."""
    from multiband_rf_pulse_design_b200 import fir_post as P
    from oracle import fir_post as O
    from test_fir_post import SPEC
    n = 17
    p = P.assemble_fir_qprog_phs(n, SPEC["f"], SPEC["a"], SPEC["d"])
    N = 2 * n
    fin = np.concatenate([p["hi"][np.isfinite(p["hi"])], p["lo"][np.isfinite(p["lo"])]])
    big = 2 * np.sqrt(n) * max(1.0, np.abs(fin).max())
    blocks = np.zeros(9); blocks[6] = N
    outs, err = oracle.mex_call(gateways["fir_solve"], 2, 0.0, p["w"], p["phase"], E, np.concatenate([np.ones(n), 2 * np.ones(n)]),
                                np.concatenate([p["q"], p["q"]]), np.ones(N), E, E, E, np.zeros((N, 1)), p["lo"].reshape(-1, 1),
                                p["hi"].reshape(-1, 1), np.full((N, 1), -big), np.full((N, 1), big), E, np.array([p["amax"] * (1 + 1e-9)]),
                                np.array([400000, 64, 8e-7, 1e-5, 2e-5]), blocks, np.array([[0.0], [0.0], [1.0], [0.0]]))
    assert err is None, err
    z, info = outs
    assert info[0, 0] == 1.0
    o = O.build_fir_qprog_phs(n, SPEC["f"], SPEC["a"], SPEC["d"])
    r = O.solve_fir_qprog_phs_reference(o)
    assert r.success and (o["A"] @ z[:, 0] - o["B"]).max() < 1e-6
    assert abs(np.linalg.norm(z[:, 0]) - np.linalg.norm(r.x)) < 1e-4 * np.linalg.norm(r.x)


def test_fir_ap_gateway_usage_errors(gateways, oracle):
    out, err = oracle.mex_call(gateways["fir_ap"], 1, 24.0, np.zeros((6, 2)))
    assert out is None and err.startswith("Usage: [H, info, X, rows] = fir_ap_mex")
    out, err = oracle.mex_call(gateways["fir_ap"], 1, 24.0, np.zeros((5, 2)), np.zeros((5, 1)), np.zeros((2, 1)), 1.0, 1e-3)
    assert out is None and err == "fir_ap_mex: F must be 2*nband-by-B"
    out, err = oracle.mex_call(gateways["fir_ap"], 1, 24.0, np.zeros((6, 2)), np.zeros((4, 1)), np.zeros((3, 1)), 1.0, 1e-3)
    assert out is None and err == "fir_ap_mex: A must be 2*nband-by-B or one column"


@pytest.mark.gpu
def test_fir_ap_gateway_like_fir_ap_cvx_batch_m(gateways, oracle):
    """fir_ap_mex with the arguments matlab/fir_ap_cvx_batch.m builds (F one design per column, shared a / d columns, obj and
    Peak vectors): the known answer of tests/golden/fir_ap_known.json and an infeasible neighbour in the same batch."""
    import json
    from multiband_rf_pulse_design_b200 import fir
    k = json.load(open(os.path.join(ROOT, "tests", "golden", "fir_ap_known.json")))["lowpass_n24"]
    n = k["n"]
    f = np.asarray(k["f"], float)
    F = np.column_stack([f, f])
    outs, err = oracle.mex_call(gateways["fir_ap"], 4, float(n), F, np.asarray(k["a"], float).reshape(-1, 1),
                                np.asarray(k["d"], float).reshape(-1, 1), np.array([k["obj"], k["obj"]]), np.array([k["peak"], 1e-7]))
    assert err is None, err
    H, info, X, rows = outs
    assert H.shape == (n, 2) and info.shape == (8, 2) and X.shape == (2 * n - 1, 2)
    assert info[0, 0] == 1.0 and info[0, 1] == 2.0                         # Peak = 1e-7 cannot carry the pass band: certificate
    hs, st, ex = fir.fir_ap_cvx_batch(n, [f], k["a"], k["d"], [k["obj"]], [k["peak"]], return_info=True)
    assert abs(info[2, 0] - ex["info"][0, 2]) <= 1e-7 * abs(ex["info"][0, 2])
    assert np.abs(H[:, 0] - hs[0]).max() < 1e-6


def test_bloch_sweep_gateway_usage_errors(gateways, oracle, mbrf):
    out, err = oracle.mex_call(gateways["bloch_sweep"], 3, np.ones(4), 1e-5)
    assert out is None and err.startswith("Usage: [mx, my, mz] = bloch_sweep_mex")
    out, err = oracle.mex_call(gateways["bloch_sweep"], 3, np.ones(4), 1e-5, 1.0, 1.0, np.zeros((0, 0)), np.ones(2), 6726.1)
    assert out is None and err == "bloch_sweep_mex: b1, df and scale must be non-empty"
    if mbrf.lib().mbrf_device_count() == 0:
        out, err = oracle.mex_call(gateways["bloch_sweep"], 3, np.ones(4) * 0.01, 1e-5, 1.0, 1.0, np.zeros(3), np.ones(2), 6726.1)
        assert out is None and "no CPU path" in err


@pytest.mark.gpu
def test_bloch_sweep_gateway_vs_one_call_per_scale(gateways, oracle, mbrf):
    """bloch_sweep_mex (sim_rf_scale.m:82-89 as one call): column k equals the blochC gateway called with b1 * scale(k)."""
    g = golden("pulses.npz")
    b1 = g["b1_cfg1_gauss"]
    dt = 4e-3 / b1.size
    df = np.linspace(-2500.0, 2500.0, 101)
    scale = np.array([0.8, 1.0, 1.2])
    outs, err = oracle.mex_call(gateways["bloch_sweep"], 3, b1, dt, 1e3, 1e3, df, scale, 6726.1)
    assert err is None, err
    assert outs[0].shape == (df.size, scale.size)
    for k, sc in enumerate(scale):
        ref, err = oracle.mex_call(gateways["blochC"], 3, (b1 * sc).reshape(-1, 1), np.zeros((b1.size, 1)), dt, 1e3, 1e3, df, 0.0, 0)
        assert err is None, err
        for o, r in zip(outs, ref):
            assert np.abs(o[:, k] - r.ravel()).max() < 1e-12
