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
    return {n: os.path.join(MEX, n + "_stub.so") for n in ("blochC", "blochH", "abrx", "fir_pdhg", "b2a", "ab2rf", "fmp2")}


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
