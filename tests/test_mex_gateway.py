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
    return {n: os.path.join(MEX, n + "_stub.so") for n in ("blochC", "blochH", "abrx")}


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
