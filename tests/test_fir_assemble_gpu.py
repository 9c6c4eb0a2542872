"""Device-side problem assembly of fir_ap_cvx (C ABI `mbrf_fir_ap_assemble`, `mbrf_fir_ap_solve`):
bit-identical to the host (numpy) assembly of the Python mirror and to the oracle's restatement of fir_ap_cvx.m:44-142, and the
one-call path (specification in, taps out) against the host-assembled path and the committed HiGHS known answers."""
import json
import os

import numpy as np
import pytest

from oracle import fir_problems as O

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
S = O.H1_DUALBAND


def _mixed_batch():
    """Designs that exercise every branch of the assembly: widened band edges (ragged union grid), per-design amplitudes and
    ripples, a sloped band, touching bands (one grid value in two bands), an edge on a base sample (f = -1), degenerate band."""
    f0, a0, d0 = S["f"], S["a"], S["d"]
    df = (f0[2:-1:2] - f0[1:-2:2]).min()
    fl, al, dl, ol, pl = [], [], [], [], []
    for k, fa in enumerate(np.linspace(0, 0.4 * df / 2, 5)):
        fn = f0.copy(); fn[0::2] -= fa; fn[1::2] += fa
        fl.append(fn); al.append(a0); dl.append(d0 * (1 + 0.1 * k)); ol.append(10.0 ** (k - 2)); pl.append(1e-3 * (k + 1))
    fl.append(np.array([-1.0, -0.5, -0.2, 0.1, 0.3, 0.7])); al.append(np.array([0.0, 0.0, 1.0, 0.8, 0.0, 0.0]))   # sloped pass band
    dl.append(np.array([0.01, 0.05, 0.02])); ol.append(1.0); pl.append(2e-3)
    fl.append(np.array([-0.6, -0.2, -0.2, 0.2, 0.4, 0.4])); al.append(np.array([0.0, 0.0, 1.0, 1.0, 0.0, 0.0]))   # touching + point band
    dl.append(np.array([0.02, 0.03, 0.04])); ol.append(0.5); pl.append(5e-3)
    fl.append(f0.copy()); al.append(a0); dl.append(d0); ol.append(1e4); pl.append(1e-3)
    return fl, al, dl, ol, pl


@pytest.mark.gpu
@pytest.mark.parametrize("n", [24, 256])
def test_device_assembly_is_bit_identical_to_host_assembly(mbrf, n):
    from multiband_rf_pulse_design_b200 import fir
    fl, al, dl, ol, pl = _mixed_batch()
    dev = fir.assemble_fir_ap_device(n, fl, al, dl, ol, pl)
    host = fir._assemble_batch_ap(n, [fir.assemble_fir_ap(n, fl[i], al[i], dl[i], ol[i], pl[i]) for i in range(len(fl))])
    assert dev["M1"] == host["M1"] and dev["ns"] == host["srows"].size and dev["M"] == host["M"]
    for k in ("w_row", "lo", "hi", "c", "bl", "bu", "rho", "sw"):
        assert np.array_equal(dev[k], host[k]), k                         # bit for bit, infinities included


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(8))
def test_device_assembly_vs_oracle_restatement(mbrf, case):
    """One design at a time against oracle/fir_problems.build_fir_ap (fir_ap_cvx.m:44-142 restated independently of the
    mirror's batching): rows in the oracle's band-then-transition order are looked up by frequency in the device's sorted grid."""
    from multiband_rf_pulse_design_b200 import fir
    fl, al, dl, ol, pl = _mixed_batch()
    n = 40
    p = O.build_fir_ap(n, fl[case], al[case], dl[case], ol[case], pl[case])
    dev = fir.assemble_fir_ap_device(n, [fl[case]], [al[case]], [dl[case]], [ol[case]], [pl[case]])
    M1 = dev["M1"]
    w = dev["w_row"][:M1]
    ix = np.searchsorted(w, p["w"])
    assert np.array_equal(w[ix], p["w"])                                  # every oracle row is a grid row, bit for bit
    hi = np.full(M1, np.inf); lo = np.full(M1, -np.inf)
    np.minimum.at(hi, ix, p["hi"]); np.maximum.at(lo, ix, p["lo"])        # a value in two bands is one row: the intersection
    assert np.array_equal(dev["hi"][:M1, 0], hi) and np.array_equal(dev["lo"][:M1, 0], lo)
    stop_rows = np.unique(ix[p["stop"]])
    assert np.array_equal(dev["w_row"][M1:], w[stop_rows])
    assert np.all(dev["hi"][M1:, 0] == 0.0) and np.all(np.isneginf(dev["lo"][M1:, 0]))
    assert dev["sw"][0] == ol[case] and np.array_equal(dev["rho"][:, 0], p["radius"][1:])
    assert dev["bu"][0, 0] == p["radius"][0] == -dev["bl"][0, 0] and np.all(np.isposinf(dev["bu"][1:, 0]))


@pytest.mark.gpu
def test_one_call_path_vs_host_assembled_path(mbrf):
    """mbrf_fir_ap_solve (assembly, solve and fmp2 on the device) against host assembly + mbrf_fir_ipm_solve + the fmp2 call."""
    from multiband_rf_pulse_design_b200 import fir
    fl, al, dl, ol, pl = _mixed_batch()
    n = 256                                                                # the dual-band specification needs about 256 taps
    hd, sd, ed = fir.fir_ap_cvx_batch(n, fl, al, dl, ol, pl, return_info=True, method="ipm", assemble="device")
    hh, sh, eh = fir.fir_ap_cvx_batch(n, fl, al, dl, ol, pl, return_info=True, method="ipm", assemble="host")
    assert sd == sh and sd.count("Solved") >= 3 and "Failed" in sd
    assert np.array_equal(ed["info"][:, 0], eh["info"][:, 0])
    for b, st in enumerate(sd):
        if st != "Solved":
            assert hd[b] is None
            continue
        assert abs(ed["info"][b, 2] - eh["info"][b, 2]) <= 1e-7 * max(1.0, abs(eh["info"][b, 2]))
        # same problem bit for bit, but the solver's reductions use atomics: the two runs stop at slightly different points of
        # the optimal face (relative gap 2e-6), so x agrees only as far as the problem determines it
        assert np.abs(ed["x"][b] - eh["x"][b]).max() < 2e-4
        assert np.abs(hd[b] - hh[b]).max() < 5e-3
        assert np.abs(hd[b] - fir.fmp2(fir._x_to_r(ed["x"][b], n))).max() < 1e-12      # the device chain = the separate call


@pytest.mark.gpu
def test_one_call_path_vs_highs_n256(mbrf):
    """N = 256 through the one-call path, straight at the C ABI: HiGHS known answers at the reference's weights (obj up to 1e5,
    widened band edges), objective and violation recomputed on the CPU from the returned x."""
    from multiband_rf_pulse_design_b200 import fir
    known = json.load(open(os.path.join(ROOT, "tests", "golden", "fir_ap_weights_known.json")))
    feas = [k for k in known.values() if k["status"] == 0]
    x, h, info, rows = fir._solve_batch_ap_device(256, [k["f"] for k in feas], [k["a"] for k in feas], [k["d"] for k in feas],
                                                  [k["obj"] for k in feas], [k["peak"] for k in feas])
    assert rows[0] >= 7686 and rows[1] > 0
    for b, k in enumerate(feas):
        assert info[b, 0] == 1.0
        p = O.build_fir_ap(k["n"], k["f"], k["a"], k["d"], k["obj"], k["peak"])
        z = np.concatenate([x[b], [info[b, 7]]])
        assert abs(p["c"] @ z - k["cone_free_obj"]) <= 1e-5 * k["cone_free_obj"]
        assert O.violation_fir_ap(p, z) <= 1e-6
        r = np.abs(np.fft.fft(h[b], 8192)) ** 2                            # |H|^2 of the returned taps = the solved spectrum
        S = np.real(np.fft.fft(np.concatenate([fir._x_to_r(x[b], 256)[255:], np.zeros(8192 - 511), fir._x_to_r(x[b], 256)[:255]])))
        assert np.abs(r - S).max() < 5e-2 * S.max()                        # the cepstral factorisation on 4096 points is itself approximate
