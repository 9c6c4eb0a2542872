"""Specification builders (multiband_rf_pulse_design_b200/spec.py: rf_ripple_GFA, rf_Mrange_desired, rf_bandedge, dinf,
spectrum_C13, the specification part of dzrf_mb.m) against the scalar restatement, the reference's own numeric self-check
(rf_ripple_GFA.m:42-79) and the dual-band H-1 specification probed in SURVEY.md 8(d).  Host logic: no GPU."""
import numpy as np
import pytest

from multiband_rf_pulse_design_b200 import spec
from oracle import spec_reference as R
from oracle.fir_problems import H1_DUALBAND


@pytest.mark.parametrize("ptype", ["ex", "sat", "inv", "se"])
def test_ripple_ranges_match_the_scalar_restatement_and_the_self_check(ptype):
    rng = np.random.default_rng(0)
    FA = np.concatenate([rng.uniform(0, 180, 300), [0.0, 90.0, 180.0, 60.0, 120.0]])
    rip = np.concatenate([10.0 ** rng.uniform(-4, -0.7, 300), [0.001, 0.05, 0.05, 0.01, 0.05]])
    rB, dB = spec.rf_ripple_GFA(FA, rip, ptype)
    rM = spec.rf_Mrange_desired(FA, rip, ptype)
    assert rB.shape == (FA.size, 2) and dB.shape == (FA.size, 2)
    for i in range(FA.size):
        want_r, want_d = R.ripple_asin(FA[i], rip[i], ptype)
        assert np.allclose(rB[i], want_r, rtol=0, atol=1e-15) and np.allclose(dB[i], want_d, rtol=0, atol=1e-15)
        assert np.allclose(rM[i], R.mrange(FA[i], rip[i], ptype), rtol=0, atol=1e-15)
        # the reference's own check (dbg >= 1): the magnetisation reached over range_B is the desired range.  It holds where the
        # map |beta| -> M is monotonic over the range (away from the 90-degree fold of 'ex', and for FA +- range inside [0, 180])
        lo, hi = R.measured_range(rB[i], ptype)
        des = rM[i]
        fold = ptype == "ex" and np.sin(np.deg2rad(FA[i])) + rip[i] >= 1
        clipped = des[0] < (0 if ptype in ("ex", "se") else -1) + 1e-12
        if not fold and not clipped and rB[i, 0] >= 0:
            assert abs(lo - des[0]) < 2e-6 and abs(hi - des[1]) < 2e-6, (ptype, FA[i], rip[i])


def test_quadratic_approximation_is_close_for_small_ripples():
    for ptype, FA in (("ex", 30.0), ("ex", 60.0), ("sat", 90.0), ("sat", 120.0), ("inv", 150.0)):
        for rip in (1e-3, 1e-2):
            got, _ = spec.rf_ripple_GFA(FA, rip, ptype, appro=1)
            want, _ = R.ripple_quad(FA, rip, ptype)
            exact, _ = R.ripple_asin(FA, rip, ptype)
            assert np.allclose(got, want, atol=1e-14)
            assert np.allclose(got, exact, atol=20 * rip ** 2 + 1e-9)


def test_errors_like_the_reference():
    with pytest.raises(ValueError, match="range of"):
        spec.rf_ripple_GFA(190, 0.01, "ex")
    with pytest.raises(ValueError, match="Unrecognized Pulse Type"):
        spec.rf_ripple_GFA(90, 0.01, "xx")
    with pytest.raises(ValueError, match="not monotonically increasing"):
        spec.rf_bandedge(200, 0.02, [0.0, 0.05], [0.2, 0.2], [60, 0], [0.01, 0.005], "ex")      # rf_bandedge.m:146-151
    with pytest.raises(ValueError, match="sampling rate is not enough"):
        spec.rf_bandedge(20, 1.0, [-0.6, 0.6], [0.1, 0.1], [60, 0], [0.01, 0.005], "ex")        # :153-155
    with pytest.raises(ValueError, match="not an integer"):
        spec.multiband_spec(201, 0.02, [0.0, 1.0], [0.1, 0.1], [60, 0], [0.01, 0.005], "ex", downsampling=2)


def test_h1_dualband_spec_known_answer():
    """specsat_H1_dualband.m:5-32 -> dzrf_mb.m:92-147 (ptype 'sat', shift_f = 1): the specification SURVEY.md 8(d) probed from
    the reference's scripts, six digits -- the constants every N = 256 solver test of this repo uses."""
    n, B0, T, d1, d2 = 260, 127794577 / (42.577 * 1e6), 26, 0.05, 0.001
    ppm = [np.array([1.8, 2.5]), np.array([3, 4.1]), np.array([4.8, 5.4])]
    ref = ppm[2].mean()
    mb_cf = [(c - ref) * B0 * 42.577 * 1e-3 for c in ppm]                   # kHz
    out = spec.multiband_spec(n, T / n, mb_cf, [0.01, 0.01, 0.01], [120, 0, 90], [d1, d2, d1], "sat", shift_f=1)
    assert np.abs(out["f"] - H1_DUALBAND["f"]).max() < 5e-7
    assert np.abs(out["a"] - H1_DUALBAND["a"]).max() < 5e-7
    assert np.abs(out["d"] - H1_DUALBAND["d"]).max() < 5e-7
    assert np.all(np.diff(out["f"]) > 0) and abs(out["b_spec"]["f"] - out["f"] - (out["b_spec"]["f"][0] - out["f"][0])).max() < 1e-15
    assert out["shift_f_back"] == pytest.approx((out["b_spec"]["f"][0] - out["f"][0]) * 0.5 / (T / n))


def test_bandedge_without_ranges_fills_the_axis():
    """mb_range = [] (rf_bandedge.m:36-131): neighbouring bands meet half-way between the given ranges, minus the transition
    width df = dinf(delta1, delta2) / T."""
    n, dt = 400, 0.02
    mb_cf = [[-2.0, -1.0], [0.0, 0.5], [2.0, 3.0]]
    fn = spec.rf_bandedge(n, dt, mb_cf, None, [0, 90, 0], [0.005, 0.01, 0.005], "ex")
    f = fn * (0.5 / dt)
    df = float(spec.dinf(np.sqrt(0.01 / 2), 0.005 / np.sqrt(2))) / (n * dt)
    assert np.all(np.diff(f) > 0)
    assert f[2] - f[1] == pytest.approx(df) and f[4] - f[3] == pytest.approx(df)
    assert (f[1] + f[2]) / 2 == pytest.approx(-0.5) and (f[3] + f[4]) / 2 == pytest.approx(1.25)


def test_c13_bssfp_spec_of_the_bench_comes_from_these_builders():
    import bench
    f, a, d, dt = bench.c13_bssfp_spec()
    assert dt == pytest.approx(0.02) and f.size == 10 and a.size == 10 and d.size == 5
    fr, names = spec.spectrum_C13(14.0)
    assert names[5] == "Urea" and fr[0] == 0.0
    assert a[0] == pytest.approx(np.sin(np.pi / 6), abs=1e-3) and np.all(a[2:] == a[2]) and d[1] == pytest.approx(0.0025, abs=1e-6)
