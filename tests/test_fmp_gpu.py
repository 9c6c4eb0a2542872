"""Batched spectral factorisation on the GPU (csrc/fmp.cu, mbrf_fmp2_batch) against the numpy restatement of
fir_ap_cvx.m:253-304 (oracle/fir_problems.py) and against the defining property |H_mp(w)|^2 = R(w)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def _autocorr(rng, n, complex_taps):
    h = rng.standard_normal(n) * np.exp(-np.arange(n) / (0.3 * n))
    if complex_taps:
        h = h + 1j * rng.standard_normal(n) * np.exp(-np.arange(n) / (0.3 * n))
    r = np.correlate(h, h, mode="full")                 # r_k = sum h[m+k] conj(h[m]), length 2n-1, Hermitian
    r[n - 1] *= 1.0 + 1e-6                              # strictly positive spectrum
    return r


@pytest.mark.parametrize("n", [5, 64, 256, 260, 512])
@pytest.mark.parametrize("complex_taps", [False, True])
def test_fmp2_matches_reference(mbrf, n, complex_taps):
    from oracle.fir_problems import fmp2_reference
    rng = np.random.default_rng(n + int(complex_taps))
    B = 7
    R = np.stack([_autocorr(rng, n, complex_taps) for _ in range(B)])
    H = mbrf.fir.fmp2_batch(R)
    assert H.shape == (B, n)
    for b in range(B):
        ref = fmp2_reference(R[b])
        assert np.abs(H[b] - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (b, np.abs(H[b] - ref).max())
    # single-sequence entry point of the mirror
    assert np.abs(mbrf.fmp2(R[0]) - H[0]).max() == 0.0


def test_fmp2_is_a_spectral_factor(mbrf):
    """|H_mp(w)|^2 reproduces the spectrum of r (up to the truncation to n taps) and H_mp is minimum phase."""
    rng = np.random.default_rng(3)
    n = 128
    r = _autocorr(rng, n, False)
    h = mbrf.fmp2(r)
    w = np.linspace(-np.pi, np.pi, 1024, endpoint=False)
    k = np.arange(-(n - 1), n)
    Rw = (r[None, :] * np.exp(-1j * w[:, None] * k[None, :])).sum(1).real
    Hw = (h[None, :] * np.exp(-1j * w[:, None] * np.arange(n)[None, :])).sum(1)
    # fmp2 is a finite (8x padded) FFT / Hilbert-transform approximation and keeps n taps: ~1 % (same bound as test_fir_gpu)
    assert np.abs(np.abs(Hw) ** 2 - Rw).max() <= 2e-2 * Rw.max()
    assert np.abs(np.roots(h)).max() <= 1.0 + 1e-3


def test_fmp2_limits(mbrf):
    lib = mbrf.lib()
    assert lib.mbrf_fmp2_max_taps() == 512
    with pytest.raises(ValueError):
        mbrf.fir.fmp2_batch(np.ones((2, 8)))            # even length: "filter length must be odd"
    with pytest.raises(mbrf.MbrfError):
        mbrf.fir.fmp2_batch(np.ones((1, 2 * 513 - 1)))  # beyond the shared-memory kernel


def test_fir_ap_cvx_returns_the_reference_taps(mbrf):
    """h returned by fir_ap_cvx == fmp2 (numpy restatement) of the solver's x: the whole return path of fir_ap_cvx.m:185-202."""
    from oracle.fir_problems import x_to_h_reference
    f = [-0.6, -0.35, -0.2, 0.18, 0.38, 0.6]
    hs, st, ex = mbrf.fir.fir_ap_cvx_batch(40, [f, f], [0.866, 0.866, 0, 0, 0.707, 0.707], [0.02, 0.03, 0.025], [0.1, 1.0],
                                           [10 ** -1.5] * 2, return_info=True, max_iter=40000)
    assert st == ["Solved", "Solved"]
    for b in range(2):
        ref = x_to_h_reference(ex["x"][b], 40)
        assert np.abs(hs[b] - ref).max() <= 1e-10


@pytest.mark.gpu
def test_fmp2_kernel_recovers_known_minimum_phase_filters(mbrf):
    """Known answer (not a comparison with the restatement): the spectral factor of the autocorrelation of a filter with all
    zeros well inside the unit circle is the filter itself (tests/test_oracle_fir.py pins the restatement the same way)."""
    from test_oracle_fir import MINPHASE_CASES, _minimum_phase_case
    from multiband_rf_pulse_design_b200 import fir
    for n, rad, tol in MINPHASE_CASES:
        cases = [_minimum_phase_case(n, rad, seed=n + 100 * k) for k in range(4)]
        H = fir.fmp2_batch(np.stack([r for _, r in cases]))
        for (h0, _), h in zip(cases, H):
            assert np.abs(h - h0).max() < tol
