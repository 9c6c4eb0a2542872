"""Batched inverse SLR on the GPU (csrc/islr.cu: mbrf_b2a_batch, mbrf_ab2rf_batch) against the numpy restatements of
rf_tools/b2a.m:13-28 and rf_tools/ab2rf.m:12-26 (oracle/ref.py), and closed through the forward SLR kernel:
abr(b2rf(b)) reproduces the beta polynomial's profile."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def _betas(n, B, seed, complex_b):
    rng = np.random.default_rng(seed)
    k = np.arange(n) - (n - 1) / 2
    out = []
    for q in range(B):
        tb = 2.0 + 6.0 * rng.random()
        b = np.sinc(k * tb / n) * np.hamming(n)
        b = b / np.abs(np.fft.fft(b, 8 * n)).max() * np.sin(0.5 * (0.2 + 2.6 * rng.random()))     # flip angle 0.2 .. 2.8 rad
        if complex_b:
            b = b * np.exp(1j * 0.3 * k * rng.random())
        out.append(b)
    return np.stack(out)


@pytest.mark.parametrize("n", [16, 64, 256, 512, 48, 260, 500])      # 2^k: radix-2; others: Bluestein (dzrf_mb uses n = 260)
@pytest.mark.parametrize("complex_b", [False, True])
def test_b2a_and_ab2rf_match_reference(mbrf, oracle, n, complex_b):
    bc = _betas(n, 5, n, complex_b)
    a = mbrf.b2a(bc)
    rf = mbrf.ab2rf(a, bc)
    for q in range(bc.shape[0]):
        a_ref = oracle.b2a_m(bc[q])
        assert np.abs(a[q] - a_ref).max() <= 1e-11, np.abs(a[q] - a_ref).max()
        rf_ref = oracle.ab2rf_m(a_ref, bc[q])
        assert np.abs(rf[q] - rf_ref).max() <= 1e-9, np.abs(rf[q] - rf_ref).max()
    if not complex_b:   # |alpha|^2 + |beta|^2 = 1 on the unit circle (what b2a constructs); real b: symmetric magnitude spectra
        A = np.fft.fft(a[:, ::-1], 8 * n, axis=1)
        Bf = np.fft.fft(bc, 8 * n, axis=1)
        assert np.abs(np.abs(A) ** 2 + np.abs(Bf) ** 2 - 1).max() <= 1e-6


def test_b2a_rescales_unphysical_beta(mbrf, oracle):
    """b2a.m:23-25: max|B| >= 1 is scaled to just below 1."""
    bc = _betas(64, 1, 3, False)[0] * 3.0
    assert np.abs(np.fft.fft(bc, 512)).max() >= 1.0
    assert np.abs(mbrf.b2a(bc) - oracle.b2a_m(bc)).max() <= 1e-9


def test_inverse_then_forward_slr_round_trip(mbrf):
    """rf = b2rf(b); the forward kernel abr(rf, x) must give back |beta(x)| = |sum_k b_k exp(-j x k)|.  abrx.c:93-101 applies
    RF and gradient of a sample as ONE rotation about the tilted axis, ab2rf.m inverts the hard-pulse model (nutation, then
    precession), so the two agree to first order in rf*x per sample, not exactly: 5e-3 of a 0.6 passband here; a convention
    error (ordering, conjugation, angle scale) would show up at the 0.1 level."""
    n = 256
    k = np.arange(n) - (n - 1) / 2
    bc = np.sinc(k * 8.0 / n) * np.hamming(n)
    bc = bc / np.abs(np.fft.fft(bc, 8 * n)).max() * np.sin(0.5 * 1.3)       # 1.3 rad flip
    rf = mbrf.b2rf(bc)
    assert abs(rf.real.sum() - 1.3) < 2e-2 and np.abs(rf.imag).max() < 1e-9  # small-tip check: sum(rf) = flip angle
    x = np.linspace(-np.pi, np.pi, 513)
    a, b = mbrf.abr(rf, np.ones(n), x)
    want = np.abs((bc[None, :] * np.exp(-1j * x[:, None] * np.arange(n)[None, :])).sum(1))
    assert want.max() > 0.55
    assert np.abs(np.abs(b.ravel()) - want).max() <= 5e-3
    assert np.abs(np.abs(a.ravel()) ** 2 + np.abs(b.ravel()) ** 2 - 1).max() <= 1e-12


def test_islr_argument_checks(mbrf):
    with pytest.raises(mbrf.MbrfError):
        mbrf.b2a(np.ones(600) * 1e-3)                      # not a power of two and > 512
    with pytest.raises(mbrf.MbrfError):
        mbrf.b2a(np.ones(2048) * 1e-3)                     # beyond the shared-memory transform
    with pytest.raises(ValueError):
        mbrf.ab2rf(np.ones(8), np.ones(9))
