"""GPU parity tests of the forward-SLR path, through the C ABI (mbrf_abr)."""
import numpy as np
import pytest

from conftest import TOL_SLR, golden, golden_files

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_files("abrx_*.npz"))
def test_abrx_golden(mbrf, name):
    g = golden(name)
    y = g["y"] if ("use_y" in g and int(g["use_y"])) else None
    a, b = mbrf.abrx(g["rf"], g["g"], g["x"], y)
    assert a.shape == g["alpha"].shape
    assert np.abs(a - g["alpha"]).max() < TOL_SLR and np.abs(b - g["beta"]).max() < TOL_SLR


def test_conventions_vs_oracle(mbrf, oracle):
    rng = np.random.default_rng(21)
    ns = 333
    rf = rng.normal(0, 0.03, ns) + 1j * rng.normal(0, 0.03, ns)
    g = rng.normal(0, 0.2, ns) + 1j * rng.normal(0, 0.2, ns)
    x = np.linspace(-9, 9, 301)
    y = np.linspace(-2, 2, 7)
    a, b = mbrf.abrx(rf, g, x, y)
    ao, bo = oracle.abrx_oracle(rf, g, x, y)
    assert np.abs(a - ao).max() < TOL_SLR and np.abs(b - bo).max() < TOL_SLR
    a, b = mbrf.abrm(rf, g, x, y)
    ao, bo = oracle.abrm_oracle(rf, g, x, y)
    assert np.abs(a - ao).max() < TOL_SLR and np.abs(b - bo).max() < TOL_SLR
    a, b = mbrf.abr(rf, g.real, x)                       # abr.m:34
    ao, bo = oracle.abrx_oracle(rf, g.real, x)
    assert np.abs(a - ao).max() < TOL_SLR and np.abs(b + np.conj(bo)).max() < TOL_SLR
    assert np.abs(np.abs(a) ** 2 + np.abs(b) ** 2 - 1).max() < 1e-12
    # 3-argument call ignores the imaginary (y) gradient, like abrx.c:50
    a3, b3 = mbrf.abrx(rf, g, x)
    a4, b4 = mbrf.abrx(rf, g.real, x)
    assert np.array_equal(a3, a4) and np.array_equal(b3, b4)
    # default gradient of the 2-argument forms (abr.m:25, abrm.m:28)
    a2, b2 = mbrf.abr(rf, x)
    a5, b5 = mbrf.abr(rf, np.ones(ns) * 2 * np.pi / ns, x)
    assert np.array_equal(a2, a5) and np.array_equal(b2, b5)


def test_edge_cases(mbrf, oracle):
    # big rotations per sample: every tier
    rng = np.random.default_rng(22)
    ns = 40
    rf = (rng.normal(0, 1, ns) + 1j * rng.normal(0, 1, ns)) * np.linspace(0.01, 4, ns)
    g = np.ones(ns)
    x = np.concatenate([np.linspace(-0.5, 0.5, 5), np.linspace(-30, 30, 7)])
    a, b = mbrf.abrx(rf, g, x)
    ao, bo = oracle.abrx_oracle(rf, g, x)
    assert np.abs(a - ao).max() < TOL_SLR and np.abs(b - bo).max() < TOL_SLR
    # phi == 0 samples: abrx keeps going (abrx.c:94-98), abrm returns NaN (abrm.m:49-50, 0/0)
    rf0 = np.array([0.0, 0.1, 0.0])
    a, b = mbrf.abrx(rf0, np.ones(3), np.array([0.0, 1.0]))
    ao, bo = oracle.abrx_oracle(rf0, np.ones(3), np.array([0.0, 1.0]))
    assert np.abs(a - ao).max() < 1e-14 and np.abs(b - bo).max() < 1e-14
    a, b = mbrf.abrm(rf0, np.ones(3), np.array([0.0, 1.0]))
    ao, bo = oracle.abrm_oracle(rf0, np.ones(3), np.array([0.0, 1.0]))
    assert np.isnan(a[0, 0]) and np.isnan(ao[0, 0]) and abs(a[1, 0] - ao[1, 0]) < 1e-14
    # real rf (mxGetPi NULL, abrx.c:92) and a single position
    a, b = mbrf.abrx(np.full(16, 0.05), np.ones(16), np.array([0.3]))
    ao, bo = oracle.abrx_oracle(np.full(16, 0.05), np.ones(16), np.array([0.3]))
    assert abs(a[0, 0] - ao[0, 0]) < 1e-13 and abs(b[0, 0] - bo[0, 0]) < 1e-13


def test_large_grid_properties(mbrf):
    """10^6 positions: unitarity and the inversion-profile identity mz = 1 - 2|b|^2 against the Bloch path."""
    g = golden("pulses.npz")
    rf = g["rf512_rad"]
    ns = rf.size
    grad = np.ones(ns) * 2 * np.pi / ns
    x = np.linspace(-40, 40, 1000)
    y = np.linspace(-1, 1, 1000)          # no y gradient: every column identical
    a, b = mbrf.abr(rf, grad, x, y)
    assert a.shape == (1000, 1000)
    assert np.abs(np.abs(a) ** 2 + np.abs(b) ** 2 - 1).max() < 1e-12
    assert np.abs(a - a[:, :1]).max() == 0.0
    gamma, dt = mbrf.GAMMA_C13, 1e-5
    df = -x * (2 * np.pi / ns) / (6.283185 * dt)
    mx, my, mz = mbrf.blochC(rf / (gamma * dt), np.zeros(ns), dt, 1e30, 1e30, df, 0.0)
    assert np.abs(mz.ravel() - (1 - 2 * np.abs(b[:, 0]) ** 2)).max() < 1e-10
