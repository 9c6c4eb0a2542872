"""CPU tests: pin the oracle restatement to the reference's own C and to the golden fixtures."""
import os

import numpy as np
import pytest

from conftest import golden, golden_files


def _bloch_args(g):
    args = [g["b1"].reshape(-1, 1), g["gr"], float(g["dt"]), float(g["t1"]), float(g["t2"]), g["df"], g["dp"],
            int(g["mode"])]
    if int(g["use_m0"]):
        args += [g["m0"][0], g["m0"][1], g["m0"][2]]
    return args


def _oracle_from_mex_args(oracle, nucleus, b1, gr, tp, t1, t2, df, dp, mode=0, mx=None, my=None, mz=None):
    """numpy restatement of the gateway's argument handling (blochC.c:571-865) on top of oracle_blochsimfz."""
    b1 = np.asarray(b1).ravel(order="F")
    nt = b1.size
    gr = np.asarray(gr, dtype=float).ravel(order="F")
    gx = gr[:nt]
    gy = gr[nt:2 * nt] if gr.size >= 2 * nt else None
    gz = gr[2 * nt:3 * nt] if gr.size >= 3 * nt else None
    tp = np.asarray(tp, dtype=float).ravel()
    if tp.size == 1:
        dt = np.full(nt, tp[0])
    else:
        iv = np.diff(np.concatenate([[0.0], tp]))
        dt = iv if np.all(iv > 0) else tp
    dp = np.atleast_2d(np.asarray(dp, dtype=float))
    if dp.shape[1] == 3:
        dx, dy, dz = dp[:, 0], dp[:, 1], dp[:, 2]
    elif dp.shape[1] == 2:
        dx, dy, dz = dp[:, 0], dp[:, 1], None
    else:
        dx, dy, dz = dp.ravel(order="F"), None, None
    df = np.asarray(df, dtype=float).ravel()
    m0 = None
    if mx is not None and np.size(mx) == np.size(my) == np.size(mz) == df.size * dx.size:
        m0 = [np.ravel(mx, order="F"), np.ravel(my, order="F"), np.ravel(mz, order="F")]
    gamma = oracle.GAMMA_C13 if nucleus == "C-13" else oracle.GAMMA_H1
    return oracle.blochsimfz_oracle(b1, gx, gy, gz, dt, t1, t2, df, dx, dy, dz, mode, m0, gamma)


@pytest.mark.parametrize("name", golden_files("bloch_rand_*.npz"))
def test_restatement_bitexact_vs_golden(oracle, name):
    g = golden(name)
    out = _oracle_from_mex_args(oracle, str(g["nucleus"]), *_bloch_args(g))
    for o, k in zip(out, ("mx", "my", "mz")):
        assert np.array_equal(o, g[k].ravel(order="F")), f"{name}:{k}"


@pytest.mark.parametrize("name", golden_files("bloch_time_*.npz"))
def test_time_vector_semantics(oracle, name):
    g = golden(name)
    out = _oracle_from_mex_args(oracle, "C-13", g["b1"], g["gr"], g["tp"], float(g["t1"]), float(g["t2"]), g["df"],
                                g["dp"], 0)
    for o, k in zip(out, ("mx", "my", "mz")):
        assert np.array_equal(o, g[k].ravel(order="F"))


def test_cfg1_and_bigangle_golden(oracle):
    g = golden("bloch_cfg1.npz")
    out = _oracle_from_mex_args(oracle, "C-13", g["b1"], np.zeros(256), float(g["dt"]), 1e3, 1e3, g["df"], 0.0, 0)
    for o, k in zip(out, ("mx", "my", "mz")):
        assert np.array_equal(o, g[k].ravel(order="F"))
    mxy = g["mx"].ravel() + 1j * g["my"].ravel()
    assert abs(abs(mxy[1000]) - 0.99998) < 1e-4          # pi/2 excitation on resonance
    assert np.abs(mxy[:200]).max() < 1e-3                # stop band
    g = golden("bloch_bigangle.npz")
    out = _oracle_from_mex_args(oracle, "C-13", g["b1"], np.zeros(g["b1"].size), float(g["dt"]), 1e3, 1e3, g["df"],
                                0.0, 0)
    for o, k in zip(out, ("mx", "my", "mz")):
        assert np.array_equal(o, g[k].ravel(order="F"))


def test_restatement_vs_live_reference(oracle):
    """Where the compiled reference is present, compare on fresh random inputs too (modes 0 and 2)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    rng = np.random.default_rng(7)
    nt = 80
    b1 = rng.normal(0, 0.05, nt) + 1j * rng.normal(0, 0.05, nt)
    g = rng.normal(0, 0.5, (nt, 3))
    dt = rng.uniform(4e-6, 2e-5, nt)
    df = rng.uniform(-3000, 3000, 11)
    dp = rng.uniform(-3, 3, (4, 3))
    for mode in (0, 2):
        for nuc, gam in (("C-13", oracle.GAMMA_C13), ("H-1", oracle.GAMMA_H1)):
            a = oracle.blochsimfz_oracle(b1, g[:, 0], g[:, 1], g[:, 2], dt, 0.5, 0.05, df, dp[:, 0], dp[:, 1],
                                         dp[:, 2], mode, gamma=gam)
            b = oracle.blochsimfz_ref(b1, g[:, 0], g[:, 1], g[:, 2], dt, 0.5, 0.05, df, dp[:, 0], dp[:, 1],
                                      dp[:, 2], mode, nucleus=nuc)
            for u, v in zip(a, b):
                assert np.array_equal(u, v)


def test_steady_state_modes_properties(oracle):
    """Modes 1/3 cannot be pinned to the reference (UB at blochC.c:132); pin them by what they must satisfy."""
    rng = np.random.default_rng(3)
    nt = 40
    b1 = rng.normal(0, 0.05, nt) + 1j * rng.normal(0, 0.05, nt)
    gx = rng.normal(0, 0.3, nt)
    dt = np.full(nt, 2e-5)
    df = rng.uniform(-500, 500, 5)
    dx = rng.uniform(-1, 1, 3)
    a = (b1, gx, None, None, dt, 0.3, 0.04, df, dx)
    m1 = oracle.blochsimfz_oracle(*a, mode=1)
    m3 = oracle.blochsimfz_oracle(*a, mode=3)
    for c in range(3):  # last recorded sample of mode 3 is the steady state again
        assert np.allclose(m3[c][nt - 1::nt], m1[c], atol=1e-12)
    # steady state is a fixed point of one period of the transient simulation
    again = oracle.blochsimfz_oracle(*a, mode=0, m0=m1)
    for c in range(3):
        assert np.allclose(again[c], m1[c], atol=1e-12)


@pytest.mark.parametrize("name", golden_files("abrx_*.npz"))
def test_abrx_restatement_vs_golden(oracle, name):
    g = golden(name)
    y = g["y"] if ("use_y" in g and int(g["use_y"])) else None
    a, b = oracle.abrx_oracle(g["rf"], g["g"], g["x"], y)
    assert np.array_equal(a, g["alpha"]) and np.array_equal(b, g["beta"])


def test_abrm_identities(oracle):
    """abrm(rf,g,x) == (a_abrx(-x), conj(b_abrx(-x))) and |a|^2+|b|^2 == 1."""
    rng = np.random.default_rng(5)
    ns = 150
    rf = rng.normal(0, 0.03, ns) + 1j * rng.normal(0, 0.03, ns)
    g = rng.normal(0, 0.2, ns) + 1j * rng.normal(0, 0.2, ns)
    x = np.linspace(-5, 5, 21)
    y = np.linspace(-2, 2, 4)
    am, bm = oracle.abrm_oracle(rf, g, x, y)
    ax, bx = oracle.abrx_oracle(rf, g, -x, -y)
    assert np.abs(am - ax).max() < 1e-14 and np.abs(bm - np.conj(bx)).max() < 1e-14
    assert np.abs(np.abs(am) ** 2 + np.abs(bm) ** 2 - 1).max() < 1e-13
    # abrm has no phi == 0 guard (abrm.m:49-50): NaN
    a0, b0 = oracle.abrm_oracle(np.zeros(4), np.ones(4), np.zeros(1))
    assert np.isnan(a0).all() and np.isnan(b0).all()


def test_abrx_gateway_errors(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    rf = np.ones(8) * 0.1
    out, err = oracle.abrx_mex_ref(rf, np.ones(7), np.zeros(3))
    assert out is None and err == "rf and gradient vectors are of different lengths"   # abrx.c:45
    out, err = oracle.abrx_mex_ref(rf, np.ones(8), np.zeros(3), nlhs=1)
    assert out is None and err.startswith("Usage:")                                     # abrx.c:41


def test_slr_matches_bloch_without_relaxation(oracle):
    """mz = 1-2|b|^2, mxy = 2 conj(a) b (abr.m:10-13) must agree with the Bloch simulator when T1,T2 -> inf."""
    g = golden("pulses.npz")
    rf = g["rf256_rad"]
    n = rf.size
    x = np.linspace(-6, 6, 41)
    grad = np.ones(n) * 2 * np.pi / n
    a, b = oracle.abrx_oracle(rf, grad, x)
    b = -np.conj(b)                                                    # abr.m:34
    mz_slr = 1 - 2 * np.abs(b[:, 0]) ** 2
    # same rotations through blochsimfz: rotx = -b1*gamma*dt = -rf (sign convention differs: compare mz only)
    gamma, dt = oracle.GAMMA_C13, 1e-5
    b1 = rf / (gamma * dt)
    df = -x * (2 * np.pi / n) / (6.283185 * dt)                         # rotz = -df*TWOPI*dt = x*g
    m = oracle.blochsimfz_oracle(b1, None, None, None, dt, 1e30, 1e30, df, np.zeros(1))
    assert np.abs(m[2] - mz_slr).max() < 1e-12


@pytest.mark.parametrize("tb,flip", [(2.0, np.pi / 2), (4.0, 0.6), (8.0, 2.5)])
def test_inverse_slr_restatement_pinned_to_the_reference_c(oracle, tb, flip):
    """The numpy restatements of rf_tools/b2a.m + ab2rf.m (the oracle of the GPU inverse SLR, tests/test_islr_gpu.py) against
    the UNMODIFIED reference C pair b2a.code.c + cabc2rf.code.c run through the reference's own b2rf mexFunction
    (oracle/_ref/libb2rf.so): real beta, n = 256 (the compiled reference overflows its static arrays for n >= 512,
    b2a.code.c:16-17,33-38).  The .m and .c variants pad differently (x8 vs x16) yet agree to 1e-15."""
    if not os.path.exists(os.path.join(oracle.REF_DIR, "libb2rf.so")):
        pytest.skip("oracle/_ref/libb2rf.so not built (needs /root/reference)")
    n = 256
    k = np.arange(n) - (n - 1) / 2
    b = np.sinc(k * tb / n) * np.hamming(n)
    b = b / np.abs(np.fft.fft(b, 8 * n)).max() * np.sin(flip / 2)
    rf_c = oracle.b2rf_ref(b)
    rf_m = oracle.ab2rf_m(oracle.b2a_m(b), b)
    assert np.abs(rf_m.imag).max() < 1e-15
    assert np.abs(np.real(rf_c) - rf_m.real).max() < 1e-14
