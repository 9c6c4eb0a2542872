"""GPU parity tests of the Bloch path, through the C ABI (mbrf_bloch / mbrf_blochsimfz)."""
import numpy as np
import pytest

from conftest import TOL_BLOCH, golden, golden_files

pytestmark = pytest.mark.gpu


def _args(g):
    a = [g["b1"].reshape(-1, 1), g["gr"], float(g["dt"]), float(g["t1"]), float(g["t2"]), g["df"], g["dp"],
         int(g["mode"])]
    if int(g["use_m0"]):
        a += [g["m0"][0], g["m0"][1], g["m0"][2]]
    return a


@pytest.mark.parametrize("name", golden_files("bloch_rand_*.npz"))
def test_golden_random(mbrf, name):
    """Outputs of the reference's own mexFunction on seeded inputs (modes 0/2, both nuclei, 1-3 axes, M0)."""
    g = golden(name)
    fn = mbrf.blochC if str(g["nucleus"]) == "C-13" else mbrf.blochH
    out = fn(*_args(g))
    for o, k in zip(out, ("mx", "my", "mz")):
        assert o.shape == g[k].shape, (o.shape, g[k].shape)          # blochC.c:880-904
        assert np.abs(o - g[k]).max() < TOL_BLOCH


@pytest.mark.parametrize("name", golden_files("bloch_time_*.npz"))
def test_golden_time_vector(mbrf, name):
    g = golden(name)
    out = mbrf.blochC(g["b1"], g["gr"], g["tp"], float(g["t1"]), float(g["t2"]), g["df"], g["dp"], 0)
    for o, k in zip(out, ("mx", "my", "mz")):
        assert o.shape == g[k].shape
        assert np.abs(o - g[k]).max() < TOL_BLOCH


def test_golden_cfg1(mbrf):
    """BASELINE config 1: dzrf 256-sample pulse, 2000 offsets."""
    g = golden("bloch_cfg1.npz")
    out = mbrf.blochC(g["b1"].reshape(-1, 1), np.zeros((256, 1)), float(g["dt"]), 1e3, 1e3, g["df"].reshape(-1, 1),
                      0.0, 0)
    for o, k in zip(out, ("mx", "my", "mz")):
        assert o.shape == g[k].shape
        assert np.abs(o - g[k]).max() < TOL_BLOCH


def test_golden_big_rotations(mbrf):
    """Rotations from ~0 to several turns per sample: all polynomial tiers and the sqrt/sincos path."""
    g = golden("bloch_bigangle.npz")
    out = mbrf.blochC(g["b1"].reshape(-1, 1), np.zeros((g["b1"].size, 1)), float(g["dt"]), 1e3, 1e3, g["df"], 0.0, 0)
    for o, k in zip(out, ("mx", "my", "mz")):
        assert np.abs(o - g[k]).max() < TOL_BLOCH


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("ngrad", [0, 1, 3])
def test_vs_oracle_all_modes(mbrf, oracle, mode, ngrad):
    rng = np.random.default_rng(100 + 10 * mode + ngrad)
    nt, nf, npos = 300, 37, 11          # more than two tiles, ragged tail, not a multiple of the block
    b1 = rng.normal(0, 0.05, nt) + 1j * rng.normal(0, 0.05, nt)
    gr = rng.normal(0, 0.4, (nt, max(ngrad, 1))) * (ngrad > 0)
    dt = rng.uniform(4e-6, 2e-5, nt)
    df = rng.uniform(-3000, 3000, nf)
    dp = rng.uniform(-3, 3, (npos, max(ngrad, 1)))
    t1, t2 = 0.5, 0.05
    g3 = [gr[:, i] if i < gr.shape[1] else None for i in range(3)]
    d3 = [dp[:, i] if i < dp.shape[1] else None for i in range(3)]
    want = oracle.blochsimfz_oracle(b1, g3[0], g3[1], g3[2], dt, t1, t2, df, d3[0], d3[1], d3[2], mode)
    got = mbrf.blochC(b1, gr, np.cumsum(dt), t1, t2, df, dp, mode)       # end times -> same intervals
    # cumsum/diff round-trips the intervals to ~1 ulp of the end time; that is input noise, not kernel error
    tol = TOL_BLOCH
    for o, w in zip(got, want):
        assert np.abs(o.ravel(order="F") - w).max() < tol


def test_inner_abi_in_place(mbrf, oracle):
    """mbrf_blochsimfz: M pre-seeded at stride ntout, updated in place (blochC.c:422-426, :838-865)."""
    rng = np.random.default_rng(11)
    nt, nf, npos = 130, 9, 6
    b1 = rng.normal(0, 0.05, nt) + 1j * rng.normal(0, 0.05, nt)
    gx, gy, gz = (rng.normal(0, 0.4, nt) for _ in range(3))
    dt = np.full(nt, 1e-5)
    df = rng.uniform(-3000, 3000, nf)
    dx, dy, dz = (rng.uniform(-3, 3, npos) for _ in range(3))
    m0 = rng.normal(0, 0.5, (3, nf * npos))
    for mode in (0, 2):
        ntout = nt if mode else 1
        bufs = [np.zeros(nf * npos * ntout) for _ in range(3)]
        for c in range(3):
            bufs[c][::ntout] = m0[c]
        mbrf.blochsimfz(b1.real, b1.imag, gx, gy, gz, dt, 0.4, 0.06, df, dx, dy, dz, *bufs, mode,
                        gamma=mbrf.GAMMA_H1)
        want = oracle.blochsimfz_oracle(b1, gx, gy, gz, dt, 0.4, 0.06, df, dx, dy, dz, mode, m0, oracle.GAMMA_H1)
        for o, w in zip(bufs, want):
            assert np.abs(o - w).max() < TOL_BLOCH


def test_edge_cases(mbrf, oracle):
    b1 = np.array([0.02 + 0.01j])
    # single sample, single spin
    out = mbrf.blochC(b1, np.zeros(1), 1e-4, 1.0, 0.1, np.array([30.0]), 0.0)
    want = oracle.blochsimfz_oracle(b1, None, None, None, 1e-4, 1.0, 0.1, np.array([30.0]), np.zeros(1))
    assert all(abs(o.ravel()[0] - w[0]) < TOL_BLOCH for o, w in zip(out, want))
    # zero field everywhere: phi == 0 identity branch (blochC.c:182-193), pure relaxation
    out = mbrf.blochC(np.zeros(10), np.zeros(10), 1e-3, 0.05, 0.02, np.zeros(4), 0.0, 0,
                      np.full((1, 4), 0.3), np.full((1, 4), -0.2), np.full((1, 4), 0.1))
    want = oracle.blochsimfz_oracle(np.zeros(10), None, None, None, 1e-3, 0.05, 0.02, np.zeros(4), np.zeros(1),
                                    m0=[np.full(4, 0.3), np.full(4, -0.2), np.full(4, 0.1)])
    assert all(np.abs(o.ravel() - w).max() < 1e-14 for o, w in zip(out, want))
    # empty frequency list -> empty outputs, no launch
    out = mbrf.blochC(np.ones(4) * 0.01, np.zeros(4), 1e-5, 1.0, 1.0, np.zeros(0), 0.0)
    assert all(o.size == 0 for o in out)
    # wrong-size initial magnetisation silently falls back to (0,0,1) (blochC.c:851-865)
    a = mbrf.blochC(b1, np.zeros(1), 1e-4, 1.0, 0.1, np.array([30.0, 10.0]), 0.0, 0, np.zeros(3), np.zeros(3),
                    np.ones(3))
    b = mbrf.blochC(b1, np.zeros(1), 1e-4, 1.0, 0.1, np.array([30.0, 10.0]), 0.0)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    # NaN in, NaN out (no silent masking)
    out = mbrf.blochC(np.array([np.nan, 0.01]), np.zeros(2), 1e-5, 1.0, 1.0, np.zeros(2), 0.0)
    assert np.isnan(out[0]).all()
    with pytest.raises(mbrf.MbrfError):
        mbrf.blochC(np.ones(4), np.zeros(4), np.ones(3), 1.0, 1.0, np.zeros(2), 0.0)   # time vector length
    with pytest.raises(mbrf.MbrfError):
        mbrf.blochC(np.ones(4), np.zeros(4), 1e-5, 1.0, 1.0, np.zeros(2), 0.0, 7)      # bad mode


def test_full_size_properties(mbrf):
    """BASELINE config 2 (512 samples x 10^6 spins): size-independent properties of the result."""
    g = golden("pulses.npz")
    b1 = g["b1_cfg2_gauss"]
    nt = b1.size
    dt = 8e-3 / nt
    gr = np.full((nt, 1), 0.05)
    df = np.linspace(-5000, 5000, 1000)
    dp = np.linspace(-5, 5, 1000).reshape(-1, 1)
    mx, my, mz = mbrf.blochC(b1.reshape(-1, 1), gr, dt, 1e3, 1e3, df.reshape(-1, 1), dp, 0)
    assert mx.shape == (1000, 1000)
    norm = np.sqrt(mx ** 2 + my ** 2 + mz ** 2)
    assert norm.max() <= 1 + 1e-9 and norm.min() > 1 - 2e-5      # T1=T2=1000 s over 8 ms: |M| shrinks < 1e-5
    # the off-resonance and the gradient act through one number, df*TWOPI + gamma*G*x (blochC.c:330):
    # spins with equal effective offset must agree.  x step 10/999 cm <-> 10/999*0.05*6726.1/6.283185 Hz
    hz_per_cm = 0.05 * 6726.1 / 6.283185
    i, j = 500, 250                                                # spin (p=i, f=j)
    eff = df[j] + dp[i, 0] * hz_per_cm
    one = mbrf.blochC(b1.reshape(-1, 1), np.zeros((nt, 1)), dt, 1e3, 1e3, np.array([eff]), 0.0, 0)
    assert abs(one[0][0, 0] - mx[i, j]) < 1e-9 and abs(one[2][0, 0] - mz[i, j]) < 1e-9
    # symmetric real pulse: Mz even in the effective offset; on resonance the flip angle is
    # sum(b1)*gamma*dt (all rotations about one axis), ~pi/2 for this 'ex' pulse
    centre = mbrf.blochC(b1.reshape(-1, 1), np.zeros((nt, 1)), dt, 1e30, 1e30, np.array([0.0, 300.0, -300.0]), 0.0, 0)
    flip = b1.real.sum() * 6726.1 * dt
    assert abs(flip - np.pi / 2) < 0.01
    assert abs(centre[2][0, 0] - np.cos(flip)) < 1e-12 and abs(centre[2][0, 1] - centre[2][0, 2]) < 1e-9


def test_full_size_sample_vs_oracle(mbrf, oracle):
    """Config 2 again: a strided sample of the 10^6 spins against the CPU oracle."""
    g = golden("pulses.npz")
    b1 = g["b1_cfg2_gauss"]
    nt = b1.size
    dt = 8e-3 / nt
    gx = np.full(nt, 0.05)
    df = np.linspace(-5000, 5000, 1000)
    dx = np.linspace(-5, 5, 1000)
    mx, my, mz = mbrf.blochC(b1.reshape(-1, 1), gx.reshape(-1, 1), dt, 1e3, 1e3, df.reshape(-1, 1),
                             dx.reshape(-1, 1), 0)
    fi = np.arange(0, 1000, 37)
    pi = np.arange(0, 1000, 41)
    want = oracle.blochsimfz_oracle(b1, gx, None, None, dt, 1e3, 1e3, df[fi], dx[pi])
    for o, w in zip((mx, my, mz), want):
        sub = o[np.ix_(pi, fi)].ravel(order="F")
        assert np.abs(sub - w).max() < TOL_BLOCH


@pytest.mark.gpu
def test_scale_sweep_matches_one_call_per_scale(mbrf, oracle):
    """The sweep extension (mbrf_bloch_scale_sweep; sim_rf_scale.m:82-89 runs one blochC / blochH call per B1 scaling): every
    (off-resonance, scale) spin of the single launch equals the oracle (bit-identical to the reference C, tests/test_oracle.py) run once per scale."""
    g = golden("pulses.npz")
    b1 = g["b1_cfg1_gauss"]
    dt = 4e-3 / b1.size
    df = np.linspace(-3000.0, 3000.0, 301)
    scales = np.array([0.5, 0.8, 1.0, 1.2, 1.5])
    for gamma, name in ((mbrf.GAMMA_C13, "C-13"), (mbrf.GAMMA_H1, "H-1")):
        mx, my, mz = mbrf.bloch_scale_sweep(b1, dt, 1e3, 1e3, df, scales, gamma)
        assert mx.shape == (scales.size, df.size)
        for k, sc in enumerate(scales):
            want = oracle.blochsimfz_oracle(b1 * sc, np.zeros(b1.size), None, None, dt, 1e3, 1e3, df, np.zeros(1), mode=0, gamma=gamma)
            assert max(np.abs(got[k] - w).max() for got, w in zip((mx, my, mz), want)) < TOL_BLOCH, (name, sc)
    r = mbrf.sim_rf_scale(b1, dt * 1e3, "pulse", None, "C-13", "ex", 2.0)       # 7-argument form: passband bandwidth in kHz
    assert r["df"].size == 2048 and r["df"][0] == -6000.0 and r["mz"].shape == (5, 2048) and np.all(r["scale"] == [0.8, 0.9, 1, 1.1, 1.2])
    mx1, my1, mz1 = mbrf.blochC(b1 * 1.1, np.zeros(b1.size), dt, 1e3, 1e3, r["df"], 0.0, 0)
    assert np.abs(r["mz"][3] - np.asarray(mz1).ravel()).max() < 1e-12 and np.abs(r["mxy"][3] - (np.asarray(mx1) + 1j * np.asarray(my1)).ravel()).max() < 1e-12
    with pytest.raises(mbrf.MbrfError):
        mbrf.bloch_scale_sweep(b1, 0.0, 1e3, 1e3, df, scales)
