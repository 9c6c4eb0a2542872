"""The whole design -> verify chain of dzrf_mb.m / specsat_H1_dualband.m on the GPU, stage by stage against the oracle chain:

    fir_ap_cvx (dzrf_mb.m:163-214)  ->  b = b(end:-1:1) (:220)  ->  a = b2a(b); rf = ab2rf(a, b) (:239-240)
        ->  rfscaleg (:244)  ->  blochH over the off-resonance grid (sim_rf_spectral.m:64-78)  ->  |Mz| against the band spec

Every stage after the solve is compared with the CPU restatement fed the SAME input; the end result is compared with the
specification the design was solved for (the physics check the reference does by eye, sim_rf_spectral.m:96-98)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("obj", [1.0, 1e4])
def test_design_to_bloch_chain(mbrf, oracle, obj):
    from oracle.fir_problems import H1_DUALBAND as S, x_to_h_reference
    n, dt_ms, gamma = 256, 0.1, 4.2576                                   # specsat_H1_dualband.m: dt = 0.1 ms, H-1
    hs, st, ex = mbrf.fir_ap_cvx_batch(n, [S["f"]], S["a"], S["d"], [obj], [1e-2], return_info=True)
    assert st == ["Solved"]
    h = hs[0]
    # stage 1: the taps are the minimum-phase spectral factor of the solved autocorrelation (fir_ap_cvx.m:185-202)
    assert np.abs(h - x_to_h_reference(ex["x"][0], n)).max() < 1e-10
    b = h[::-1]                                                          # dzrf_mb.m:220
    # stage 2: inverse SLR
    a = mbrf.b2a(b)
    assert np.abs(a - oracle.b2a_m(b)).max() < 1e-11
    rf = mbrf.ab2rf(a, b)
    assert np.abs(rf - oracle.ab2rf_m(oracle.b2a_m(b), b)).max() < 1e-9
    # stage 3: Gauss, Bloch simulation over the off-resonance grid (sim_rf_spectral.m:64-78: G = 0, T1 = T2 = 1e3 s, mode 0)
    rfg = oracle.rfscaleg(rf, dt_ms * n, gamma)                          # rfscaleg.m:10-12 (host arithmetic)
    fs_hz = 1e3 / dt_ms
    df = np.linspace(-0.06, 0.06, 1537) * fs_hz / 2
    mx, my, mz = mbrf.blochH(rfg, np.zeros(n), dt_ms * 1e-3, 1e3, 1e3, df, 0.0, 0)
    want = oracle.blochsimfz_oracle(rfg, np.zeros(n), None, None, dt_ms * 1e-3, 1e3, 1e3, df, np.zeros(1), mode=0, gamma=oracle.GAMMA_H1)
    assert max(np.abs(g.ravel(order="F") - w).max() for g, w in zip((mx, my, mz), want)) < 1e-9
    # stage 4: the simulated profile meets the band specification the design was solved for: |beta| = sin(theta/2),
    # Mz = 1 - 2 |beta|^2 (abr.m:10-13), to the accuracy of the SLR hard-pulse approximation
    mz = np.asarray(mz).ravel()
    fn = df / (fs_hz / 2)
    f, amp, d = np.array(S["f"]), np.array(S["a"]), np.array(S["d"])
    for k in range(3):
        sel = (fn >= f[2 * k] + 2e-3) & (fn <= f[2 * k + 1] - 2e-3)
        assert sel.sum() > 20
        beta = np.sqrt(np.clip((1 - mz[sel]) / 2, 0, 1))
        assert np.all(beta <= amp[2 * k] + d[k] + 2e-3) and np.all(beta >= amp[2 * k] - d[k] - 2e-3), (k, beta.min(), beta.max())
