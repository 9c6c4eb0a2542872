"""The drop-in host call spread over several GPUs inside ONE call (mbrf_set_fanout), and pageable result arrays."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fanout_and_pageable_results_match(mbrf, oracle):
    lib = mbrf.lib()
    ndev = lib.mbrf_device_count()
    rng = np.random.default_rng(3)
    nt = 96
    b1 = rng.normal(0, 0.05, nt) + 1j * rng.normal(0, 0.05, nt)
    gr = rng.normal(0, 0.2, (nt, 1))
    df = np.linspace(-2000, 2000, 401)
    dp = np.linspace(-3, 3, 523).reshape(-1, 1)                     # 209 723 spins: more than one device's minimum share
    try:
        assert lib.mbrf_set_fanout(1) == 0 and lib.mbrf_get_fanout() == 1
        one = mbrf.blochC(b1, gr, 1e-5, 0.3, 0.04, df, dp, 0)
        assert lib.mbrf_set_fanout(0) == 0 and lib.mbrf_get_fanout() == max(1, ndev)      # 0 = every device of the box
        many = mbrf.blochC(b1, gr, 1e-5, 0.3, 0.04, df, dp, 0)
        for a, b in zip(one, many):
            assert np.array_equal(a, b)                              # same kernel, same arithmetic, whatever the device
        want = oracle.blochsimfz_oracle(b1, gr[:, 0], None, None, 1e-5, 0.3, 0.04, df[:3], dp[:, 0], mode=0)
        got = mbrf.blochC(b1, gr, 1e-5, 0.3, 0.04, df[:3], dp, 0)
        assert max(np.abs(g.ravel(order="F") - w).max() for g, w in zip(got, want)) < 1e-9
        # mode 2 with an initial magnetisation through the chunked pipeline
        m0 = [rng.normal(0, 0.3, (523, 401)) for _ in range(3)]
        a2 = mbrf.blochC(b1, gr, 1e-5, 0.3, 0.04, df, dp, 2, *m0)
        lib.mbrf_set_fanout(1)
        b2 = mbrf.blochC(b1, gr, 1e-5, 0.3, 0.04, df, dp, 2, *m0)
        assert all(np.array_equal(x, y) for x, y in zip(a2, b2)) and a2[0].shape == (nt, 523, 401)
        lib.mbrf_set_fanout(0)
        x = np.linspace(-8, 8, 300001)
        rf = rng.normal(0, 0.03, 150) + 1j * rng.normal(0, 0.03, 150)
        al, be = mbrf.abrx(rf, np.full(150, 0.1), x)
        lib.mbrf_set_fanout(1)
        al1, be1 = mbrf.abrx(rf, np.full(150, 0.1), x)
        assert np.array_equal(al, al1) and np.array_equal(be, be1)
    finally:
        lib.mbrf_set_fanout(1)
    assert lib.mbrf_set_fanout(-1) != 0
