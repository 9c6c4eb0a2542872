"""The tcgen05 split-integer product (csrc/tc_gemm.cuh) against fp64: the batched A*X of the PDHG iterations
(fir_ap_cvx.m:160-169 / fir_linprog.m:246-252 re-expressed) must be an fp64-grade product, and the
solver must take the same path with either product kernel."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _product(mbrf, A, X, nd, nslab):
    import torch
    lib = mbrf.lib()
    R, kdim = A.shape
    Bp = X.shape[1]
    dA = torch.from_numpy(A).cuda()
    dX = torch.from_numpy(X).cuda()
    dC = torch.empty((nslab, R, Bp), dtype=torch.float64, device="cuda")
    ms = (C.c_float * 2)()
    rc = lib.mbrf_tc_product_device(dA.data_ptr(), R, kdim, dX.data_ptr(), Bp, nd, nslab, dC.data_ptr(), 2,
                                    C.cast(C.byref(ms, 0), C.c_void_p), C.cast(C.byref(ms, 4), C.c_void_p), None)
    assert rc == 0, mbrf.lib().mbrf_last_error().decode()
    torch.cuda.synchronize()
    return dC.sum(0).cpu().numpy(), ms[0], ms[1]


@pytest.mark.parametrize("shape", [(512, 256, 128, 1), (512, 1216, 192, 1), (1984, 512, 64, 5), (7872, 512, 512, 9)])
@pytest.mark.parametrize("nd", [5, 6])
def test_tc_product_matches_fp64(mbrf, shape, nd):
    kdim, R, Bp, nslab = shape
    rng = np.random.default_rng(7)
    # Fourier-like rows of very different magnitude, iterates with exact zeros, a dead design and a wide dynamic range
    A = np.cos(np.outer(np.arange(1, R + 1) * 1e-3, np.arange(1, kdim + 1))) * (2.0 ** ((np.arange(R) % 9) - 4))[:, None]
    X = rng.standard_normal((kdim, Bp)) * (10.0 ** rng.uniform(-6, 2, size=(1, Bp)))
    X[rng.random((kdim, Bp)) < 0.3] = 0.0
    X[:, 3] = 0.0
    X[::17, :] *= 1e-7
    Cg, _, _ = _product(mbrf, np.ascontiguousarray(A), np.ascontiguousarray(X), nd, nslab)
    ref = (A.astype(np.longdouble) @ X.astype(np.longdouble)).astype(np.float64)
    scale = np.abs(A).max(1)[:, None] * np.abs(X).max(0)[None, :] * np.sqrt(kdim)
    tol = {5: 1e-10, 6: 1e-12}[nd]          # measured 1e-11 / 5e-14 of |row|max * |column|max * sqrt(k)
    err = np.abs(Cg - ref)
    assert np.all(err <= tol * scale + 1e-300), float((err / np.maximum(scale, 1e-300)).max())
    assert np.all(Cg[:, 3] == 0.0)


def test_tc_product_rejects_bad_shapes(mbrf):
    import torch
    lib = mbrf.lib()
    d = torch.zeros(64 * 64, dtype=torch.float64, device="cuda")
    for R, kdim, Bp, nd, ns in [(63, 64, 64, 5, 1), (64, 64, 32, 5, 1), (64, 64, 64, 3, 1), (64, 64, 64, 7, 1), (64, 64, 64, 5, 0)]:
        assert lib.mbrf_tc_product_device(d.data_ptr(), R, kdim, d.data_ptr(), Bp, nd, ns, d.data_ptr(), 1, None, None, None) != 0


def test_solver_same_answer_on_both_product_kernels(mbrf):
    """64 designs of the dual-band spec at n = 64: fp64 DMMA tiles vs tcgen05 int8 tiles give the same statuses and
    objectives (1e-6 relative: the two runs differ only in product rounding, 1e-11 vs 1e-16)."""
    from multiband_rf_pulse_design_b200 import fir
    lib = mbrf.lib()
    f = [-0.6, -0.35, -0.2, 0.18, 0.38, 0.6]
    a = [0.866, 0.866, 0, 0, 0.707, 0.707]
    d = [0.02, 0.03, 0.025]
    objs = np.logspace(-2, 0.5, 16)
    peaks = np.logspace(-2.2, -1.2, 4)
    out = {}
    try:
        for mode in (1, 2):
            assert lib.mbrf_pdhg_set_gemm(mode) == 0
            out[mode] = fir.fir_ap_cvx_sweep(48, f, a, d, objs, peaks, [0.0], batch=64, max_iter=40000, method="pdhg")["info"]
    finally:
        lib.mbrf_pdhg_set_gemm(2)
    assert np.array_equal(out[1][:, 0], out[2][:, 0])
    ok = out[1][:, 0] == 1
    assert ok.sum() >= 32
    rel = np.abs(out[1][ok, 2] - out[2][ok, 2]) / np.abs(out[1][ok, 2])
    assert rel.max() < 1e-6, rel.max()
    assert out[2][ok, 4].max() <= 1e-6


def test_seeded_sweep_matches_cold_sweep(mbrf):
    """fir_ap_cvx_sweep(seed_stride=...) (cold seeds + one warm-started hop, mbrf_fir_pdhg_warm_start) returns the same
    optima as the cold sweep: both stop at a 5e-5 relative gap, so objectives agree within 1e-4, and every design that the
    cold sweep solves is solved."""
    from multiband_rf_pulse_design_b200 import fir
    f = [-0.6, -0.35, -0.2, 0.18, 0.38, 0.6]
    a = [0.866, 0.866, 0, 0, 0.707, 0.707]
    d = [0.02, 0.03, 0.025]
    objs = np.logspace(-1, 0, 96)              # 0.0105 decades apart -> "auto" picks stride 9
    cold = fir.fir_ap_cvx_sweep(48, f, a, d, objs, [10 ** -1.5], [0.0], batch=96, max_iter=40000, method="pdhg")
    warm = fir.fir_ap_cvx_sweep(48, f, a, d, objs, [10 ** -1.5], [0.0], batch=96, max_iter=40000, seed_stride="auto", method="pdhg")
    ok = cold["info"][:, 0] == 1
    assert ok.sum() >= 90
    assert np.all(warm["info"][ok, 0] == 1)
    rel = np.abs(warm["info"][ok, 2] - cold["info"][ok, 2]) / np.abs(cold["info"][ok, 2])
    assert rel.max() < 1e-4, rel.max()
    assert warm["info"][ok, 4].max() <= 1e-6
    assert warm["info"][ok, 1].sum() < cold["info"][ok, 1].sum()      # fewer iterations in total
    assert np.abs(warm["x"][ok] - cold["x"][ok]).max() < 5e-3         # same design, up to the tolerance of a first-order solve


def test_halpern_option_solves_to_the_same_tolerances(mbrf):
    """mbrf_pdhg_set_halpern: the reflected Halpern iteration (default, mode 2) and the averaged restarts (mode 0) reach the
    same optima (objective within 1e-4) on a small trade-off grid."""
    from multiband_rf_pulse_design_b200 import fir
    lib = mbrf.lib()
    f = [-0.6, -0.35, -0.2, 0.18, 0.38, 0.6]
    a = [0.866, 0.866, 0, 0, 0.707, 0.707]
    d = [0.02, 0.03, 0.025]
    objs = np.logspace(-2, 0, 8)
    out = {}
    try:
        for mode in (0, 1, 2):
            assert lib.mbrf_pdhg_set_halpern(mode) == 0
            out[mode] = fir.fir_ap_cvx_sweep(48, f, a, d, objs, [10 ** -1.5], [0.0], batch=8, max_iter=40000, method="pdhg")["info"]
    finally:
        lib.mbrf_pdhg_set_halpern(2)
    for mode in (1, 2):
        assert np.all(out[0][:, 0] == 1) and np.all(out[mode][:, 0] == 1)
        rel = np.abs(out[0][:, 2] - out[mode][:, 2]) / np.abs(out[0][:, 2])
        assert rel.max() < 1e-4, (mode, rel.max())
        assert out[mode][:, 4].max() <= 1e-6
