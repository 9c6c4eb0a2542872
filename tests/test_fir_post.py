"""fir_flip_zero / fir_qprog_phs: oracle restatements pinned by properties on the CPU, the GPU path
(C ABI `mbrf_flip_zero_batch`, `mbrf_fir_pdhg_solve2`) against the oracle on the same inputs."""
import warnings

import numpy as np
import pytest

from oracle import fir_post as O


def _filter(n_sb, n_pb, seed, real=False):
    """A filter with n_sb zeros on the unit circle (stop band) and n_pb zeros well off it (pass band)."""
    rng = np.random.default_rng(seed)
    if real:
        ang = np.linspace(0.45 * np.pi, 0.95 * np.pi, n_sb // 2)
        zs = np.concatenate([np.exp(1j * ang), np.exp(-1j * ang)])
        rad = rng.uniform(0.55, 0.9, n_pb // 2)
        pa = rng.uniform(0.05, 0.35, n_pb // 2) * np.pi
        zp = np.concatenate([rad * np.exp(1j * pa), rad * np.exp(-1j * pa)])
        h = np.real(O.poly_reference(np.concatenate([zs, zp])))
        return h * 0.01
    ang = np.linspace(0.3 * np.pi, 1.7 * np.pi, n_sb)
    zs = np.exp(1j * ang)
    rad = np.where(rng.random(n_pb) < 0.5, rng.uniform(0.5, 0.9, n_pb), rng.uniform(1.1, 1.6, n_pb))
    zp = rad * np.exp(1j * rng.uniform(-0.25, 0.25, n_pb) * np.pi)
    z = np.concatenate([zs, zp])
    return O.poly_reference(z[rng.permutation(z.size)]) * (0.02 - 0.01j)


def _mag(h, m=1024):
    return np.abs(np.fft.fft(h, m))


# ------------------------------------------------------------------ CPU: the restatements themselves
def test_flip_patterns_match_the_recursive_table():
    from multiband_rf_pulse_design_b200 import fir_post as P
    for n in (1, 2, 3, 7, 12):
        assert np.array_equal(P.flip_patterns(n).T, O.combination_2power(n))       # fir_flip_zero.m:119-138
    m = P.flip_patterns(15, rng=3)
    assert m.shape == (4096, 15) and np.unique(m, axis=0).shape[0] == 4096           # distinct patterns, :55-58
    code = (m.astype(np.int64) * (1 << np.arange(15)[::-1])).sum(1)
    assert np.all(np.diff(code) < 0)                                                  # sorted column indices (bits are inverted)
    assert P.flip_patterns(25, rng=1).shape == (4096, 25)


def test_flip_zero_restatement_keeps_the_magnitude_response():
    h = _filter(14, 6, seed=1)
    r = O.flip_zero_reference(h)
    assert r["idx_pb"].size == 6 and r["h_array"].shape == (21, 64)
    ref = _mag(h)
    for i in range(64):                                                               # fir_flip_zero.m:4-6
        assert np.abs(_mag(r["h_array"][:, i]) - ref).max() < 1e-9 * ref.max()
    assert np.isclose(r["h_array"].sum(0), h.sum()).all()                             # :71
    assert r["peak"][r["best"]] == r["peak"].min() <= np.abs(h).max() + 1e-12


def test_qprog_phs_assembly_agrees_with_the_restatement():
    from multiband_rf_pulse_design_b200 import fir_post as P
    f = [-0.6, -0.35, -0.15, 0.15, 0.35, 0.6]
    a = [0, 0, 1, 1, 0, 0]
    d = [0.02, 0.05 * np.exp(0.2j), 0.02]
    for n in (15, 16):
        o = O.build_fir_qprog_phs(n, f, a, d)
        p = P.assemble_fir_qprog_phs(n, f, a, d)
        ang = np.outer(p["w"], p["q"]) + p["phase"][:, None]
        K = np.hstack([np.cos(ang), np.sin(ang)])
        up = np.isfinite(p["hi"])
        A = np.vstack([K[up], -K[~up]])
        B = np.concatenate([p["hi"][up], -p["lo"][~up]])
        assert A.shape == o["A"].shape
        assert np.abs(A - o["A"]).max() < 1e-12 and np.abs(B - o["B"]).max() < 1e-15
    assert P.assemble_fir_qprog_phs(16, [-1, -0.5, 0.2, 0.6], [1, 1, 0, 0], [0.1, 0.1]) is None     # ss/fir_qprog_phs.m:190-201
    with pytest.raises(ValueError, match="sloped"):
        P.assemble_fir_qprog_phs(15, [0, 0.2, 0.4, 1], [1, 0.9, 0, 0], [0.1, 0.1])


# ------------------------------------------------------------------ GPU: flip zero
def flip_zero_check(h, Z, idx_pb, mask, got):
    """Hold GPU candidates (dict with peak / power / best / h_new [/ h_array [Num x N]]) to the accuracy of the reference's own
    loop: poly() in fp64 loses digits (intermediate coefficients grow and cancel), in MATLAB and numpy alike, so the yardstick
    is the restatement repeated in 80-bit arithmetic and the tolerance is what the fp64 restatement itself loses against it."""
    o64 = O.flip_zero_reference(h, Z=Z, mask=mask.T)
    o80 = O.flip_zero_reference(h, Z=Z, mask=mask.T, dtype=np.clongdouble)
    assert np.array_equal(idx_pb, o64["idx_pb"])
    scale = float(np.abs(o80["h_array"]).max())
    tol = 10 * float(np.abs(o64["h_array"] - o80["h_array"]).max()) + 1e-12 * scale
    if "h_array" in got:
        assert float(np.abs(got["h_array"].T - o80["h_array"]).max()) < tol             # every candidate, every tap
    assert float(np.abs(got["peak"] - o80["peak"]).max()) < tol
    assert float(np.abs(got["power"] - o80["power"]).max()) < 4 * tol * scale * np.sqrt(h.size)
    # complementary patterns have the same peak in exact arithmetic, so min() picks between them by rounding noise -- in the
    # reference too: the choice must be A minimiser (within the noise) and the taps those of the chosen pattern
    b = got["best"]
    assert float(o80["peak"][b]) <= float(o80["peak"].min()) + 2 * tol
    assert float(np.abs(got["h_new"] - o80["h_array"][:, b]).max()) < tol
    return o80


@pytest.mark.gpu
@pytest.mark.parametrize("n_sb,n_pb,seed,real", [(14, 6, 1, False), (40, 10, 2, False), (24, 8, 3, True), (200, 12, 4, False),
                                                 (9, 0, 5, False)])
def test_flip_zero_candidates_vs_oracle(mbrf, n_sb, n_pb, seed, real):
    from multiband_rf_pulse_design_b200 import fir_post as P
    h = _filter(n_sb, n_pb, seed, real)
    h_new, info = P.fir_flip_zero(h, rng=5, return_info=True)
    if n_sb < 100:                                                                    # roots() of a 212-zero filter misplaces some
        assert info["idx_pb"].size == n_pb
    res = P.flip_zero_candidates(info["Z"], info["idx_pb"], info["mask"], h.sum(), want_all=True)
    assert res["best"] == info["best"]
    flip_zero_check(h, info["Z"], info["idx_pb"], info["mask"], res)
    assert np.abs(h_new - res["h_new"]).max() < 1e-15 + (1e-9 * np.abs(h_new).max() if np.isrealobj(h_new) else 0)
    if real:
        # poly(): real coefficients iff the chosen zeros are closed under conjugation (one zero of a pair flipped alone is a
        # legitimate, complex, candidate)
        Zb = info["Z"].copy()
        sel = info["idx_pb"][info["mask"][info["best"]].astype(bool)]
        Zb[sel] = 1 / np.conj(Zb[sel])
        closed = np.allclose(np.sort_complex(Zb), np.sort_complex(np.conj(Zb)), atol=1e-12)
        assert np.isrealobj(h_new) == closed
    if n_sb < 100:
        assert np.abs(_mag(h_new) - _mag(h)).max() < 1e-6 * _mag(h).max()
        assert np.abs(h_new).max() <= np.abs(h).max() * (1 + 1e-6)


@pytest.mark.gpu
def test_flip_zero_sampled_patterns_vs_oracle(mbrf):
    """N_z > 12: 2^12 sampled patterns (fir_flip_zero.m:52-63), the same mask handed to the oracle."""
    from multiband_rf_pulse_design_b200 import fir_post as P
    h = _filter(30, 15, seed=7)
    h_new, info = P.fir_flip_zero(h, rng=11, return_info=True)
    assert info["mask"].shape == (4096, 15)
    flip_zero_check(h, info["Z"], info["idx_pb"], info["mask"], info)


@pytest.mark.gpu
def test_flip_zero_argument_errors(mbrf):
    from multiband_rf_pulse_design_b200 import fir_post as P
    with pytest.raises(mbrf.MbrfError, match="out of range or repeated"):
        P.flip_zero_candidates(np.array([0.5, 2.0]), [0, 0], np.zeros((1, 2), np.uint8), 1.0)
    with pytest.raises(ValueError):
        P.flip_zero_candidates(np.array([0.5, 2.0]), [0], np.zeros((1, 2), np.uint8), 1.0)


# ------------------------------------------------------------------ GPU: fir_qprog_phs
SPEC = dict(f=[-0.6, -0.35, -0.15, 0.15, 0.35, 0.6], a=[0, 0, 1, 1, 0, 0], d=[0.02, 0.05 * np.exp(0.2j), 0.02])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [21, 22])
def test_fir_qprog_phs_vs_cpu_qp(mbrf, n):
    """Same QP (restated from ss/fir_qprog_phs.m) solved by SciPy SLSQP on the CPU: strictly convex, so the minimiser itself
    is compared, not only the objective."""
    from multiband_rf_pulse_design_b200 import fir_post as P
    h, st, ex = P.fir_qprog_phs(n, SPEC["f"], SPEC["a"], SPEC["d"], return_info=True)
    assert st == "Solved" and h.shape == (n,)
    o = O.build_fir_qprog_phs(n, SPEC["f"], SPEC["a"], SPEC["d"])
    res = O.solve_fir_qprog_phs_reference(o)
    assert res.success, res.message
    x = ex["x"]
    assert (o["A"] @ x - o["B"]).max() < 1e-6                                         # violation, recomputed on the CPU
    e_gpu, e_cpu = np.linalg.norm(x), np.linalg.norm(res.x)
    assert abs(e_gpu - e_cpu) < 1e-4 * e_cpu                                          # objective (north star: 1e-4)
    assert np.abs(x - res.x).max() < 2e-3 * np.abs(res.x).max()


@pytest.mark.gpu
def test_fir_qprog_phs_infeasible_and_rejected(mbrf):
    from multiband_rf_pulse_design_b200 import fir_post as P
    h, st = P.fir_qprog_phs(7, SPEC["f"], SPEC["a"], SPEC["d"])                       # far too short: certificate
    assert st == "Failed" and h.size == 0
    h, st = P.fir_qprog_phs(16, [-1, -0.5, 0.2, 0.6], [1, 1, 0, 0], [0.1, 0.1])      # even n, non-zero at fs/2
    assert st == "Failed" and h.size == 0


@pytest.mark.gpu
def test_fir_min_order_qprog_phs(mbrf):
    """The bisection of ss/fir_min_order_qprog_phs.m against the same search driven by the CPU QP."""
    from multiband_rf_pulse_design_b200 import fir_post as P
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        h, st = P.fir_min_order_qprog_phs(24, SPEC["f"], SPEC["a"], SPEC["d"], even_odd=1)
    assert st == "Solved"

    def cpu_feasible(n):
        o = O.build_fir_qprog_phs(n, SPEC["f"], SPEC["a"], SPEC["d"])
        r = O.solve_fir_qprog_phs_reference(o)
        return bool(r.success and (o["A"] @ r.x - o["B"]).max() < 1e-7)
    n = len(h)
    assert n % 2 == 1 and cpu_feasible(n) and not cpu_feasible(n - 2)
