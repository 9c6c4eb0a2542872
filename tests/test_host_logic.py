"""CPU tests of host-side search logic (no GPU): the speculative, concurrent bisections must take exactly the
decisions of the reference's serial loops (ss/fir_min_order_linprog.m:84-232, fir_ap.m:143-162)."""
import numpy as np
import pytest


def _serial_min_order(n, feasible):
    """Literal transcription of the control flow of ss/fir_min_order_linprog.m:84-232 (probe = feasible(n_tap))."""
    hb_o = hb_e = None
    n_odd_max, n_even_max = 2 * ((n - 1) // 2) + 1, 2 * (n // 2)
    n_bot, n_top = 1, (n_odd_max + 1) // 2
    n_cur = n_top
    while n_top - n_bot > 1:
        if feasible(2 * n_cur - 1):
            hb_o, n_top = 2 * n_cur - 1, n_cur
            n_cur = n_bot if n_top == n_bot + 1 else int(np.ceil((n_top + n_bot) / 2))
        else:
            n_bot = n_cur
            n_cur = int(np.ceil((n_bot + n_top) / 2))
    n_bot = 1
    n_top = n_even_max // 2 if hb_o is None else min(n_even_max // 2, (hb_o + 1) // 2)
    n_cur = n_top
    while n_top - n_bot > 1:
        if feasible(2 * n_cur):
            hb_e, n_top = 2 * n_cur, n_cur
            n_cur = n_bot if n_top == n_bot + 1 else int(np.ceil((n_top + n_bot) / 2))
        else:
            n_bot = n_cur
            n_cur = int(np.ceil((n_bot + n_top) / 2))
    return hb_o, hb_e


@pytest.mark.parametrize("n,min_odd,min_even", [(40, 27, 28), (40, 19, 20), (40, 5, 40), (41, 39, 39), (64, 99, 99),
                                                  (33, 3, 2), (512, 301, 288)])
def test_speculative_min_order_search_equals_serial(n, min_odd, min_even):
    from multiband_rf_pulse_design_b200 import fir

    def feasible(nt):
        return nt >= (min_odd if nt % 2 else min_even)

    def solve(nt, hw):
        return (np.zeros(nt), "Solved") if feasible(nt) else (np.zeros(0), "Failed")

    hb_o, hb_e = _serial_min_order(n, feasible)
    h, st = fir._min_order_search(n, None, None, None, 0, solve, pick_longer=False)
    if hb_o is None and hb_e is None:
        assert st == "Failed" and h.size == 0
    else:
        want = hb_e if hb_o is None else (hb_o if (hb_e is None or hb_o < hb_e) else hb_e)   # :220-228
        assert st == "Solved" and h.size == want
    h, st = fir._min_order_search(n, None, None, None, 0, solve, pick_longer=True)           # fir_min_order.m:222-226
    if not (hb_o is None and hb_e is None):
        assert h.size == (hb_o if (hb_o or 0) > (hb_e or 0) else hb_e)


def test_speculation_lists_only_reachable_probes():
    from multiband_rf_pulse_design_b200 import fir

    def step(st, solved):
        bot, top, mid = st
        bot, top = (bot, mid) if solved else (mid, top)
        return (bot, top, int(np.ceil((top + bot) / 2)) if top - bot > 1 else None)

    probes = fir._speculate((2, 256, 129), step, 3)
    assert probes[0] == 129 and set(probes) == {129, 66, 193, 34, 98, 161, 225}
    assert fir._speculate((2, 3, None), step, 3) == []


def test_seeded_sweep_solves_every_design_once_and_warm_starts_from_the_nearest_seed(monkeypatch):
    """fir._solve_seeded host logic with a fake solver: cold seeds every `stride`-th member of a (band edges, Peak) chain plus
    the largest weight, every other design solved exactly once in the second pass, started from the nearest seed's solution."""
    import numpy as np
    from multiband_rf_pulse_design_b200 import fir
    calls = []

    def fake(n, designs, warm=None, want_dual=False, **kw):
        ids = [d["id"] for d in designs]
        calls.append((ids, warm))
        B = len(designs)
        x = np.array([[d["id"]] * (2 * n - 1) for d in designs], float)
        info = np.zeros((B, 8)); info[:, 0] = 1
        if want_dual:
            return x, np.zeros(B), info, np.array([[100.0 + d["id"]] * 5 for d in designs]), np.array([2.0 + d["id"] for d in designs])
        return x, np.zeros(B), info
    monkeypatch.setattr(fir, "_solve_batch_ap", fake)
    objs = list(np.logspace(-1, 0, 40)) * 2
    peaks = [1e-3] * 40 + [2e-3] * 40
    perm = np.random.default_rng(0).permutation(80)
    designs = [dict(id=int(i)) for i in perm]
    x, t, info = fir._solve_seeded(4, designs, [b"f"] * 80, [peaks[i] for i in perm], [objs[i] for i in perm], 8)
    assert np.array_equal(x[:, 0], perm)                      # every design got its own answer back, in the caller's order
    assert len(calls) == 2 and calls[0][1] is None            # pass 1 cold
    seeds, rest = calls[0][0], calls[1][0]
    assert sorted(seeds + rest) == list(range(80)) and not set(seeds) & set(rest)
    for chain in (range(0, 40), range(40, 80)):
        s = sorted(i for i in seeds if i in chain)
        assert s == sorted(set(list(chain)[4::8]) | {chain[-1]})     # stride 8 from offset 4, plus the largest weight
    x0, y0, om = calls[1][1]
    for k, i in enumerate(rest):
        chain = [s for s in seeds if (s < 40) == (i < 40)]
        dist = min(abs(np.log(objs[s]) - np.log(objs[i])) for s in chain)
        used = int(x0[k, 0])
        assert used in chain and abs(abs(np.log(objs[used]) - np.log(objs[i])) - dist) < 1e-12      # a nearest seed (ties: either)
        assert y0[k, 0] == 100.0 + used and om[k] == 2.0 + used


def test_set_solver_options_names_and_ranges():
    # host-side knobs only: no device involved.  Values written back are the library defaults.
    from multiband_rf_pulse_design_b200 import fir
    fir.set_solver_options(eta_factor=0.9, beta_sufficient=0.2, beta_necessary=0.8, beta_artificial=0.36, omega_smoothing=0.5,
                           min_restart_interval=1, halpern=2, gemm=2, tc_digits=5)
    with pytest.raises(TypeError):
        fir.set_solver_options(beta_nec=0.9)
    with pytest.raises(ValueError):
        fir.set_solver_options(min_restart_interval=0)
    with pytest.raises(ValueError):
        fir.set_solver_options(tc_digits=9)


def test_fill_opt_param_inverts_fill_h():
    """ss/fir_linprog.m: fill_opt_param (:298-373) maps a previous filter h0 to the optimisation vector, fill_h (:274-296) maps
    the vector to the filter -- for a filter of the same length they are inverses, for a shorter one the taps that fit are copied."""
    from multiband_rf_pulse_design_b200 import fir
    rng = np.random.default_rng(0)
    for n, f in ((21, [0, 0.3, 0.5, 1]), (20, [0, 0.3, 0.5, 0.9]), (21, [-0.8, -0.5, -0.2, 0.3, 0.5, 0.9]), (20, [-0.8, -0.5, -0.2, 0.3, 0.5, 0.9])):
        nb = len(f) // 2
        p = fir.assemble_fir_linprog(n, f, [1, 1, 0, 0, 0, 0][:2 * nb], [0.1] * nb)
        x = rng.standard_normal(p["col_type"].size)
        h = fir._fill_h(x, p)
        assert h.size == n and np.array_equal(fir._fill_opt_param(h, p), x)
        assert fir._fill_opt_param(h[:-1], p) is None and fir._fill_opt_param(None, p) is None      # other parity / no h0: FFT init
        ps = fir.assemble_fir_linprog(n - 4, f, [1, 1, 0, 0, 0, 0][:2 * nb], [0.1] * nb)          # a shorter previous filter
        xs = rng.standard_normal(ps["col_type"].size)
        x0 = fir._fill_opt_param(fir._fill_h(xs, ps), p)
        assert np.abs(fir._fill_h(x0, p)[2:-2] - fir._fill_h(xs, ps)).max() == 0 and np.all(fir._fill_h(x0, p)[[0, 1, -2, -1]] == 0)


def test_sweep_batch_composition_keeps_the_instance_order(monkeypatch):
    """fir_ap_cvx_sweep composes its batches of designs with similar Peak / band edges (long batches first) and runs several
    batches from host threads; whatever the composition, results come back in ascending instance order, every instance once,
    each row belonging to its own design (the solve is stubbed out: host logic only)."""
    from multiband_rf_pulse_design_b200 import fir
    seen = []

    def fake(n, f_list, a_list, d_list, obj_list, peak_list, ipm_max_iter=None, want_h=True, oversamp=15):
        B = len(f_list)
        seen.append(sorted(set(peak_list)))
        x = np.zeros((B, 2 * n - 1))
        x[:, 0], x[:, 1], x[:, 2] = obj_list, peak_list, [f[0] for f in f_list]
        info = np.ones((B, 8))
        info[:, 7] = np.asarray(obj_list) * 2
        return x, None, info, (10, 2)
    monkeypatch.setattr(fir, "_solve_batch_ap_device", fake)
    f = np.array([-0.5, -0.2, 0.1, 0.4])
    objs, peaks, fadds = np.logspace(-2, 2, 5), np.logspace(-4, -2, 4), np.linspace(0, 0.02, 3)
    fl, ol, pl = fir.sweep_grid(f, objs, peaks, fadds)
    for world, order, conc in ((1, "grouped", 3), (1, "natural", 0), (2, "grouped", 2), (3, "grouped_ascending", 0)):
        parts = []
        for rank in range(world):
            seen.clear()
            r = fir.fir_ap_cvx_sweep(8, f, [1, 1, 0, 0], [0.1, 0.1], objs, peaks, fadds, rank=rank, world=world, batch=16,
                                     order=order, concurrent_batches=conc)
            assert np.array_equal(r["index"], np.arange(rank, len(fl), world))
            for k, i in enumerate(r["index"]):
                assert r["x"][k, 0] == ol[i] and r["x"][k, 1] == pl[i] and r["x"][k, 2] == fl[i][0] and r["ripple_stop"][k] == 2 * ol[i]
            if order == "grouped" and world == 1:
                assert all(len(p) <= 2 for p in seen)            # a batch of 16 spans at most two of the four Peak values
            parts.append(r["index"])
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(len(fl)))
