#!/usr/bin/env python
"""bench.py — Bloch hot path on B200 (BASELINE.json metric "Bloch spin-steps/sec"), with the forward-SLR and the
convex-FIR-design paths reported in the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (`config.workload`): BASELINE configs[1] — a 512-sample excitation pulse simulated
over 10^6 spins per GPU (1000 off-resonances x 1000 positions with a 0.05 G/cm x-gradient,
T1 = T2 = 1000 s, mode 0, 13C gamma), i.e. 5.12e8 spin-steps per step per GPU.  At N GPUs the
job is 1000*N off-resonances x 1000 positions, sharded by contiguous spin range
(s = p + npos*f, blochC.c:468-473); the final gather is fused into the kernels: every rank stores
its slice straight into rank 0's peer-mapped result over NVLink.

One JSON line is printed by rank 0:
  value        device-resident throughput (inputs already in HBM, one mbrf_bloch_device per rank writing
               into rank 0's result), timed with CUDA events on the launching stream, max over ranks
  e2e          the same metric through ONE reference-facing call  blochC(b1,gr,tp,t1,t2,df,dp,mode)
               (C ABI mbrf_bloch) on rank 0 for the whole job: HOST buffers, the result arrays freshly
               allocated PAGEABLE memory as a MEX gateway gets them, all N GPUs used inside the call
               (mbrf_set_fanout); H2D, D2H and the host-side copies are inside the timed region
  solver_*     short copies of the second hot path's numbers: BASELINE config 4 (4096 fir_ap_cvx designs, the same
               grid at every N), config 3 (fir_qp_cvx, one design), config 5 (C-13 order search); details under "solver"
  roofline     FP64-pipe roofline of the dominant kernel (76 algorithmic flops per
               spin-step; the path is FP64-bound, 56 B of HBM traffic per spin) against the FP64
               FMA peak measured in this run by a DFMA-only kernel; `hbm` shows why it is not HBM-bound
  cpu_baseline the reference's own C (oracle/_ref/libblochC.so: blochC.c compiled unmodified)
               on this box's host cores over a bounded sample of the same workload

`--impl reference` times that CPU implementation alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_REAL_STDOUT = sys.stdout
METRIC = "Bloch spin-steps/sec"
UNIT = "spin-steps/s"
NTIME = 512
NF_PER_GPU = 1000
NPOS = 1000
FLOPS_PER_SPIN_STEP = 76        # blochC.c:330-361 as written, sqrt/div/sin/cos at 1 each
FP64_INST_PER_SPIN_STEP = 35    # what our kernel issues on the FP64 pipe for this workload (cuobjdump -sass:
                                # constant dt and gradient -> constant-rz loop, |phi| <= 1 rad -> tier TINY)
BYTES_PER_SPIN = 56             # df (8) + M0/positions amortised (24) + M out (24): SURVEY.md 8(d)
WORKLOAD = ("bloch cfg2: 512-sample dzrf 'ex' pulse x 1e6 spins per GPU "
            "(1000 df x 1000 dp, Gx=0.05 G/cm, T1=T2=1e3 s, mode 0, blochC gamma)")


# ----------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------
def make_pulse():
    """512-sample b1 in Gauss: the committed fixture made with the reference's own dzrf recipe
    (tests/golden/make_golden.py); an inline Hamming-sinc stand-in if the fixture is absent."""
    p = os.path.join(ROOT, "tests", "golden", "pulses.npz")
    if os.path.exists(p):
        return np.load(p)["b1_cfg2_gauss"].astype(np.complex128), "tests/golden/pulses.npz:b1_cfg2_gauss"
    n, m = NTIME, 2
    x = np.arange(-n / 2, n / 2) / (n / 2)
    snc = np.sin(m * 2 * np.pi * x + 1e-5) / (m * 2 * np.pi * x + 1e-5)
    rf = snc * (0.54 + 0.46 * np.cos(np.pi * x))
    rf = rf / rf.sum() * (np.pi / 2)
    return (rf / (2 * np.pi * 1.0705 * (8.0 / n))).astype(np.complex128), "inline msinc stand-in"


def make_workload(world):
    b1, src = make_pulse()
    nt = b1.size
    return dict(b1=b1, gx=np.full(nt, 0.05), dt=8e-3 / nt, t1=1e3, t2=1e3,
                df=np.linspace(-5000.0, 5000.0, NF_PER_GPU * world), dx=np.linspace(-5.0, 5.0, NPOS), src=src)


# ----------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [l.split(",") for (t, l) in self.lines if t0 <= t <= t1 + 0.03] or [l.split(",") for _, l in self.lines]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------
# CPU arm: the reference's own C on host cores
# ----------------------------------------------------------------------------
def cpu_reference_run(wl, nf_per_thread, threads):
    """Time the reference blochsimfz over threads x nf_per_thread offsets x NPOS positions.
    Returns (spin_steps_per_s, seconds, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref
    ref.build()
    use_ref = ref.have_ref()
    nt = wl["b1"].size
    df_all = wl["df"]

    def work(i):
        df = np.ascontiguousarray(df_all[(i * nf_per_thread) % max(1, df_all.size - nf_per_thread):][:nf_per_thread])
        if use_ref:
            lib = ref._lib(os.path.join(ref.REF_DIR, "libblochC.so"))
            fn = lib.blochsimfz
            fn.argtypes = ref._BLOCHSIMFZ_ARGS
        else:
            lib = ref._lib(os.path.join(ref.HERE, "liboracle.so"))
            fn = lib.oracle_blochsimfz
            fn.argtypes = ref._BLOCHSIMFZ_ARGS + [C.c_double]
        fn.restype = C.c_int
        b1r, b1i, gx, gy, gz, dt, n, dfp, dx, dy, dz, ntout, out = ref._prep_sim(
            wl["b1"], wl["gx"], None, None, wl["dt"], df, wl["dx"], None, None, 0, None)
        a = [ref._ptr(b1r), ref._ptr(b1i), ref._ptr(gx), ref._ptr(gy), ref._ptr(gz), ref._ptr(dt), n, wl["t1"],
             wl["t2"], ref._ptr(dfp), dfp.size, ref._ptr(dx), ref._ptr(dy), ref._ptr(dz), dx.size,
             ref._ptr(out[0]), ref._ptr(out[1]), ref._ptr(out[2]), 0]
        if not use_ref:
            a.append(ref.GAMMA_C13)
        fn(*a)                       # ctypes releases the GIL: threads run the C loop in parallel
        return float(out[2][0])

    with ref._quiet_stdout():        # the reference prints "%d%% Complete." above 40 000 spins (blochC.c:502)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, range(threads)))
        sec = time.perf_counter() - t0
    steps = threads * nf_per_thread * NPOS * nt
    return steps / sec, sec, ("reference" if use_ref else "port")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    wl = make_workload(1)
    nf_pt = 20                      # 20 offsets x 1000 positions x 512 samples ~ 0.65 s per thread per step
    for _ in range(args.warmup):
        cpu_reference_run(wl, nf_pt, threads)
    tot_steps, tot_sec, kind = 0.0, 0.0, "reference"
    for _ in range(args.steps):
        v, sec, kind = cpu_reference_run(wl, nf_pt, threads)
        tot_steps += v * sec
        tot_sec += sec
    value = tot_steps / tot_sec
    sample = f"{threads} threads x {nf_pt} df x {NPOS} dp x {NTIME} samples per step (bounded sample of the workload)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_sec / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pulse": wl["src"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    return 0


# ----------------------------------------------------------------------------
# solver legs (second hot path: "N=256 FIR pulse designs solved/sec"), reported in the same JSON line
# ----------------------------------------------------------------------------
H1_DUALBAND = dict(f=[-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006],       # SURVEY.md 8(d):
                   a=[0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886],                     # specsat_H1_dualband.m:5-32
                   d=[0.014436, 0.022361, 0.017683])                                         # after dzrf_mb's shift_f


def c13_bssfp_spec(B0=14.0, n=200, T=4.0, FA=60.0, d1=0.01, d2=0.005):
    """The band specification of BASELINE config 5 from the reference's input script bSSFP_pulse_lp_ap.m:8-64 (urea selected, 5 bands
    of 0.1 kHz, 'ex'), built with the package's mirrors of the specification builders (spec.py: spectrum_C13.m, rf_bandedge.m,
    rf_ripple_GFA.m, dzrf_mb.m:92-121).  Returns f (normalised to [-1,1]), a, d, dt (ms)."""
    from multiband_rf_pulse_design_b200 import spec
    fhz, _ = spec.spectrum_C13(B0)                                         # bSSFP_pulse_lp_ap.m:25
    pick = np.array([6, 1, 3, 4, 2]) - 1                                   # :29: urea pyr ala pyr-H2O lac
    cf = fhz[pick]
    cf = (cf - cf[0]) * 1e-3                                               # kHz, selected compound on resonance (:55, :64)
    out = spec.multiband_spec(n, T / n, list(cf), [0.1] * 5, [FA, 0, 0, 0, 0], [d1, d2, d2, d2, d2], "ex")   # :30, :52-55, :68-74
    return out["f"], out["a"], out["d"], out["dt"]


def leg_cfg4(args, m, lib, rank, world, dev, max_over_ranks, barrier):
    """BASELINE config 4: the dual-band H-1 saturation trade-off sweep -- 16 obj in logspace(-2,4) x 16 Peak in logspace(-4,-2)
    x 16 band-edge expansions f_add in linspace(0, 0.9 df_min/2) (fir_ap.m:70-83) = 4096 fir_ap_cvx designs at N = 256, the SAME
    grid at every GPU count (strong scaling): instance i goes to rank i mod world, every rank solves its share in batches of 512
    on the interior-point solver, the results (x, ripple_stop, status) are gathered on rank 0.  One timed pass through the
    public API: every batch is ONE C call (mbrf_fir_ap_solve: specification in, per-design assembly by kernels, solve, D2H of
    the solutions) plus the gather, all inside the timed region."""
    from multiband_rf_pulse_design_b200 import fir
    from multiband_rf_pulse_design_b200.shard import gather_sweep
    n = 256
    no, npk, nfa = args.solver_grid
    f = np.array(H1_DUALBAND["f"])
    df_min = float((f[2:-1:2] - f[1:-2:2]).min())
    objs, peaks, fadds = np.logspace(-2, 4, no), np.logspace(-4, -2, npk), np.linspace(0.0, 0.9 * df_min / 2, nfa)
    total = no * npk * nfa
    fir.fir_ap_cvx_sweep(n, f, H1_DUALBAND["a"], H1_DUALBAND["d"], objs[:2], peaks[-2:], fadds[:2], batch=8, method="ipm")   # warm-up
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.mbrf_launch_count()
    t0 = time.perf_counter()
    # three batches in flight per GPU (three host threads, three streams): the per-design kernels of a batch leave SMs idle once
    # most of its designs have finished, the other batches fill them; batches are composed of designs with similar Peak / band
    # edges and the long ones go first (fir_ap_cvx_sweep order="grouped").  Measured on the full grid, one B200: natural order,
    # 2 in flight 318 designs/s; grouped 345; grouped, 3 in flight 377; 4 x 256 is slower (285).
    conc = 3
    local = -(-total // world)
    batch = 512 if local > 512 * conc else max(64, -(-local // conc // 64) * 64)
    r = fir.fir_ap_cvx_sweep(n, f, H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, fadds, rank=rank, world=world, batch=batch,
                             method="ipm", concurrent_batches=conc)
    t_solve = time.perf_counter() - t0
    full = gather_sweep(r, n, total, device=dev if world > 1 else None)
    t1 = time.perf_counter()
    sec = max_over_ranks(t1 - t0)
    sec_solve = max_over_ranks(t_solve)
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = lib.mbrf_launch_count() - l0
    barrier()
    if rank != 0:
        return None
    info = full["info"]
    st = info[:, 0]
    ok = st == 1
    out = {"metric": "N=256 FIR pulse designs solved/sec", "value": total / sec, "unit": "designs/s", "designs": total,
           "seconds": sec, "seconds_solve_only": sec_solve, "scaling": "strong",
           "workload": "cfg4: fir_ap_cvx, dual-band H-1 sat spec, N=256, 7686-row grid, %d obj in logspace(-2,4) x %d Peak in "
                       "logspace(-4,-2) x %d f_add in linspace(0, 0.9 df_min/2); same grid at every GPU count" % (no, npk, nfa),
           "method": "interior point (mbrf_fir_ap_solve): problem assembly on the device (band masks, bounds, stop rows, radii by "
                     "kernels; only the union grid is built on the host), structured normal matrix from Toeplitz/Hankel moments, "
                     "batched double-double Cholesky", "assembly": fir.DEFAULT_ASSEMBLE, "batch": batch, "concurrent_batches": conc, "batch_order": "grouped by (Peak descending, band edges, weight)",
           "status_counts": {"solved": int(ok.sum()), "infeasible_certificate": int((st == 2).sum()),
                             "iteration_limit": int((st == 3).sum())},
           "iterations_mean": float(info[:, 1].mean()), "iterations_max": float(info[:, 1].max()),
           "max_violation_solved": float(info[ok, 4].max()) if ok.any() else None,
           "max_rel_gap_solved": float((np.abs(info[ok, 2] - info[ok, 3]) / np.maximum(np.abs(info[ok, 2]), 1e-300)).max()) if ok.any() else None,
           "gpu_launches_rank0": int(launches), "clocks": clocks,
           "tolerances": {"feastol": fir.IPM_FEASTOL, "reltol": fir.IPM_RELTOL},
           "note": "most of the grid is infeasible at N = 256 (Peak < ~1e-3, or band edges widened by more than ~0.15 df_min/2): those "
                   "designs end with a Farkas certificate ('Failed', like CVX's Infeasible); all 4096 are counted in designs/s"}
    # the step after the solve (fir_ap_cvx.m:185-202): x -> minimum-phase taps h = fmp2(r), batched on the GPU
    if ok.any():
        R = np.stack([fir._x_to_r(x, n) for x in full["x"][ok]])
        fir.fmp2_batch(R[:8])
        t2 = time.perf_counter()
        fir.fmp2_batch(R)
        dt = time.perf_counter() - t2
        out["fmp2"] = {"designs": int(ok.sum()), "seconds": dt, "designs_per_s": float(ok.sum() / dt)}
    return out


def solver_roofline(lib, tf_peak):
    """Roofline of the solver's dominant kernel, ipm::cholesky_kernel<dd> (42 % of a sweep, profiles/r2_ipm_launches_summary.txt):
    the batched double-double Cholesky of the 512 x 512 normal matrices of a 512-design batch, timed alone with CUDA events
    inside libmbrf.  FP64-pipe bound: algorithmic work = n^3/6 double-double multiply-subtracts per matrix, 15 FP64-pipe
    instructions each (dd.cuh dd_fnma: 1 DMUL, 3 DFMA, 11 DADD); peak = the DFMA issue rate measured in this run."""
    from multiband_rf_pulse_design_b200._lib import check
    nv, B = 512, 512
    ms = C.c_float()
    check(lib.mbrf_ipm_cholesky_bench(nv, B, 1, 5, C.byref(ms)))
    ms64 = C.c_float()
    check(lib.mbrf_ipm_cholesky_bench(nv, B, 0, 5, C.byref(ms64)))
    macs = nv ** 3 / 6.0 * B
    ach = 2.0 * 15.0 * macs / (ms.value * 1e-3) / 1e12          # FP64-pipe instructions counted as 2 flops each, like a DFMA
    return {"bound": "fp64", "kernel": "mbrf::ipm::cholesky_kernel<dd>", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s (FP64-pipe issue, 2 per instruction)",
            "frac": ach / tf_peak, "traffic": None, "kernel_ms": ms.value, "shape": {"order": nv, "matrices": B},
            "algorithmic": "n^3/6 double-double multiply-subtracts per matrix x 15 FP64 instructions",
            "dd_gflops_equiv": 2.0 * macs / (ms.value * 1e-3) / 1e9, "fp64_variant_ms": ms64.value,
            "peak_source": "FP64 FMA peak measured in this run (mbrf_measure_fp64_peak)"}


def leg_cfg3(m, lib, hbm_peak):
    """BASELINE config 3: fir_qp_cvx(256, f, a, d, 120, 1e6) (dzrf_mb.m:211-213), one design, on the reference grid
    (oversamp 10, 2566 points) and on the 4096-point grid (oversamp 16).  Single designs stay on the first-order solver:
    vectorised matrix-vector passes over K and K^T (no interior-point path for the disk rows of fir_qp_cvx).
    HBM roofline per SURVEY.md 8(d): 2 * sizeof(K) algorithmic bytes per iteration."""
    from multiband_rf_pulse_design_b200 import fir
    out = {}
    for tag, os_ in (("grid2566", 10), ("grid4096", 16)):
        t1 = time.perf_counter()
        _, st3, ex3 = fir.fir_qp_cvx(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 120, 1e6, return_info=True,
                                     oversamp=os_, max_iter=1500000)
        sec = time.perf_counter() - t1
        mrows = 2 * ex3["problem"]["w"].size + 2 * 256
        kbytes = 8.0 * mrows * 512
        it = float(ex3["info"][1])
        gbs = 2.0 * kbytes * it / sec / 1e9
        out[tag] = {"status": st3, "seconds": sec, "iterations": it, "objective": float(ex3["info"][2]), "dual_bound": float(ex3["info"][3]),
                    "max_violation": float(ex3["info"][4]), "rows": int(mrows), "us_per_iteration": sec / max(it, 1) * 1e6,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                 "algorithmic_bytes_per_iteration": 2.0 * kbytes,
                                 "note": "whole call (assembly + solve) / iterations; K and K^T (%.0f MB together) stay in the 126 MB L2" % (2 * kbytes / 1e6)}}
    out["workload"] = "cfg3: fir_qp_cvx min-energy/min-peak multiband FIR, N=256, dual-band H-1 spec, k=120, obj=1e6, single design"
    return out


def leg_cfg5(m, lib, rank, world, allgather_obj, barrier):
    """BASELINE config 5: C-13 bSSFP duration search at 14 T -- the minimal order of the arbitrary-phase design
    fir_ap(512, f, a, d, Peak=1e-3, min_order=1) (fir_ap.m:143-166, reached from dzrf_mb 'ap_minorder_cvx') and of the
    linear-phase design fir_min_order_linprog(512, ...) (ss/fir_min_order_linprog.m), on the spec of bSSFP_pulse_lp_ap.m.
    The arbitrary-phase search probes the next three levels of the reference's bisection tree at once (different orders =
    different matrices); the probes of a round are spread over the GPUs, the tree is walked with the reference's decisions."""
    from multiband_rf_pulse_design_b200 import fir
    f, a, d, dt = c13_bssfp_spec()
    lam, peak, n_top = 0.1, 1e-3, 512

    def step(st, solved):                                                  # fir_ap.m:143-162
        bot, top, mid = st
        if solved:
            top = mid
        else:
            bot = mid
        return (bot, top, int(np.ceil((top + bot) / 2)) if top - bot > 1 else None)

    barrier()
    t0 = time.perf_counter()
    cache, probes, rounds = {}, 0, 0
    state = (2, n_top, int(np.ceil((n_top + 2) / 2)))
    first = True
    while state[2] is not None or first:
        need = ([n_top] if first else []) + [q for q in fir._speculate(state, step, 3) if q not in cache]
        first = False
        mine = need[rank::world]
        res = fir._solve_concurrently([(lambda q=q: fir.fir_ap_cvx(q, f, a, d, lam, peak, method="ipm")[1]) for q in mine])
        merged = {}
        for part in allgather_obj(dict(zip(mine, res))):
            merged.update(part)
        cache.update(merged)
        probes += len(need)
        rounds += 1
        if cache.get(n_top) == "Failed":
            break                                                          # "original parameters are too tight", fir_ap.m:52-54
        for _ in range(3):
            if state[2] is None:
                break
            state = step(state, cache[state[2]] != "Failed")
    sec = time.perf_counter() - t0
    out = None
    if rank == 0:
        out = {"workload": "cfg5: C-13 bSSFP spec at 14 T (bSSFP_pulse_lp_ap.m), order search of fir_ap(512, f, a, d, Peak=1e-3, min_order=1)",
               "minimal_order_arbitrary_phase": int(state[1]), "probes_solved": probes, "rounds": rounds, "seconds": sec,
               "instances_per_s": probes / sec, "duration_ms": float(state[1] * dt), "gpus": world}
        t1 = time.perf_counter()
        h, st = fir.fir_min_order_linprog(512, f, a, d, 0, 0, method="ipm")
        out["linear_phase"] = {"call": "fir_min_order_linprog(512, f, a, d)", "status": st, "taps": int(len(h)), "seconds": time.perf_counter() - t1,
                               "duration_ms": float(len(h) * dt)}
    barrier()
    return out


def leg_post(lib):
    """The steps around the solve, one GPU: fir_flip_zero at the reference's cap of 2^12 flip patterns
    on a 256-tap filter (fir_flip_zero.m:45-99: one poly() per pattern in a MATLAB loop; here one CTA per pattern), with the
    numpy restatement of that loop timed on one host core for a sample of the patterns; and one fir_qprog_phs design."""
    from multiband_rf_pulse_design_b200 import fir_post as P
    from oracle import fir_post as O
    rng = np.random.default_rng(0)
    nsb, npb = 243, 12
    zs = np.exp(1j * np.linspace(0.25 * np.pi, 1.75 * np.pi, nsb))
    zp = rng.uniform(0.6, 0.9, npb) * np.exp(1j * rng.uniform(-0.2, 0.2, npb) * np.pi)
    Z = np.concatenate([zs, zp])[rng.permutation(nsb + npb)]
    idx = np.nonzero(np.abs(np.abs(Z) - 1) > 1e-2)[0]
    mask = P.flip_patterns(idx.size)
    P.flip_zero_candidates(Z, idx, mask, 1.0)
    l0 = lib.mbrf_launch_count()
    t = time.perf_counter()
    reps = 10
    for _ in range(reps):
        r = P.flip_zero_candidates(Z, idx, mask, 1.0)
    gpu = (time.perf_counter() - t) / reps
    launches = (lib.mbrf_launch_count() - l0) // reps
    t = time.perf_counter()
    sample = 64
    for i in range(sample):
        Ze = Z.copy()
        Ze[idx] = np.where(mask[i].astype(bool), 1 / np.abs(Z[idx]) * np.exp(1j * np.angle(Z[idx])), Z[idx])
        O.poly_reference(Ze)
    cpu = (time.perf_counter() - t) / sample * mask.shape[0]
    out = {"flip_zero": {"taps": int(Z.size + 1), "patterns": int(mask.shape[0]), "seconds_call": gpu, "patterns_per_s": mask.shape[0] / gpu,
                         "gpu_launches_per_call": int(launches), "best": int(r["best"]),
                         "cpu_baseline": {"seconds": cpu, "cores": 1, "kind": "port",
                                          "sample": "%d of the %d patterns through oracle/fir_post.py:poly_reference (numpy), extrapolated" % (sample, mask.shape[0])}}}
    spec = dict(f=[-0.6, -0.35, -0.15, 0.15, 0.35, 0.6], a=[0, 0, 1, 1, 0, 0], d=[0.02, 0.05 * np.exp(0.2j), 0.02])
    P.fir_qprog_phs(15, spec["f"], spec["a"], spec["d"])
    t = time.perf_counter()
    h, st, ex = P.fir_qprog_phs(63, spec["f"], spec["a"], spec["d"], return_info=True)
    out["fir_qprog_phs"] = {"n": 63, "status": st, "seconds": time.perf_counter() - t, "iterations": float(ex["info"][1]),
                            "rows": int(ex["problem"]["w"].size), "energy": float(np.linalg.norm(ex["x"])),
                            "max_violation": float(ex["info"][4])}
    return out


def solver_cpu_baseline():
    """CPU baseline of the solver path: CVX/SeDuMi/linprog are not installable offline, so the restated problem
    of ONE design of the sweep goes to HiGHS (SciPy) on one host core -- the LP without the 2-D Peak cones.  Bounded sample."""
    from oracle.fir_problems import build_fir_ap, solve_fir_ap_highs
    pr = build_fir_ap(256, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 0.1, 1e-2)
    t3 = time.perf_counter()
    res, _ = solve_fir_ap_highs(pr, 0)
    dt3 = time.perf_counter() - t3
    return {"value": 1.0 / dt3, "unit": "designs/s", "cores": 1, "kind": "port", "seconds": dt3, "highs_status": int(res.status),
            "objective": float(res.fun) if res.status == 0 else None,
            "sample": "1 design of the sweep (obj=0.1, cones dropped), N=256, 7686-row grid: HiGHS via scipy.optimize.linprog on the "
                      "restated LP (oracle/fir_problems.py); obj=1e4 / 1e5 take 400 / 760 s (tests/golden/fir_ap_weights_known.json)"}


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import multiband_rf_pulse_design_b200 as m
    from multiband_rf_pulse_design_b200._lib import check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmbrf has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = m.lib()
    check(lib.mbrf_set_device(local))
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")     # host-side barriers that leave the GPUs idle (an NCCL barrier spins on them)

    wl = make_workload(world)
    nt = wl["b1"].size
    nf, npos = wl["df"].size, wl["dx"].size
    nlocal = NF_PER_GPU * npos
    spin0 = rank * nlocal
    steps_per_gpu = nlocal * nt

    def T(a):
        return torch.tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)

    b1r, b1i, gx = T(wl["b1"].real), T(wl["b1"].imag), T(wl["gx"])
    dts, df, dx = T(np.full(nt, wl["dt"])), T(wl["df"]), T(wl["dx"])
    out = torch.empty((3, nlocal), dtype=torch.float64, device=dev)
    ws = torch.empty(int(lib.mbrf_bloch_workspace_bytes(nt)), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()

    from multiband_rf_pulse_design_b200.shard import bloch_sharded
    dev_args = dict(b1r=b1r.data_ptr(), b1i=b1i.data_ptr(), gx=gx.data_ptr(), gy=None, gz=None, dt=dts.data_ptr(),
                    ntime=nt, t1=wl["t1"], t2=wl["t2"], df=df.data_ptr(), nf=nf, dx=dx.data_ptr(), dy=None, dz=None,
                    npos=npos)

    # The gather of the path is fused into the kernel: rank 0 owns the final [3, S] result, every other rank
    # maps it (CUDA IPC) and its kernel stores its slice straight into it over NVLink -- one kernel per rank, no transfer
    # step; a 1-element all-reduce on the stream orders completion.  (The NCCL variant -- batched send/recv into the final
    # layout, pipelined in two pieces -- stays available: bloch_sharded(..., chunks=[0.85, 0.15]) without `peer`.)
    from multiband_rf_pulse_design_b200.shard import PeerResult
    peer = PeerResult(lib, 3, nf * npos) if world > 1 else None
    peer_t = peer.tensor() if peer is not None else None
    tok = torch.zeros(1, device=dev)

    def step_device():
        r = bloch_sharded(lib, dev_args, nf * npos, out, ws.data_ptr(), stream, 0, m.GAMMA_C13, peer=peer)
        if world > 1:
            dist.all_reduce(tok)         # returns on this stream once every rank's kernel (and its remote stores) has finished
            return peer_t
        return r

    def step_device_nccl():
        return bloch_sharded(lib, dev_args, nf * npos, out, ws.data_ptr(), stream, 0, m.GAMMA_C13, chunks=[0.85, 0.15])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def allgather_obj(o):
        if world == 1:
            return [o]
        outl = [None] * world
        dist.all_gather_object(outl, o, group=cpu_group)
        return outl

    # ---- device-resident leg ---------------------------------------------------------------
    full = None
    for _ in range(args.warmup):
        full = step_device()
    # the gathered planes against this rank simulating other ranks' ranges itself (first and last shard)
    gather_check = None
    if rank == 0 and full is not None:
        torch.cuda.synchronize()
        gather_check = 0.0
        ref = torch.empty((3, nlocal), dtype=torch.float64, device=dev)
        for r in sorted({0, world - 1}):
            check(lib.mbrf_bloch_device(b1r.data_ptr(), b1i.data_ptr(), gx.data_ptr(), None, None, dts.data_ptr(), nt,
                                        wl["t1"], wl["t2"], df.data_ptr(), nf, dx.data_ptr(), None, None, npos,
                                        r * nlocal, nlocal, None, None, None, 1, ref[0].data_ptr(), ref[1].data_ptr(),
                                        ref[2].data_ptr(), 0, m.GAMMA_C13, ws.data_ptr(), stream.cuda_stream))
            torch.cuda.synchronize()
            gather_check = max(gather_check, float((full[:, r * nlocal:(r + 1) * nlocal] - ref).abs().max()))
        del ref
    full = None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    barrier()
    launches0 = lib.mbrf_launch_count()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                # evict L2 between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = lib.mbrf_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    per_step = [a.elapsed_time(b) for a, b in ev]
    ms_dev = sum(per_step) / args.steps
    ms_dev = max_over_ranks(ms_dev)
    ms_dev_min, ms_dev_med = max_over_ranks(min(per_step)), max_over_ranks(statistics.median(per_step))
    value = world * steps_per_gpu / (ms_dev * 1e-3)

    # the NCCL gather variant, for the record
    ms_nccl = None
    if world > 1:
        for _ in range(3):
            step_device_nccl()
        barrier()
        nev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(20, args.steps))]
        for k in range(len(nev)):
            flush.zero_()
            nev[k][0].record(stream)
            step_device_nccl()
            nev[k][1].record(stream)
        barrier()
        ms_nccl = max_over_ranks(sum(a.elapsed_time(b) for a, b in nev) / len(nev))

    # dominant kernel alone (prep + spin kernel, no gather) for the roofline, same stream, L2 flushed
    barrier()
    for k in range(args.steps):
        flush.zero_()
        kev[k][0].record(stream)
        check(lib.mbrf_bloch_device(b1r.data_ptr(), b1i.data_ptr(), gx.data_ptr(), None, None, dts.data_ptr(), nt,
                                    wl["t1"], wl["t2"], df.data_ptr(), nf, dx.data_ptr(), None, None, npos,
                                    spin0, nlocal, None, None, None, 1,
                                    out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), 0, m.GAMMA_C13,
                                    ws.data_ptr(), stream.cuda_stream))
        kev[k][1].record(stream)
    torch.cuda.synchronize()
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    # ---- end-to-end leg: ONE reference-facing call for the whole job, host buffers ------------------
    # What a MEX gateway does: blochC(b1,gr,tp,t1,t2,df,dp,mode) -> mbrf_bloch on rank 0, the result arrays freshly allocated
    # PAGEABLE memory (mxCreateDoubleMatrix), all `world` GPUs used inside the call (mbrf_set_fanout: contiguous spin ranges,
    # each device DMA-ing its slice into a pinned ring, host threads copying it out).  The other ranks idle at a host barrier.
    e2e = None
    host_barrier()
    if rank == 0:
        h_b1 = wl["b1"].reshape(-1, 1)
        h_gr = wl["gx"].reshape(-1, 1)
        h_df = np.ascontiguousarray(wl["df"]).reshape(-1, 1)          # all world * 1000 off-resonances
        h_dp = wl["dx"].reshape(-1, 1)
        ntot = nf * npos
        h2d = 8 * (2 * nt + nt + nt + h_df.size + h_dp.size) * world   # the small inputs go to every device
        d2h = 3 * 8 * ntot
        check(lib.mbrf_set_fanout(world))
        res = {}
        pin = tuple(torch.empty(ntot, dtype=torch.float64).pin_memory().numpy() for _ in range(3))
        for kind in ("pageable", "pinned"):
            def step_e2e():
                return m.blochC(h_b1, h_gr, wl["dt"], wl["t1"], wl["t2"], h_df, h_dp, 0, out=pin if kind == "pinned" else None)
            for _ in range(max(3, args.warmup)):
                last = step_e2e()
            ts = []
            for _ in range(args.steps):
                t0 = time.perf_counter()
                last = step_e2e()                            # synchronous: returns with the result in host memory
                ts.append(time.perf_counter() - t0)
            res[kind] = (sum(ts) / len(ts) * 1e3, min(ts) * 1e3, statistics.median(ts) * 1e3, last)
        check(lib.mbrf_set_fanout(1))
        # the e2e result must be the same numbers as the device-resident leg (rank 0's shard is the first nlocal spins)
        chk = float(np.abs(res["pageable"][3][2].ravel(order="F")[:4096] - out[2][:4096].cpu().numpy()).max())
        ms_e2e = res["pageable"][0]
        e2e = {"value": world * steps_per_gpu / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e2e, "ms_min": res["pageable"][1], "ms_median": res["pageable"][2],
               "call": "ONE blochC(b1,gr,tp,t1,t2,df,dp,0) -> mbrf_bloch (C ABI) for all %d spins, result in freshly allocated PAGEABLE "
                       "arrays (what mxCreateDoubleMatrix gives a MEX gateway), fan-out over %d GPU(s) inside the call" % (ntot, world),
               "check_vs_device_leg": chk,
               "pinned_result_arrays": {"value": world * steps_per_gpu / (res["pinned"][0] * 1e-3), "ms_per_step": res["pinned"][0],
                                        "ms_min": res["pinned"][1], "note": "same call with page-locked result arrays (a C / Python host can "
                                        "supply them, MATLAB cannot): one GPU stores straight into them, several DMA their slices"}}
        del pin, res
    host_barrier()

    # ---- forward SLR (abrx) on the same pulse: 10^6 positions per GPU x 512 samples, device-resident, sharded by
    # contiguous position range (abrx.c:67-78) with the alpha/beta planes gathered on rank 0 like the Bloch planes ----------
    from multiband_rf_pulse_design_b200.shard import abr_sharded
    rf = np.load(os.path.join(ROOT, "tests", "golden", "pulses.npz"))["rf512_rad"] \
        if os.path.exists(os.path.join(ROOT, "tests", "golden", "pulses.npz")) else wl["b1"] * (2 * np.pi * 1.0705 * wl["dt"] * 1e3)
    ns, nx_local = rf.size, 1_000_000
    nx = nx_local * world
    d_rfr, d_rfi = T(rf.real), T(rf.imag)
    d_g, d_x = T(np.full(ns, 2 * np.pi / ns)), T(np.linspace(-40, 40, nx))
    ab = torch.empty((4, nx_local), dtype=torch.float64, device=dev)
    ws2 = torch.empty(int(lib.mbrf_abr_workspace_bytes(ns)), dtype=torch.uint8, device=dev)
    slr_args = dict(rfr=d_rfr.data_ptr(), rfi=d_rfi.data_ptr(), gx=d_g.data_ptr(), gy=None, ns=ns, x=d_x.data_ptr(), nx=nx, y=None, ny=1)

    peer_ab = PeerResult(lib, 4, nx) if world > 1 else None
    peer_ab_t = peer_ab.tensor() if peer_ab is not None else None

    def step_slr():
        r = abr_sharded(lib, slr_args, nx, ab, ws2.data_ptr(), stream, 0, peer=peer_ab)
        if world > 1:
            dist.all_reduce(tok)
            return peer_ab_t
        return r
    for _ in range(3):
        full_ab = step_slr()
    barrier()
    reps = 20
    sev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for k in range(reps):
        flush.zero_()
        sev[k][0].record(stream)
        full_ab = step_slr()
        sev[k][1].record(stream)
    barrier()
    ms_slr = max_over_ranks(sum(a.elapsed_time(b) for a, b in sev) / reps)
    slr = None
    if rank == 0:
        unit = float((full_ab[0] ** 2 + full_ab[1] ** 2 + full_ab[2] ** 2 + full_ab[3] ** 2 - 1).abs().max().item())
        slr = {"metric": "SLR position-steps/sec", "value": nx * ns / (ms_slr * 1e-3), "unit": "position-steps/s",
               "ms_per_call": ms_slr, "positions": nx, "samples": ns, "unitarity_max_err": unit, "n_gpus": world,
               "flops_per_position_step": 50, "call": "mbrf_abr_device (abrx convention) per rank, alpha/beta planes stored straight into rank 0's peer-mapped result, device-resident"}
    full_ab = None

    # ---- second hot path: convex FIR design step (BASELINE metric "N=256 FIR pulse designs solved/sec") --------
    solver = cfg5 = None
    if not args.no_solver:
        solver = leg_cfg4(args, m, lib, rank, world, dev, max_over_ranks, barrier)
        cfg5 = leg_cfg5(m, lib, rank, world, allgather_obj, host_barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline ---------------------------------------------------------------------------
    tf, kms = C.c_double(), C.c_double()
    check(lib.mbrf_measure_fp64_peak(C.byref(tf), C.byref(kms)))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved_tf = FLOPS_PER_SPIN_STEP * steps_per_gpu / (ms_kernel * 1e-3) / 1e12
    pipe_tf = 2 * FP64_INST_PER_SPIN_STEP * steps_per_gpu / (ms_kernel * 1e-3) / 1e12
    hbm_gbs = BYTES_PER_SPIN * nlocal / (ms_kernel * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "bloch_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            traffic = None
    roofline = {
        "bound": "fp64", "kernel": "mbrf::bloch::bloch_kernel<0,1,1,false>",
        "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved_tf / tf.value,
        "traffic": traffic,
        "peak_source": "FP64 FMA peak measured in this run by mbrf_measure_fp64_peak (DFMA-only kernel); "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "flops_per_spin_step": FLOPS_PER_SPIN_STEP, "kernel_ms": ms_kernel,
        "fp64_pipe": {"inst_per_spin_step": FP64_INST_PER_SPIN_STEP, "issued_tflops_equiv": pipe_tf,
                      "frac": pipe_tf / tf.value},
        "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                "note": "56 algorithmic bytes per spin per call: the path is FP64-bound, not HBM-bound"},
    }

    if slr:
        slr["roofline"] = {"bound": "fp64", "achieved": slr["value"] * 50 / 1e12 / world, "peak": tf.value, "unit": "TFLOP/s per GPU",
                           "frac": slr["value"] * 50 / 1e12 / world / tf.value}
    if solver is not None:
        solver["roofline"] = solver_roofline(lib, tf.value)
        solver["single_design"] = leg_cfg3(m, lib, hbm_peak) if not args.skip_single_design else None
        solver["order_search"] = cfg5
        solver["post"] = leg_post(lib)
        if world == 1 and not args.no_cpu_baseline:
            solver["cpu_baseline"] = solver_cpu_baseline()

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nf_pt = 30
        v, sec, kind = cpu_reference_run(wl, nf_pt, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "seconds": sec,
               "sample": f"{threads} threads x {nf_pt} df x {NPOS} dp x {nt} samples of the same workload "
                         f"(blochC.c compiled -O2, one blochsimfz call per thread)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "ms_per_step_min": ms_dev_min, "ms_per_step_median": ms_dev_med,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "pulse": wl["src"], "spins_per_gpu": nlocal, "ntime": nt,
                   "sharding": f"contiguous spin ranges over {world} GPU(s); every rank's kernel stores its slice straight into rank 0's [3, S] result "
                               f"over NVLink (peer-mapped, CUDA IPC), completion ordered by a 1-element all-reduce; NCCL send/recv gather variant: {ms_nccl} ms per step"
                   if world > 1 else "single GPU", "l2": "flushed between timed iterations (256 MiB memset)", "gather_check_max_abs": gather_check},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        # short top-level copies of the second hot path's headline numbers (the full objects follow)
        "solver_designs_per_s": solver["value"] if solver else None,
        "solver_roofline_frac": solver["roofline"]["frac"] if solver else None,
        "solver_cfg3_seconds": solver["single_design"]["grid4096"]["seconds"] if solver and solver.get("single_design") else None,
        "solver_cfg3_hbm_frac": solver["single_design"]["grid4096"]["roofline"]["frac"] if solver and solver.get("single_design") else None,
        "solver_cfg5_seconds": solver["order_search"]["seconds"] if solver and solver.get("order_search") else None,
        "slr_position_steps_per_s": slr["value"] if slr else None,
        "slr": slr,
        "solver": solver,
    }
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """Keep the real stdout for the ONE JSON line; anything libraries print on fd 1 (e.g. NCCL's version banner)
    goes to stderr instead."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solver", action="store_true", help="skip the FIR-design leg")
    ap.add_argument("--solver-grid", type=lambda v: tuple(int(x) for x in v.split(",")), default=(16, 16, 16),
                    help="cfg4 sweep: numbers of obj, Peak and f_add values (default 16,16,16 = 4096 designs, BASELINE config 4)")
    ap.add_argument("--skip-single-design", action="store_true",
                    help="skip cfg3 (a million small launches: only for runs under a profiler)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 10:
            args.steps, args.warmup = 10, 3
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
