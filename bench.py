#!/usr/bin/env python
"""bench.py — Bloch hot path on B200 (BASELINE.json metric "Bloch spin-steps/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (`config.workload`): BASELINE configs[1] — a 512-sample excitation pulse simulated
over 10^6 spins per GPU (1000 off-resonances x 1000 positions with a 0.05 G/cm x-gradient,
T1 = T2 = 1000 s, mode 0, 13C gamma), i.e. 5.12e8 spin-steps per step per GPU.  At N GPUs the
job is 1000*N off-resonances x 1000 positions, sharded by contiguous spin range
(s = p + npos*f, blochC.c:468-473) with one NCCL gather of mx/my/mz to rank 0 per step.

One JSON line is printed by rank 0:
  value        device-resident throughput (inputs already in HBM, mbrf_bloch_device + gather),
               timed with CUDA events on the launching stream, max over ranks
  e2e          the same metric through the reference-facing call  blochC(b1,gr,tp,t1,t2,df,dp,mode)
               (C ABI mbrf_bloch) with pinned HOST buffers; H2D of the inputs and the result
               landing in host memory are inside the timed region
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
# solver leg (reported inside the same JSON line under "solver")
# ----------------------------------------------------------------------------
H1_DUALBAND = dict(f=[-0.047006, -0.027115, -0.016335, 0.013779, 0.029671, 0.047006],       # SURVEY.md 8(d):
                   a=[0.865905, 0.865905, 0.0, 0.0, 0.706886, 0.706886],                     # specsat_H1_dualband.m:5-32
                   d=[0.014436, 0.022361, 0.017683])                                         # after dzrf_mb's shift_f


def solver_leg(args, m, lib, rank, world, max_over_ranks, barrier):
    """BASELINE config 4 shape: N=256 fir_ap_cvx designs of the dual-band H-1 saturation spec, a slice of the
    obj x Peak trade-off grid (band edges fixed), `--solver-designs` per GPU, sharded by design instance.
    One untimed warm-up batch of 64 designs, then ONE timed solve of the whole local share through the public
    API fir_ap_cvx_sweep (host assembly + H2D + GPU solve + D2H inside the timed region)."""
    from multiband_rf_pulse_design_b200 import fir
    n = 256
    per_gpu = args.solver_designs
    total = per_gpu * world
    n_obj = max(1, total // 8)                          # 8 Peak values x n_obj weights
    objs = np.logspace(-2, 1, n_obj)                    # stop-band weight; SURVEY.md 8(d) asks logspace(-2,4): above ~10 the
                                                        # ripple-dominated objective converges too slowly for a bench leg
    peaks = np.logspace(-3.2, -2, 8)                    # Peak values that keep the spec feasible at N=256
    fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs[:8], peaks, [0.0],
                         max_iter=512)                  # warm-up: context, buffers, kernels
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.mbrf_launch_count()
    t0 = time.perf_counter()
    r = fir.fir_ap_cvx_sweep(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], objs, peaks, [0.0], rank=rank,
                             world=world, batch=per_gpu, max_iter=60000, seed_stride="auto")
    t1s = time.perf_counter()
    sec = max_over_ranks(t1s - t0)
    clocks = sampler.stop(t0, t1s) if sampler else None
    launches = lib.mbrf_launch_count() - l0
    barrier()
    single = None
    if rank == 0:
        t1 = time.perf_counter()
        _, st3, ex3 = fir.fir_qp_cvx(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 120, 1.0, return_info=True)
        single = {"workload": "cfg3: fir_qp_cvx min-energy multiband FIR, N=256, dual-band H-1 spec, k=120, obj=1, single design",
                  "status": st3, "seconds": time.perf_counter() - t1, "iterations": float(ex3["info"][1]),
                  "objective": float(ex3["info"][2]), "max_violation": float(ex3["info"][4])}
    roof, fmp, osearch = None, None, None
    if rank == 0 and getattr(args, "order_search", False):
        # BASELINE config 5 shape: fir_ap(..., min_order=1) = bisection over the order with fir_ap_cvx probes (fir_ap.m:143-162),
        # here from n = 512 on the dual-band H-1 spec (the C-13 bSSFP spec needs the toolbox's spec builders, out of scope);
        # every round of the search solves its speculative probes concurrently (different orders = different matrices)
        t4 = time.perf_counter()
        _, st4, n_op, _ = fir.fir_ap(512, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], 1e-3, 1, 0, 0, max_iter=60000)
        osearch = {"workload": "fir_ap(512, f, a, d, Peak=1e-3, min_order=1): minimal order of the dual-band H-1 spec, orders searched 1..512",
                   "status": st4, "minimal_order": int(n_op), "seconds": time.perf_counter() - t4}
    if rank == 0:
        roof = solver_roofline(lib, per_gpu)
        # the step after the solve (fir_ap_cvx.m:185-202): x -> minimum-phase taps h = fmp2(r), batched on the GPU
        ok = r["info"][:, 0] == 1
        if ok.any():
            R = np.stack([fir._x_to_r(x, n) for x in r["x"][ok]])
            fir.fmp2_batch(R[:8])
            t2 = time.perf_counter()
            fir.fmp2_batch(R)
            dt = time.perf_counter() - t2
            fmp = {"call": "fmp2_batch -> mbrf_fmp2_batch (one CTA per design, four 4096-point fp64 FFTs in shared memory), host buffers",
                   "designs": int(ok.sum()), "seconds": dt, "designs_per_s": float(ok.sum() / dt)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU baseline of the solver path: CVX/SeDuMi/linprog are not installable offline, so the restated
        # problem of ONE design of this sweep goes to HiGHS (SciPy) on one host core -- the LP without the 2-D Peak cones,
        # i.e. less work than the reference solve.  Bounded sample: one design (~50 s).
        from oracle.fir_problems import build_fir_ap, solve_fir_ap_highs
        o_mid, p_mid = float(objs[len(objs) // 2]), float(peaks[len(peaks) // 2])
        pr = build_fir_ap(n, H1_DUALBAND["f"], H1_DUALBAND["a"], H1_DUALBAND["d"], o_mid, p_mid)
        t3 = time.perf_counter()
        res, _ = solve_fir_ap_highs(pr, 0)
        dt3 = time.perf_counter() - t3
        cpu = {"value": 1.0 / dt3, "unit": "designs/s", "cores": 1, "kind": "port", "seconds": dt3,
               "highs_status": int(res.status), "objective": float(res.fun) if res.status == 0 else None,
               "sample": "1 design of the same sweep (obj=%.3g, Peak=%.3g), N=256, 7686-row grid: HiGHS via scipy.optimize.linprog on "
                         "the restated LP without the Peak cones (oracle/fir_problems.py); CVX/SeDuMi are not installable offline" % (o_mid, p_mid)}
    info = r["info"]
    solved = int((info[:, 0] == 1).sum())
    iters = float(info[:, 1].max()) if info.size else 0.0
    # the dominant kernels are the two fp64 GEMMs of every iteration: 2 * 2*Mp*Np*Bp flops per iteration
    Mp, Np, Bp = C.c_int(), C.c_int(), C.c_int()
    lib.mbrf_pdhg_padded_sizes(7686 + 118, 2 * n, per_gpu, C.byref(Mp), C.byref(Np), C.byref(Bp))
    # useful GEMM work: every design pays 2 products of 2*Mp*Np flops per iteration it was still in the batch
    flops = 4.0 * Mp.value * Np.value * float(info[:, 1].sum()) * world
    return {"metric": "N=256 FIR pulse designs solved/sec", "value": total / sec, "unit": "designs/s",
            "designs": total, "designs_per_gpu": per_gpu, "sweep": "seed_stride=auto: where the obj grid is finer than 0.0125 decades "
            "(4 and 8 GPUs) every ~0.1 decade is solved cold and the other designs start from the nearest seed; coarser grids run cold", "solved_on_rank0": solved, "local_designs_rank0": int(info.shape[0]),
            "seconds": sec, "iterations_max": iters, "gpu_launches": int(launches),
            "workload": "cfg4 slice: fir_ap_cvx, dual-band H-1 sat spec, N=256, 7686-row grid, obj x Peak trade-off grid",
            "single_design": single, "roofline": roof, "fmp2": fmp, "cpu_baseline": cpu, "clocks": clocks, "order_search": osearch,
            "tolerances": {"eps_pr": fir.EPS_PR, "eps_gap_rel": fir.EPS_GAP, "eps_dr": fir.EPS_DR},
            "caveat": "the termination test is checked against HiGHS within 1e-4 for stop-band weights <= 1; for the weights > 2.5 of "
                      "this grid (a quarter of the designs) the numpy twin of the solver ends 2.6e-4 (weight 4) to 6.7e-4 (weight 10) "
                      "above the HiGHS optimum (DESIGN.md 6, known weak spot)",
            "gemm_tflops_useful": flops / sec / 1e12, "iterations_mean": float(info[:, 1].mean()) if info.size else 0.0,
            "note": "fp64 restarted PDHG; the two products of every iteration run on tcgen05 int8 tiles (split-integer, 5 base-256 "
                    "digit planes, exact int32 level sums in TMEM, ~1e-11 relative), convergence checks on fp64 mma.sync tiles; whole "
                    "call timed on the host (assembly + PCIe + solve); gemm_tflops_useful = fp64-equivalent 4*Mp*Np*sum_b(iterations_b) "
                    "/ wall time (rank 0's designs x world)"}


def solver_roofline(lib, per_gpu, nd=5):
    """Roofline of the solver's dominant kernel, tc::tc_i8_gemm_kernel<5>: the two products of one PDHG iteration at the
    bench shape (K: 7808 x 512, `per_gpu` designs), each timed alone with CUDA events inside libmbrf
    (mbrf_tc_product_device).  Tensor bound: int8 operations 2*R*k*B per digit-plane product, nd(nd+1)/2 = 15 products."""
    import torch
    from multiband_rf_pulse_design_b200._lib import check
    Mp, Np, Bp = C.c_int(), C.c_int(), C.c_int()
    lib.mbrf_pdhg_padded_sizes(7686 + 118, 511, per_gpu, C.byref(Mp), C.byref(Np), C.byref(Bp))
    Mp, Np, Bp = Mp.value, Np.value, Bp.value
    if Bp < 64:
        return None
    g = torch.Generator(device="cuda").manual_seed(0)
    K = torch.rand((Mp, Np), dtype=torch.float64, device="cuda", generator=g) - 0.5
    KT = K.t().contiguous()
    z = torch.rand((Np, Bp), dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.rand((Mp, Bp), dtype=torch.float64, device="cuda", generator=g) - 0.5
    tiles = (Np // 64) * ((Bp + 127) // 128)
    P = max(1, min(32, Mp // 256, (2 * 148) // tiles))          # same split as pdhg.cu:split_k_tc
    out = torch.empty((max(P, 1), max(Mp, Np), Bp), dtype=torch.float64, device="cuda")
    ms = {}
    for name, (A, X, R, kd, ns) in {"K*zbar": (K, z, Mp, Np, 1), "K^T*y": (KT, y, Np, Mp, P)}.items():
        a, b = C.c_float(), C.c_float()
        check(lib.mbrf_tc_product_device(A.data_ptr(), R, kd, X.data_ptr(), Bp, nd, ns, out.data_ptr(), 20,
                                         C.cast(C.byref(a), C.c_void_p), C.cast(C.byref(b), C.c_void_p), None))
        ms[name] = (a.value, b.value)
    torch.cuda.synchronize()
    ops = 2.0 * Mp * Np * Bp * (nd * (nd + 1) // 2)              # int8 multiply-adds x2, per product
    t = ms["K*zbar"][0] + ms["K^T*y"][0]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    if "bf16_tflops" in peaks:
        peak, src = 2.0 * float(peaks["bf16_tflops"]), "2 x MEASURED_PEAKS.json bf16_tflops (int8 issues at twice the bf16 rate; no int8 figure is recorded)"
    else:
        peak, src = 4500.0, "fallback: nominal dense int8 4.5 POP/s (B200_PROFILING.md: 2 x bf16 2.25 PFLOP/s)"
    ach = 2 * ops / (t * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "solver_tc_traffic.json"))).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    return {"bound": "tensor", "kernel": "mbrf::tc::tc_i8_gemm_kernel<%d>" % nd, "achieved": ach, "peak": peak, "unit": "TOP/s (int8)",
            "frac": ach / peak, "traffic": traffic, "peak_source": src,
            "fp64_equivalent_tflops": 2 * 2.0 * Mp * Np * Bp / (t * 1e-3) / 1e12,
            "kernel_ms": {k: v[0] for k, v in ms.items()}, "with_digit_planes_ms": {k: v[1] for k, v in ms.items()},
            "shape": {"Mp": Mp, "Np": Np, "Bp": Bp, "digit_planes": nd, "split_k": P},
            "note": "algorithmic work = 2*Mp*Np*Bp int8 MACs x2 per digit-plane product x 15 products x 2 GEMMs per iteration; the "
                    "fp64 tensor path this replaced (mma.sync m8n8k4) ran the same two products in 0.139 + 0.148 ms"}


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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

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

    def step_device():
        # this rank's contiguous spin range, then the one collective of the path: the final gather (NCCL/NVLink)
        # (pipelined in two pieces: the first piece's transfer hides behind the simulation of the second)
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
    ms_dev = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    ms_dev = max_over_ranks(ms_dev)
    value = world * steps_per_gpu / (ms_dev * 1e-3)

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

    # ---- end-to-end leg: the reference-facing call on pinned host buffers --------------------
    h_out = [torch.empty(nlocal, dtype=torch.float64).pin_memory() for _ in range(3)]
    h_out_np = tuple(t.numpy() for t in h_out)
    h_b1 = wl["b1"].reshape(-1, 1)
    h_gr = wl["gx"].reshape(-1, 1)
    h_df = np.ascontiguousarray(wl["df"][rank * NF_PER_GPU:(rank + 1) * NF_PER_GPU]).reshape(-1, 1)
    h_dp = wl["dx"].reshape(-1, 1)
    h2d = 8 * (2 * nt + nt + nt + h_df.size + h_dp.size)   # b1 re/im, gradient, time steps, df, dp
    d2h = 3 * 8 * nlocal

    def step_e2e():
        m.blochC(h_b1, h_gr, wl["dt"], wl["t1"], wl["t2"], h_df, h_dp, 0, out=h_out_np)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()                                   # synchronous: returns with the result in host memory
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) / args.steps * 1e3
    ms_e2e = max_over_ranks(ms_e2e)
    barrier()
    e2e_value = world * steps_per_gpu / (ms_e2e * 1e-3)
    # the e2e result must be the same numbers as the device-resident leg
    chk = float(np.abs(h_out_np[2][:4096] - out[2][:4096].cpu().numpy()).max())

    # ---- forward SLR (abrx) on the same pulse: 10^6 positions x 512 samples, device-resident -------------
    slr = None
    if rank == 0:
        rf = np.load(os.path.join(ROOT, "tests", "golden", "pulses.npz"))["rf512_rad"] \
            if os.path.exists(os.path.join(ROOT, "tests", "golden", "pulses.npz")) else wl["b1"] * (2 * np.pi * 1.0705 * wl["dt"] * 1e3)
        ns, nx = rf.size, 1_000_000
        d_rfr, d_rfi = T(rf.real), T(rf.imag)
        d_g, d_x = T(np.full(ns, 2 * np.pi / ns)), T(np.linspace(-40, 40, nx))
        ab = torch.empty((4, nx), dtype=torch.float64, device=dev)
        ws2 = torch.empty(int(lib.mbrf_abr_workspace_bytes(ns)), dtype=torch.uint8, device=dev)

        def step_slr():
            check(lib.mbrf_abr_device(d_rfr.data_ptr(), d_rfi.data_ptr(), d_g.data_ptr(), None, ns, d_x.data_ptr(), nx,
                                      None, 1, 0, 0, nx, ab[0].data_ptr(), ab[1].data_ptr(), ab[2].data_ptr(),
                                      ab[3].data_ptr(), ws2.data_ptr(), stream.cuda_stream))
        for _ in range(3):
            step_slr()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for _ in range(reps):
            step_slr()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_slr = e0.elapsed_time(e1) / reps
        unit = float((ab[0] ** 2 + ab[1] ** 2 + ab[2] ** 2 + ab[3] ** 2 - 1).abs().max().item())
        slr = {"metric": "SLR position-steps/sec", "value": nx * ns / (ms_slr * 1e-3), "unit": "position-steps/s",
               "ms_per_call": ms_slr, "positions": nx, "samples": ns, "unitarity_max_err": unit,
               "flops_per_position_step": 50, "call": "mbrf_abr_device (abrx convention), device-resident"}

    # ---- second hot path: convex FIR design step (BASELINE metric "N=256 FIR pulse designs solved/sec") --------
    solver = None
    if not args.no_solver:
        solver = solver_leg(args, m, lib, rank, world, max_over_ranks, barrier)

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
        slr["roofline"] = {"bound": "fp64", "achieved": slr["value"] * 50 / 1e12, "peak": tf.value, "unit": "TFLOP/s",
                           "frac": slr["value"] * 50 / 1e12 / tf.value}

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
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "pulse": wl["src"], "spins_per_gpu": nlocal, "ntime": nt,
                   "sharding": f"contiguous spin ranges over {world} GPU(s), one NCCL gather to rank 0 per step"
                   if world > 1 else "single GPU", "l2": "flushed between timed iterations (256 MiB memset)", "gather_check_max_abs": gather_check},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "call": "blochC(b1,gr,tp,t1,t2,df,dp,0) -> mbrf_bloch (C ABI), pinned host buffers",
                "check_vs_device_leg": chk},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "solver": solver,
        "slr": slr,
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
    ap.add_argument("--solver-designs", type=int, default=512, help="fir_ap_cvx designs per GPU in the solver leg")
    ap.add_argument("--order-search", action="store_true",
                    help="also time the arbitrary-phase order search from n=512 (BASELINE config 5 shape, ~16 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 10:
            args.steps, args.warmup = 10, 3
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
