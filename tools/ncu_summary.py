#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics + stall-sample breakdown per kernel.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__sass_inst_executed_op_local_ld.sum", "sm__sass_inst_executed_op_local_st.sum"]
for r in rows[2:]:
    print("=== kernel:", r[hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:70s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    st = []
    for i, h in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                st.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("  warp-state samples (pc sampling):")
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"    {h:28s} {100 * v / tot:5.1f} %")
