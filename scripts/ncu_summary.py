#!/usr/bin/env python3
"""Summarises an .ncu-rep (one kernel capture, `ncu --set full --import-source on`) into a small text
file for profiles/: launch shape, duration, DRAM traffic, pipe utilisation, stall mix, SASS opcode mix.
Usage: scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt   (runs without a GPU)"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2",  # 2 x DFMA thread instructions = executed DFMA flops (millions)
]
for r in rows[2:]:
    print("== raw metrics ==")
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"{h:75s} {r[i]:>24s} {units[i]}")
    print("-- warp stall reasons (avg warps stalled per issue-active cycle) --")
    st = [(h, float(r[i])) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for h, v in sorted(st, key=lambda kv: -kv[1])[:8]:
        print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    ops, samples, tot, tots = collections.Counter(), collections.Counter(), 0, 0
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        try:
            n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += n; samples[op] += s; tot += n; tots += s
    print("== SASS opcode mix (warp instructions executed; share of stall samples) ==")
    for op, n in ops.most_common(18):
        print(f"  {op:10s} {n:14d} {100 * n / tot:6.1f}%   samples {100 * samples[op] / max(tots, 1):5.1f}%")
    tma = [op for op in ops if op.startswith("UBLKCP") or op.startswith("UTMA") or op.startswith("SYNCS")]
    print("  TMA / mbarrier opcodes present:", ", ".join(f"{o}={ops[o]}" for o in tma) or "none")
