#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers we track; usage: ncu_summary.py rep [kernel-regex]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_global_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("=" * 100)
    d = dict(zip(hdr, r))
    for w in want:
        if w in d: print(f"{w:75s} {d[w]:>18s} {units[hdr.index(w)]}")
    st = {h: float(d[h]) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and d[h]}
    tot = sum(st.values()) or 1
    print("stall samples:", ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_','')}={100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
