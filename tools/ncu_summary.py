#!/usr/bin/env python
"""Key raw metrics of one ncu report:  python tools/ncu_summary.py rep.ncu-rep [events]"""
import csv, subprocess, sys
rep = sys.argv[1]
events = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_elapsed.sum", "smsp__cycles_active.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for k in keys:
    if k in m:
        print(f"{k:85s} {m[k][0]} {m[k][1]}")
wi = float(m["smsp__inst_executed.sum"][0]); ti = wi * float(m["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
ce = float(m["smsp__cycles_elapsed.sum"][0])
print(f"issue-slot utilisation (inst_executed / cycles_elapsed) {wi / ce:.4f}")
print(f"lane efficiency (thread_inst / 32 inst)               {ti / (32 * wi):.4f}")
if events:
    print(f"warp-instructions per event                            {wi / events:.2f}")
    print(f"lane-instructions per event                            {ti / events:.2f}")
