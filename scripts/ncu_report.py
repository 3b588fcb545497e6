#!/usr/bin/env python
"""One text summary of an .ncu-rep for profiles/: headline metrics (ncu_summary.py), stall reasons per issue, DRAM traffic.
usage: ncu_report.py <report.ncu-rep> > profiles/<name>.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
print(f"# {rep}  (ncu --set full --clock-control none, first profiled launch)")
for i, k in enumerate(h):
    if k in want:
        print(f"{k:72s} {v[i]:>24s} {u[i]}")
print("# warp stall reasons, per issued instruction")
for i, k in enumerate(h):
    if "per_issue_active" in k and "stalled" in k:
        try:
            x = float(v[i])
        except ValueError:
            continue
        if x > 0.02:
            print(f"{k.replace('smsp__average_warps_issue_stalled_', ''):72s} {x:24.3f}")
