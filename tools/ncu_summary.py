#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` dump: one column per captured kernel launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__warps_eligible.avg.per_cycle_active"]
want += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for w in want:
    if w not in idx:
        continue
    i = idx[w]
    vals = []
    for r in data:
        v = r[i]
        try:
            f = float(v.replace(",", "")); v = "%.3g" % f if abs(f) < 1e6 else "%.3e" % f
        except ValueError:
            v = v[-34:]
        vals.append(v)
    if w.startswith("smsp__average_warps_issue_stalled") and all(float(x) < 0.15 for x in vals): continue
    print("%-72s %-6s %s" % (w.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")[:72], units[i][:6], " | ".join(vals)))
