#!/usr/bin/env python3
"""Compact per-kernel table of the metrics that matter for the trace kernels (development aid)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = [("ms", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("fmaH%", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("l1tp%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), ("l1hit", "l1tex__t_sector_hit_rate.pct"), ("l2tp%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2hit", "lts__t_sector_hit_rate.pct"), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("Minst", "smsp__inst_executed.sum"),
        ("st_long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"), ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"), ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("st_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"), ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("st_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"), ("st_nois", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        ("st_disp", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"), ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")]
ki = hdr.index("Kernel Name")
idx = [(n, hdr.index(m)) for n, m in want if m in hdr]
print("%-22s" % "kernel" + "".join("%9s" % n for n, _ in idx))
for r in rows[2:]:
    name = r[ki].replace("void ", "").split("(")[0][:22]
    out = "%-22s" % name
    for n, i in idx:
        try:
            v = float(r[i].replace(",", ""))
            out += "%9.2f" % (v / 1e6 if n == "Minst" else v)
        except ValueError:
            out += "%9s" % r[i][:8]
    print(out)
