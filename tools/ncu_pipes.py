#!/usr/bin/env python3
"""Print the pipe / stall metrics of every kernel in an .ncu-rep (development aid)."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
pat = sys.argv[2:] or ["gpu__time_duration.sum", "launch__registers", "warps_active.avg.pct", "issue_active.avg.pct", "inst_executed_pipe_", "pipe_fma", "pipe_alu",
                       "thread_inst_executed_per_inst", "l1tex__throughput", "l1tex__t_sector_hit", "lts__t_sector_hit", "smsp__inst_executed.sum", "warp_issue_stalled.*pct", "lsu"]
import re
cols = [i for i, h in enumerate(hdr) if any(re.search(p, h) for p in pat)]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:40])
    for i in cols:
        v = r[i]
        try:
            if float(v.replace(",", "")) == 0: continue
        except ValueError:
            pass
        print("   %-90s %s %s" % (hdr[i], v, rows[1][i]))
