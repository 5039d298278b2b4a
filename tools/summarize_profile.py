#!/usr/bin/env python3
"""Turn gpurun_out/*.ncu-rep / launch lists into the small tracked summaries under profiles/."""
import collections
import csv
import subprocess
import sys


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("trd::", "")
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        tot[name] += ms
        cnt[name] += 1
    s = sum(tot.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, v in tot.most_common():
            f.write("%s,%d,%.4f,%.2f\n" % (k, cnt[k], v, 100 * v / s))


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
    want = [w for w in want if w in idx]
    with open(out, "w") as f:
        f.write(",".join(want) + "\n")
        f.write(",".join(rows[1][idx[w]] for w in want) + "\n")
        for r in rows[2:]:
            f.write(",".join('"%s"' % r[idx[w]].split("(")[0] if w == "Kernel Name" else r[idx[w]].replace(",", "") for w in want) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
