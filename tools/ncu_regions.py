#!/usr/bin/env python3
"""Bucket per-instruction counters of one kernel (ncu --page source --csv) into address regions (development aid).
usage: ncu_regions.py src.csv <kernel index> 0x0:0x5020:prologue 0x5020:0x5d40:node ..."""
import csv, sys
allrows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
which = int(sys.argv[2])          # index of the kernel inside the file
rows = allrows[starts[which]:starts[which + 1]]
print(rows[0][1][:60])
hdr = rows[1]
ia, ie, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
regs = []
for a in sys.argv[3:]:
    lo, hi, name = a.split(":")
    regs.append((int(lo, 16), int(hi, 16), name))
tot = {}
for r in rows[2:]:
    off = int(r[ia], 16) - base
    name = "other"
    for lo, hi, n in regs:
        if lo <= off < hi: name = n
    t = tot.setdefault(name, [0, 0, 0, 0])
    t[0] += int(r[ie]); t[1] += int(r[it]); t[2] += int(r[isamp]); t[3] += 1
E = sum(t[0] for t in tot.values()); S = sum(t[2] for t in tot.values())
for n, t in tot.items():
    print("%-12s sass %5d  warp-inst %8.1fM (%5.1f%%)  lanes %5.2f  samples %5.1f%%" % (n, t[3], t[0] / 1e6, 100 * t[0] / E, t[1] / max(t[0], 1), 100 * t[2] / max(S, 1)))
