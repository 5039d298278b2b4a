#!/usr/bin/env python3
"""Attribute a kernel's executed instructions / stall samples (ncu --page source --csv) to source lines, using the
inline chain nvdisasm -gi prints (development aid).
usage: ncu_by_line.py all.dis <mangled kernel> src.csv <kernel index in csv> <file suffix> [depth]
Aggregates by the line of <file suffix> that is closest to the kernel in the inline chain (depth 0) etc."""
import csv, re, sys, collections
dis, kern, src, kidx, fsuffix = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
lines = open(dis).read().split("\n")
start = lines.index(".text.%s:" % kern)
addr_chain = {}
chain = []
pending = []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or (l.startswith(".text.") and l.endswith(":")):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m:
        if pending:
            chain = pending
            pending = []
        addr_chain[int(m.group(1), 16)] = chain
rows = list(csv.reader(open(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
k = rows[starts[kidx]:starts[kidx + 1]]
hdr = k[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(k[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0])
for r in k[2:]:
    off = int(r[ia], 16) - base
    ch = addr_chain.get(off, [])
    key = None
    # the chain is innermost first; take the OUTERMOST entry inside the wanted file (closest to the kernel)
    for f, ln in reversed(ch):
        if f.endswith(fsuffix):
            key = ln
            break
    if key is None:
        key = -1
    agg[key][0] += int(r[ie]); agg[key][1] += int(r[isamp])
E = sum(v[0] for v in agg.values()); S = sum(v[1] for v in agg.values())
src_lines = {}
try:
    for i, l in enumerate(open("/root/repo/tauray_b200/csrc/" + fsuffix.split("/")[-1]), 1):
        src_lines[i] = l.rstrip()
except Exception:
    pass
for key, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%5d  inst %5.1f%%  samples %5.1f%%  %s" % (key, 100 * e / E, 100 * s / max(S, 1), src_lines.get(key, "")[:110].strip()))
