#!/usr/bin/env python3
"""Developer tool: summarise an `ncu --page source --csv` export: stall samples by SASS opcode class and the top lines.
usage: ncu_src_top.py source.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# an export can hold several views (or kernels) one after the other: keep the first
for i in range(2, len(rows)):
    if rows[i] and rows[i][0] == "Kernel Name":
        rows = rows[:i]
        break
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]; SRC = ix["Source"]; EX = ix["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; byop = collections.Counter(); exop = collections.Counter(); bystall = collections.Counter()
lines = []
for r in rows[2:]:
    if len(r) <= S: continue
    n = int(r[S] or 0); tot += n
    op = r[SRC].split()[0] if r[SRC].split() else "?"
    if op.startswith("@"): op = r[SRC].split()[1]
    op = op.split(".")[0] + ("." + r[SRC].split(".")[1].split()[0] if op in ("LDG", "STG", "LDS", "STS", "LDL", "STL") and "." in r[SRC] else "")
    byop[op] += n; exop[op] += int(r[EX] or 0)
    for s in stalls: bystall[s] += int(r[ix[s]] or 0)
    lines.append((n, r[SRC].strip(), {s: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}))
print("total samples", tot, " instructions executed", sum(exop.values()))
print("by stall:", {k: v for k, v in bystall.most_common(8)})
print("samples by opcode (where the warp waits):")
for op, n in byop.most_common(14):
    print("  %-14s %7d %5.1f%%   executed %d" % (op, n, 100.0 * n / tot, exop[op]))
print("top lines:")
for n, src, st in sorted(lines, key=lambda x: -x[0])[:N]:
    print("  %6d %-60s %s" % (n, src[:60], dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])))
