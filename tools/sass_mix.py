#!/usr/bin/env python
"""Instruction mix of one profiled kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view):
warp-level executed instructions and stall samples per opcode.  Usage: sass_mix.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "# Samples", "Thread Instructions Executed")}
mix = collections.Counter(); smp = collections.Counter(); n = collections.Counter()
tot = tots = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    src = r[ci["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] if "--full" not in sys.argv else op
    e = int(r[ci["Instructions Executed"]]); s = int(r[ci["# Samples"]])
    mix[op] += e; smp[op] += s; n[op] += 1; tot += e; tots += s
print(f"static SASS instructions {sum(n.values())}, executed warp instructions {tot}, samples {tots}")
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 30
for op, e in mix.most_common(top):
    print(f"{op:12s} static {n[op]:5d}  executed {e:12d} {100*e/tot:5.1f}%   samples {100*smp[op]/max(tots,1):5.1f}%")
