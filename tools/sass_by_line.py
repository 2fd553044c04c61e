#!/usr/bin/env python
"""Join `nvdisasm -g` line info of a kernel with the per-instruction counters of an ncu source page (SASS view, csv)
and report executed warp instructions / stall samples per source region.
Usage: sass_by_line.py <cubin> <mangled kernel substring> <ncu_source.csv> [--lines]"""
import csv, re, subprocess, sys, collections
cubin, kern, csvf = sys.argv[1:4]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
line_of = {}
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m: line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(csvf)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[h + 1][0], 16)
# regions = the functions of elastic_math.h / local_bodies.h, found from the source text (pass the directory with --src DIR)
import os
SRC = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "admm-elastic-sca_b200", "csrc")
def functions(path):
    out = []
    if not os.path.exists(path): return out
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r'^(?:\s{0,1})(?:static\s+)?(?:ADMMB_\w+|__device__ __forceinline__)\s+[\w:<>\s\*&]*?\b(\w+)\s*\(', l)
        if m and not l.startswith("\t\t"): out.append((i, m.group(1)))
    return out
FUNCS = {f: functions(os.path.join(SRC, f)) for f in ("elastic_math.h", "local_bodies.h")}
def region(f, ln):
    fs = FUNCS.get(f)
    if not fs: return f
    name = f
    for start, nm in fs:
        if start <= ln: name = nm
        else: break
    return name
ex = collections.Counter(); sm = collections.Counter(); fp = collections.Counter(); perline = collections.Counter(); pls = collections.Counter()
tot = tots = 0
for r in rows[h + 1:]:
    if len(r) < len(hdr): continue
    off = int(r[0], 16) - base
    (f, ln), ins = line_of.get(off, (("?", 0), ""))
    e, s = int(r[iE]), int(r[iN])
    k = region(f, ln)
    ex[k] += e; sm[k] += s; tot += e; tots += s
    op = r[iS].split()[1] if r[iS].strip().startswith("@") else r[iS].split()[0]
    if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"): fp[k] += e
    perline[(f, ln)] += e; pls[(f, ln)] += s
print(f"{'region':18s} {'executed':>12s} {'%':>6s} {'fp64-pipe':>12s} {'samples%':>8s}")
for k, e in ex.most_common():
    print(f"{k:18s} {e:12d} {100*e/tot:6.1f} {fp[k]:12d} {100*sm[k]/tots:8.1f}")
print(f"{'total':18s} {tot:12d}        {sum(fp.values()):12d}")
if "--lines" in sys.argv:
    for (f, ln), e in perline.most_common(60):
        print(f"{f}:{ln:5d} {e:12d} {100*e/tot:5.1f}%  samples {100*pls[(f,ln)]/tots:5.1f}%")
