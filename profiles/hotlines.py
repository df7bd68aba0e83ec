#!/usr/bin/env python
"""Joins an `ncu --page source --csv` SASS export with `nvdisasm --print-line-info` of the same cubin and prints
per-source-line executed-instruction counts and stall samples for one kernel.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm --print-line-info X.cubin > X.sass
    python profiles/hotlines.py sass.csv X.sass <kernel-substring> [top]
"""
import csv, re, sys
from collections import defaultdict

csv_path, sass_path, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
insts = [(r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)) for r in rows[2:] if len(r) > 5]

# walk nvdisasm output of the wanted function: track current line marker, collect per instruction
lines, cur, infunc = [], None, False
for ln in open(sass_path):
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infunc = kernel in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', ln)
        cur = (m.group(1).split("/")[-1], int(m.group(2)), (inl.group(1).split("/")[-1], int(inl.group(2))) if inl else None)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append(cur)
if len(lines) != len(insts):
    print(f"warning: {len(lines)} disassembled vs {len(insts)} profiled instructions", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for (src, n, s), loc in zip(insts, lines):
    key = loc[:2] if loc else ("?", 0)
    agg[key][0] += n; agg[key][1] += s
    tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]}:{key[1]:<5d} inst {n:>12d} ({100*n/max(tot_i,1):5.1f}%)  samples {s:>7d} ({100*s/max(tot_s,1):5.1f}%)")
