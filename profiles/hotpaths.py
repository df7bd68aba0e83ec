#!/usr/bin/env python
"""Like hotlines.py, but with inlining context: joins an `ncu --page source --csv` SASS export with
`nvdisasm --print-line-info-inline` of the same cubin and aggregates executed warp-instructions by the call-site
chain (outermost first) truncated at a given depth.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm --print-line-info-inline X.cubin > X_inl.sass
    python profiles/hotpaths.py sass.csv X_inl.sass <kernel-substring> <depth> [top]
"""
import csv, re, sys
from collections import defaultdict

csv_path, sass_path, kernel, depth = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
insts = [(r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)) for r in rows[2:] if len(r) > 5]

chains, cur, pending, infunc = [], (), [], False
for ln in open(sass_path):
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infunc = kernel in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        if pending:
            cur = tuple(reversed(pending))  # outermost first
            pending = []
        chains.append(cur)
if len(chains) != len(insts):
    print(f"warning: {len(chains)} disassembled vs {len(insts)} profiled instructions", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for (src, n, s), ch in zip(insts, chains):
    key = ch[:depth]
    agg[key][0] += n; agg[key][1] += s
    tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    name = " > ".join(f"{f.replace('lighting.cu','L').replace('particles.cu','P').replace('ilb_device.cuh','D')}:{l}" for f, l in key)
    print(f"inst {n:>12d} ({100*n/max(tot_i,1):5.1f}%)  samples {s:>7d} ({100*s/max(tot_s,1):5.1f}%)  {name}")
