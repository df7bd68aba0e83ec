#!/usr/bin/env python
"""Summarises an `ncu --set full` report (every kernel launch in it) into a short text file and updates profiles/traffic.json
(the sum over the report's launches under <kernel-key>).

    python profiles/summarize.py gpurun_out/prof_X.ncu-rep <kernel-key> profiles/<name>.txt
"""
import csv, io, json, subprocess, sys
from pathlib import Path

rep, key, out = sys.argv[1], sys.argv[2], Path(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]
lines = [f"# ncu --set full --clock-control none, report {Path(rep).name}"]
traffic = 0.0
for vals in rows[2:]:
    if len(vals) < len(hdr) // 2:
        continue
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    for w in want:
        if w in m:
            lines.append(f"{w:75s} {m[w][0]} {m[w][1]}")
    def mb(name):
        v, u = m.get(name, ("0", "byte"))
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    t = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    traffic += t
    lines.append(f"traffic (dram read + write) of this launch: {t/1e6:.1f} MB")
    stalls = sorted(((float(v[0].replace(',', '')), h) for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled")), reverse=True)[:8]
    for v, h in stalls:
        lines.append(f"{h:75s} {v:.3f}")
    lines.append("")
lines.append(f"traffic (dram read + write), all launches above: {traffic/1e6:.1f} MB")
out.write_text("\n".join(lines) + "\n")
tj = out.parent / "traffic.json"
d = json.loads(tj.read_text()) if tj.exists() else {}
d[key] = traffic
tj.write_text(json.dumps(d, indent=1) + "\n")
print("\n".join(lines))
