#!/usr/bin/env python
"""Key figures of every launch in an .ncu-rep (read here, no GPU): duration, registers, pipe utilisation, issue slots,
warp stall breakdown, DRAM bytes.  usage: ncu_summary.py report.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
stall2 = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") or (h.startswith("smsp__average_warps_issue_stalled") and "ratio" in h)]
out = []
for r in rows[2:]:
    d = {w: r[idx[w]] for w in want if w in idx}
    st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[idx[h]]) for h in stall if r[idx[h]]}
    d["stall_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
    out.append(d)
units_d = {w: units[idx[w]] for w in want if w in idx}
res = {"_report": rep, "_units": units_d, "launches": out}
s = json.dumps(res, indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(s)
print(s)
