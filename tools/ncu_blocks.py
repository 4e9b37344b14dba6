#!/usr/bin/env python
"""Where the time of one profiled kernel goes, per basic block: reads the source page of an `ncu --set full
--import-source on` report and prints, for every block that matters, its static size, FP64-pipe instructions, dynamic
instructions per warp-tile, share of the warp-state samples and the split of those samples over the stall reasons.
usage: ncu_blocks.py report.ncu-rep rays_per_launch rays_per_warp_tile      (e.g. 83868160 96 for 3 rays/thread)"""
import csv
import re
import subprocess
import sys

rep, rays, per_warp = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
print("#", rows[0][1] if len(rows[0]) > 1 else "")
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
addr = [int(r[ix['Address']], 16) for r in data]
base = addr[0]
targets = set()
for r in data:
    m = re.search(r'\b(?:BRA|BSSY|CALL)\S*\s.*?(0x[0-9a-f]+)', r[ix['Source']])
    if m:
        targets.add(int(m.group(1), 16))
blocks, cur = [], []
for a, r in zip(addr, data):
    if (a - base) in targets and cur:
        blocks.append(cur); cur = []
    cur.append(r)
    if re.search(r'\b(BRA|EXIT|RET)\b', r[ix['Source']]):
        blocks.append(cur); cur = []
if cur:
    blocks.append(cur)
T = rays / per_warp                      # warp-tiles per launch
st = ['stall_wait', 'stall_math', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_dispatch', 'stall_no_inst',
      'stall_branch_resolving']
print(f"# {tot} samples, {len(data)} instructions, {T:.0f} warp-tiles; columns: offset, static instructions, of which FP64 pipe, "
      f"executions per warp-tile, dynamic instructions per warp-tile, % of samples | % of samples by stall reason")
print("off      n fp64  exec  instr  samp% | " + " ".join(f"{s[6:14]:>8s}" for s in st))
ti = fi = 0.0
for b in blocks:
    ex = [int(r[ix['Instructions Executed']] or 0) for r in b]
    s = sum(int(r[ix['# Samples']] or 0) for r in b)
    ni = sum(ex) / T
    ops = [re.sub(r'^@!?U?P\d+\s+', '', r[ix['Source']]).split()[0].split('.')[0] for r in b]
    isf = [o in ('DFMA', 'DMUL', 'DADD', 'DSETP') for o in ops]
    ti += ni; fi += sum(e for e, f in zip(ex, isf) if f) / T
    if ni < 1 and s < 0.005 * tot:
        continue
    print("%05x %4d %4d %5.2f %6.1f %6.2f | " % (int(b[0][ix['Address']], 16) - base, len(b), sum(isf), max(ex) / T, ni, 100 * s / tot) +
          " ".join("%8.2f" % (100 * sum(int(r[ix[k]] or 0) for r in b) / tot) for k in st))
print(f"# dynamic instructions per warp-tile: {ti:.1f}, of which FP64 pipe {fi:.1f}")
