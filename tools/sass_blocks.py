#!/usr/bin/env python
"""Basic-block statistics of a `cuobjdump -sass` listing: per block the FP64-pipe instructions (DFMA, DMUL, DADD, DSETP),
the MUFU seeds and everything else -- the static counterpart of the dynamic mix in profiles/*_ncu_summary.json.
usage: sass_blocks.py listing.sass [min_instructions]"""
import re
import sys

ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
targets = set()
for a, t in ins:
    m = re.search(r'\b(?:BRA|BSSY|CALL)\S*\s.*?(0x[0-9a-f]+)', t)
    if m:
        targets.add(int(m.group(1), 16))
minn = int(sys.argv[2]) if len(sys.argv) > 2 else 8
blk = []
def flush():
    if len(blk) >= minn:
        ops = [re.sub(r'^@!?U?P\d+\s+', '', t).split()[0] for _, t in blk]
        base = [o.split('.')[0] for o in ops]
        f = sum(o in ('DFMA', 'DMUL', 'DADD', 'DSETP') for o in base)
        mu = [o.split('.')[1] for o in ops if o.startswith('MUFU')]
        print('%05x-%05x n=%3d fp64=%3d other=%3d mufu=%s end=%s' % (blk[0][0], blk[-1][0], len(blk), f, len(blk) - f, ','.join(mu), blk[-1][1][:40]))
for a, t in ins:
    if a in targets:
        flush(); blk = []
    blk.append((a, t))
    if re.search(r'\b(BRA|EXIT|RET|BRX)\b', t):
        flush(); blk = []
flush()
