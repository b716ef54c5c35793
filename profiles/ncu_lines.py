#!/usr/bin/env python
"""Per-source-line dynamic instruction counts from an ncu report (needs -lineinfo).
usage: python profiles/ncu_lines.py report.ncu-rep [rays_per_launch] [top]"""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
rays = float(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
path, hdr, funcs = None, None, 0
acc = defaultdict(lambda: [0, 0, 0, ''])   # (file, line) -> [instr, fp64 instr, samples, text]
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        path = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        funcs += 1; continue
    if r[0] == 'Line No':
        hdr = r
        il, isrc, isass = 0, 1, 3
        ie, iss = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= ie:
        continue
    try:
        n, s = int(r[ie]), int(r[iss])
    except ValueError:
        continue
    key = (path, r[il])
    a = acc[key]
    a[0] += n; a[2] += s
    op = [o for o in r[isass].strip().split() if not o.startswith('@')]
    if op and op[0].split('.')[0] in ('DFMA', 'DMUL', 'DADD', 'DSETP'):
        a[1] += n
    if r[isrc].strip():
        a[3] = r[isrc].strip()[:100]
tot = sum(a[0] for a in acc.values())
scale = (rays / 32) if rays else 1.0
print(f'# total warp instructions {tot}' + (f' = {tot / scale:.1f} per 32 rays' if rays else ''))
print('percent,per_32_rays,fp64_per_32_rays,samples,file:line,source')
for (f, l), a in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * a[0] / tot:5.2f},{a[0] / scale:7.1f},{a[1] / scale:7.1f},{a[2]},{f}:{l},"{a[3]}"')
