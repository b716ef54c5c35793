#!/usr/bin/env python
"""Warp-stall samples per innermost source line (ncu pc sampling joined with nvdisasm -g line info).
usage: python profiles/ncu_stalls.py report.ncu-rep libxrt.so <kernel-substring> [top]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib, kname = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
    # one cubin per translation unit: take the one that holds the kernel
    dis = ''
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith('.cubin')):
        txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        if kname in txt:
            dis = txt
            break
line_of, cur, inside = {}, ('?', 0), False
for ln in dis.splitlines():
    if ln.startswith('\t.section\t.text.'):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, ie, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
names = ['stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_math', 'stall_dispatch', 'stall_branch_resolving',
         'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_barrier', 'stall_lg', 'stall_mio']
cols = {k: hdr.index(k) for k in names if k in hdr}
acc, tot, base, launches = defaultdict(lambda: defaultdict(int)), 0, None, 0
for r in rows[2:]:
    if r and r[0] == 'Address':
        launches += 1
        continue
    if launches or len(r) <= ie or not r[ie].isdigit():
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    key = line_of.get(a - base, ('?', 0))
    s = int(r[isamp])
    acc[key]['samples'] += s
    tot += s
    for k, i in cols.items():
        acc[key][k] += int(r[i])
agg = defaultdict(int)
for d in acc.values():
    for k in cols:
        agg[k] += d[k]
print('# total samples', tot, {k[6:]: round(100 * v / tot, 1) for k, v in agg.items()})
print('percent_of_samples,file:line,top stall reasons')
for key, d in sorted(acc.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    best = sorted(((d[k], k) for k in cols), reverse=True)[:3]
    print(f"{100 * d['samples'] / tot:5.2f},{key[0]}:{key[1]}," + ' '.join(f"{k[6:]}={100 * v / max(d['samples'], 1):.0f}%" for v, k in best))
