#!/usr/bin/env python
"""Innermost-source-line attribution of executed instructions: joins nvdisasm -g line info of the
built library with the per-address execution counts of an ncu report.
usage: python profiles/ncu_hot.py report.ncu-rep libxrt.so <kernel-substring> [rays_per_launch] [top]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib, kname = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
rays = float(sys.argv[4]) if len(sys.argv) > 4 else None
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
    # one cubin per translation unit: take the one that holds the kernel
    dis = ''
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith('.cubin')):
        txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        if kname in txt:
            dis = txt
            break

line_of, text_of, cur, inside = {}, {}, ('?', 0), False
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
        off = int(m.group(1), 16)
        line_of[off] = cur
        text_of[off] = m.group(2).strip()

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, ie, isrc = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Source')
addr = [(int(r[ia], 16), int(r[ie]), r[isrc]) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
base = addr[0][0]
acc = defaultdict(lambda: [0, 0])
tot = 0
for a, n, s in addr:
    key = line_of.get(a - base, ('?', 0))
    acc[key][0] += n
    op = [o for o in s.strip().split() if not o.startswith('@')]
    if op and op[0].split('.')[0] in ('DFMA', 'DMUL', 'DADD', 'DSETP'):
        acc[key][1] += n
    tot += n
scale = rays / 32 if rays else 1.0
srcs = {}
def text(f, l):
    for d in ('xicsrt_b200/csrc', 'include'):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in srcs:
                srcs[p] = open(p).read().splitlines()
            return srcs[p][l - 1].strip()[:95] if 0 < l <= len(srcs[p]) else ''
    return ''
print(f'# {kname}: {tot} warp instructions' + (f' = {tot / scale:.1f} per 32 rays' if rays else ''))
print('percent,per_32_rays,fp64_per_32_rays,file:line,source')
for (f, l), (n, d) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * n / tot:5.2f},{n / scale:7.1f},{d / scale:6.1f},{f}:{l},"{text(f, l)}"')
