#!/usr/bin/env python
"""Digest of an ncu report: key metrics + dynamic instruction mix per opcode.
usage: python profiles/ncu_digest.py report.ncu-rep [rays_per_launch]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
rays = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
keep = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__average_warps_issue_stalled']
print('metric,unit,' + ','.join(f'launch{i}' for i in range(len(vals))))
for i, h in enumerate(hdr):
    if any(h.startswith(k) for k in keep) and '.sum.' not in h:
        print(f'{h},{units[i]},' + ','.join(v[i] for v in vals))

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
byop, bythr, tot, launches = Counter(), Counter(), 0, 0
for r in rows:
    if r and r[0] == 'Address':
        h = r; launches += 1
        ia, ie, it = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')
        continue
    if h is None or len(r) <= it or launches > 1:
        continue
    try:
        n, t = int(r[ie]), int(r[it])
    except ValueError:
        continue
    op = [o for o in r[ia].strip().split() if not o.startswith('@')]
    name = op[0].split('.')[0] if op else '?'
    byop[name] += n; bythr[name] += t; tot += n
print(f'\n# dynamic warp instructions (launch 0): {tot}' + (f' = {tot / (rays / 32):.1f} per 32 rays' if rays else ''))
print('opcode,warp_instructions,percent,' + ('per_32_rays,' if rays else '') + 'avg_active_threads')
for k, v in byop.most_common(28):
    print(f'{k},{v},{100 * v / tot:.2f},' + (f'{v / (rays / 32):.1f},' if rays else '') + f'{bythr[k] / v:.1f}')
