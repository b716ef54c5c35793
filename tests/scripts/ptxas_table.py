#!/usr/bin/env python
"""Registers / spills of every kernel of the library: python tests/scripts/ptxas_table.py (forces a verbose rebuild)."""
import re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
res = subprocess.run([sys.executable, '-m', 'xicsrt_b200.build', '--force', '--verbose'] , capture_output=True, text=True)
txt = res.stderr
name = None
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void xrt::', '')
    m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', ln)
    if m: spill = m.groups()
    m = re.search(r'Used (\d+) registers', ln)
    if m and name:
        print(f'{name:60s} regs {m.group(1):>4s} stack {spill[0]:>4s} spill st/ld {spill[1]:>4s}/{spill[2]:>4s}')
        name = None
if res.returncode: print(txt[-3000:])
