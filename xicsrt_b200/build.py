# -*- coding: utf-8 -*-
"""
Builds ``xicsrt_b200/libxrt.so`` (the C-ABI library of include/xrt.h) in-tree
with nvcc for sm_100a.  The library is git-ignored but travels to the GPU box
with the working tree, so nothing is compiled there.

    python -m xicsrt_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libxrt.so')
SOURCES = ['xrt.cu']
HEADERS = ['xrt_math.cuh', 'xrt_fastmath.cuh', 'xrt_trace.cuh', 'xrt_mesh.cuh', 'xrt_plasma.cuh', os.path.join('..', '..', 'include', 'xrt.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--shared', '-Xcompiler', '-fPIC']


def nvcc_path():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile libxrt.so if it is missing or older than its sources.  Returns the path."""
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError('nvcc not found: cannot build xicsrt_b200/libxrt.so')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
          ['-o', LIB + '.tmp'] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libxrt.so')
    os.replace(LIB + '.tmp', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
