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
# one translation unit per compiled feature set (they build in parallel) + the host side
SOURCES = ['xrt.cu', 'v_cull.cu', 'v_lean.cu', 'v_mid.cu', 'v_mosaic.cu', 'v_src.cu', 'v_mesh.cu', 'v_full.cu']
HEADERS = ['xrt_math.cuh', 'xrt_fastmath.cuh', 'xrt_trace.cuh', 'xrt_mesh.cuh', 'xrt_plasma.cuh', 'xrt_kernels.cuh',
           'xrt_meshsort.cuh', 'xrt_select.cuh', 'xrt_variants.h', os.path.join('..', '..', 'include', 'xrt.h')]
OBJ_DIR = os.path.join(os.path.dirname(HERE), 'build', 'obj')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']


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


def _compile(nvcc, src, obj, verbose, extra):
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res


def build(force=False, verbose=False, extra_flags=(), lib=None):
    """Compile libxrt.so if it is missing or older than its sources.  Returns the path."""
    lib = lib or LIB
    if not force and lib == LIB and not is_stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError('nvcc not found: cannot build xicsrt_b200/libxrt.so')
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = OBJ_DIR if lib == LIB and not extra_flags else OBJ_DIR + '_' + os.path.basename(lib)
    os.makedirs(obj_dir, exist_ok=True)
    header_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    jobs, objs = [], []
    for f in SOURCES:
        src = os.path.join(CSRC, f)
        obj = os.path.join(obj_dir, os.path.splitext(f)[0] + '.o')
        objs.append(obj)
        fresh = os.path.exists(obj) and os.path.getmtime(obj) > max(header_time, os.path.getmtime(src))
        if force or not fresh:
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        results = list(pool.map(lambda j: _compile(nvcc, j[0], j[1], verbose, extra_flags), jobs))
    for src, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(f'==== {os.path.basename(src)}\n' + res.stdout + res.stderr)
    if any(res.returncode != 0 for _, res in results):
        raise RuntimeError('nvcc failed building libxrt.so')
    res = subprocess.run([nvcc, '--shared', '-o', lib + '.tmp'] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed linking libxrt.so')
    os.replace(lib + '.tmp', lib)
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
