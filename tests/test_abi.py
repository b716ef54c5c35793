# -*- coding: utf-8 -*-
"""
The C ABI without a GPU: libxrt.so loads, exports every function include/xrt.h
declares, and the ctypes mirror of the structs has the C compiler's layout.
No compute entry point is called here (there is no device in the build container).
"""
import ctypes as C
import os
import re
import subprocess

import pytest

from xicsrt_b200 import _lib as L
from xicsrt_b200 import build as xbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'xrt.h')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(xrt_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = xbuild.build()
    lib = C.CDLL(path)
    names = declared_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), f'{name} is declared in include/xrt.h but not exported'
    assert sorted(L.SYMBOLS) == names, 'the ctypes binding and the header disagree on the entry points'


def test_version_and_error_string_need_no_device():
    lib = L.load()
    assert lib.xrt_version() == L.XRT_VERSION
    assert isinstance(lib.xrt_last_error(), bytes)


def test_no_cpu_fallback_scene_create_fails_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a device is present')
    lib = L.load()
    desc = L.XrtSceneDesc()
    desc.version = L.XRT_VERSION
    handle = C.c_void_p()
    rc = lib.xrt_scene_create(C.byref(desc), C.byref(handle))
    assert rc == L.ECUDA
    assert b'no CUDA device' in lib.xrt_last_error()


PROBE = r'''
#include <stdio.h>
#include <stddef.h>
#include "xrt.h"
#define S(T) printf(#T " %zu\n", sizeof(T))
#define O(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
    S(XrtAperture); S(XrtMesh); S(XrtOpticDesc); S(XrtSightline); S(XrtBundle); S(XrtSourceDesc); S(XrtPlasmaDesc);
    S(XrtSceneDesc); S(XrtOutputs); S(XrtHistory); S(XrtRaysIn); S(XrtInject); S(XrtSourceInject);
    O(XrtOpticDesc, origin); O(XrtOpticDesc, center); O(XrtOpticDesc, root_idx); O(XrtOpticDesc, two_d);
    O(XrtOpticDesc, n_aperture); O(XrtOpticDesc, apertures); O(XrtOpticDesc, mesh); O(XrtOpticDesc, npix);
    O(XrtOpticDesc, image_offset); O(XrtOpticDesc, cull_t2); O(XrtOpticDesc, cull_inv_r); O(XrtOpticDesc, mosaic_scan); O(XrtOpticDesc, mosaic_err);
    O(XrtSourceDesc, axis_basis); O(XrtSourceDesc, cone_par); O(XrtSourceDesc, wave_par); O(XrtSourceDesc, n_table);
    O(XrtSourceDesc, table_cdf); O(XrtSourceDesc, sightlines); O(XrtSourceDesc, n_bundles); O(XrtSourceDesc, voxel_size); O(XrtSourceDesc, bundle_x); O(XrtSourceDesc, bundle_hint_shift);
    O(XrtPlasmaDesc, cone); O(XrtPlasmaDesc, origin); O(XrtPlasmaDesc, inject_u); O(XrtPlasmaDesc, sightlines);
    O(XrtSceneDesc, source); O(XrtSceneDesc, optics); O(XrtSceneDesc, kn32);
    O(XrtMesh, n_tri); O(XrtMesh, grid_nx); O(XrtMesh, grid_x0); O(XrtMesh, vgrid_items);
    O(XrtOutputs, found_capacity); O(XrtOutputs, lost_threshold);
    return 0;
}
'''


def test_ctypes_structs_match_the_c_layout(tmp_path):
    src = tmp_path / 'probe.c'
    src.write_text(PROBE)
    exe = tmp_path / 'probe'
    subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        what, value = line.split()
        value = int(value)
        if '.' in what:
            struct, field = what.split('.')
            assert getattr(getattr(L, struct), field).offset == value, what
        else:
            assert C.sizeof(getattr(L, what)) == value, what


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without libxrt.so the binding raises instead of computing something else."""
    monkeypatch.setattr(L, '_lib', None)
    monkeypatch.setattr(L, 'LIB_PATH', str(tmp_path / 'libxrt.so'))
    with pytest.raises(ImportError, match='no CPU fallback'):
        L.load()


def test_public_api_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a device is present')
    import xicsrt_b200
    from oracle import scenes
    with pytest.raises(RuntimeError, match='no CPU path'):
        xicsrt_b200.raytrace(scenes.get('sphere'))
