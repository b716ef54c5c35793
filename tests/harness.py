# -*- coding: utf-8 -*-
"""
Shared plumbing of the GPU parity tests: run the oracle with every random draw
recorded, push the very same rays and draws through the CUDA path
(``xrt_trace_injected`` / ``xrt_source_injected``, called through the C ABI via
ctypes) and hand both results back for comparison.
"""
import copy
import ctypes as C

import numpy as np

import oracle
from xicsrt_b200 import _lib as L
from xicsrt_b200 import config as xconfig
from xicsrt_b200 import scene as xscene


def device_scene(config, poisson=None):
    """Product-side preparation of a user config -> (DeviceScene, layout, optic params)."""
    cfg = xconfig.get_config(xconfig.to_numpy(copy.deepcopy(config)))
    _, sname, sparam, sfilters, optics = xscene.prepare(cfg, poisson=poisson)
    desc, layout, keep = xscene.flatten(sname, sparam, sfilters, optics)
    scene = xscene.DeviceScene(desc, layout)
    return scene, layout, sparam, optics


def draws_for_optics(stream, optics, n):
    """Recorded draws scattered to full length: ({k: u[depth,n]}, {k: xy[depth,2,n]})."""
    u, xy = {}, {}
    for k, (name, param) in enumerate(optics.items()):
        kind = param['_interact']
        if kind not in ('crystal', 'mosaic'):
            continue
        depth = int(param['mosaic_depth']) if kind == 'mosaic' else 1
        sites = {s for s, _, _, _ in stream.log}
        if any(s.startswith(f'opt.{k}.u.') for s in sites):
            u[k] = np.stack([stream.scattered(f'opt.{k}.u.{layer}', n) for layer in range(depth)])
        if kind == 'mosaic':
            xy[k] = np.stack([stream.scattered(f'opt.{k}.xy.{layer}', n, width=2).T for layer in range(depth)])
    return u, xy


def run_injected(torch, scene, layout, rays0, u, xy, want_history=True):
    """
    rays0: dict origin (n,3), direction (n,3), wavelength (n,), mask (n,) -- the source rays.
    Returns (history {name: rays}, counts {name: int}, images {name: array or None}).
    """
    dev = torch.device('cuda', 0)
    n = len(rays0['mask'])
    n_elem = 1 + len(layout.optic_names)

    def up(a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(dev)

    t_o, t_d = up(rays0['origin'], np.float64), up(rays0['direction'], np.float64)
    t_w, t_m = up(rays0['wavelength'], np.float64), up(rays0['mask'], np.uint8)
    rin = L.XrtRaysIn()
    rin.origin, rin.direction, rin.wavelength, rin.mask = t_o.data_ptr(), t_d.data_ptr(), t_w.data_ptr(), t_m.data_ptr()

    inj = L.XrtInject()
    held = []
    for k, a in u.items():
        t = up(a, np.float64)
        held.append(t)
        inj.u[k] = t.data_ptr()
    for k, a in xy.items():
        t = up(a, np.float64)
        held.append(t)
        inj.xy[k] = t.data_ptr()

    packed = torch.zeros(n_elem + max(layout.n_pixels, 1), dtype=torch.int64, device=dev)
    out = L.XrtOutputs()
    out.counts = packed.data_ptr()
    out.images = packed.data_ptr() + 8 * n_elem

    hist_rays = torch.full((n_elem, 7, max(n, 1)), -7.0, dtype=torch.float64, device=dev)
    hist_mask = torch.full((n_elem, max(n, 1)), 9, dtype=torch.uint8, device=dev)
    h = L.XrtHistory()
    h.rays, h.mask, h.capacity = hist_rays.data_ptr(), hist_mask.data_ptr(), max(n, 1)

    lib = L.load()
    L.check(lib.xrt_trace_injected(scene.handle, C.byref(rin), C.byref(inj), n, C.byref(out),
                                   C.byref(h) if want_history else None, None))
    torch.cuda.synchronize()

    host = packed.cpu().numpy()
    names = layout.element_names
    counts = {name: int(host[i]) for i, name in enumerate(names)}
    images = {}
    for name in layout.optic_names:
        spec = layout.images[name]
        if spec is None:
            images[name] = None
        else:
            off, nx, ny = spec
            images[name] = host[n_elem + off:n_elem + off + nx * ny].astype(np.float64).reshape(nx, ny)
    history = {}
    if want_history:
        r = hist_rays.cpu().numpy()
        m = hist_mask.cpu().numpy()
        for e, name in enumerate(names):
            history[name] = {'origin': r[e, 0:3, :n].T, 'direction': r[e, 3:6, :n].T,
                             'wavelength': r[e, 6, :n], 'mask': m[e, :n].astype(bool)}
    return history, counts, images


def oracle_and_cuda(torch, config):
    """Oracle iteration with recorded draws + the CUDA path on the same rays and draws."""
    single, stream, oscene = oracle.trace_recorded(config)
    scene, layout, sparam, optics = device_scene(config, poisson=lambda lam: int(single['meta'][oscene.source_name]['num_out']))
    rays0 = single['history'][oscene.source_name]
    n = len(rays0['mask'])
    u, xy = draws_for_optics(stream, optics, n)
    hist, counts, images = run_injected(torch, scene, layout, rays0, u, xy)
    scene.close()
    return single, hist, counts, images, layout


def assert_rays_close(got, ref, what, rtol=1e-9):
    """
    The bar of BASELINE.json: masks bit-equal; positions / directions / wavelengths within
    1e-9 relative (scaled by the largest component of the row, so that a component that is
    ~0 by cancellation is judged against the vector it belongs to).
    """
    gm, rm = np.asarray(got['mask'], bool), np.asarray(ref['mask'], bool)
    bad = np.flatnonzero(gm != rm)
    assert bad.size == 0, f'{what}: {bad.size} mask mismatches, first at ray {bad[:5]}'
    for key in ('origin', 'direction', 'wavelength'):
        a, b = np.asarray(got[key]), np.asarray(ref[key])
        assert a.shape == b.shape, f'{what}: {key} shape'
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        assert np.array_equal(nan_a, nan_b), f'{what}: {key} NaN pattern differs on {np.sum(nan_a != nan_b)} entries'
        if a.ndim == 2:
            scale = np.max(np.where(nan_b, 0.0, np.abs(b)), axis=1, keepdims=True)
        else:
            scale = np.abs(b)
        scale = np.where(np.isfinite(scale) & (scale > 0), scale, 1.0)
        err = np.abs(a - b) / scale
        err = np.where(nan_b, 0.0, err)
        worst = np.max(err) if err.size else 0.0
        assert worst <= rtol, f'{what}: {key} relative error {worst:.3e} > {rtol}'
