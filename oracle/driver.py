# -*- coding: utf-8 -*-
"""
oracle.driver -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Run / iteration loops, found-lost sorting and result combination, restating
``xicsrt/xicsrt_raytrace.py:28-393`` and ``xicsrt/objects/_Dispatcher.py:142-196``
so that ``oracle.raytrace(config)`` returns the same dictionary (same keys,
same arrays) as ``xicsrt.raytrace(config)`` for the same ``random_seed``.
"""
import copy
import multiprocessing

import numpy as np

from xicsrt_b200 import config as xconfig
from xicsrt_b200 import elements

from oracle import optics, sources
from oracle.stream import LegacyStream

RAY_KEYS = ('origin', 'direction', 'mask', 'wavelength')


def _copy_rays(rays):
    return {k: np.array(v, copy=True) for k, v in rays.items() if not k.startswith('_')}


class Scene:
    """Prepared filters / source / optics of one run (order = config order)."""

    def __init__(self, config, stream):
        strict = config['general']['strict_config_check']
        self.filters = {}
        cfg_filters = {}
        for name, c in (config.get('filters') or {}).items():
            cfg_filters[name], self.filters[name] = elements.prepare_filter(c, strict=strict)
        if 'filters' in config:
            config['filters'] = cfg_filters

        if len(config['sources']) == 0:
            raise Exception('No ray sources defined.')
        if len(config['sources']) != 1:
            raise NotImplementedError('Multiple ray sources are not currently supported.')
        cfg_sources = {}
        for name, c in config['sources'].items():
            cfg_sources[name], self.source = elements.prepare_source(
                c, strict=strict, poisson=lambda lam: stream.poisson(lam, site='src.count'))
            self.source_name = name
        config['sources'] = cfg_sources
        self.source_filters = [self.filters[f] for f in self.filters
                               if self.source.get('filters') is not None and f in self.source['filters']]

        self.optics = {}
        cfg_optics = {}
        for name, c in config['optics'].items():
            cfg_optics[name], self.optics[name] = elements.prepare_optic(c, strict=strict)
        config['optics'] = cfg_optics
        self.config = config


def generate(scene, stream):
    if scene.source['_kind'].startswith('plasma'):
        rays, _ = sources.plasma_source(scene.source, stream, scene.source_filters)
    else:
        rays = sources.box_source(scene.source, stream, scene.source_filters)
    return rays


def run_iteration(scene, stream, keep_history=True, keep_images=True, keep_meta=True):
    """xicsrt_raytrace.py:178-226 with _Dispatcher.py:142-196 inlined."""
    meta, image, history = {}, {}, {}
    rays = generate(scene, stream)
    if keep_meta:
        meta[scene.source_name] = {'num_out': np.sum(rays['mask'])}
    if keep_history:
        history[scene.source_name] = _copy_rays(rays)
    for k, (name, param) in enumerate(scene.optics.items()):
        rays = optics.trace_optic(param, rays, stream, f'opt.{k}')
        if keep_meta:
            meta[name] = {'num_out': np.sum(rays['mask'])}
        if keep_history:
            history[name] = _copy_rays(rays)
        if keep_images:
            image[name] = optics.bin_image(param, rays['origin'], rays['mask'])
    return {'config': scene.config, 'meta': meta, 'image': image, 'history': history}


def _skeleton(config):
    return {'config': config,
            'total': {'meta': {}, 'image': {}},
            'found': {'meta': {}, 'history': {}},
            'lost': {'meta': {}, 'history': {}}}


def sort_iteration(single, stream, max_lost=None):
    """xicsrt_raytrace.py:229-278."""
    if max_lost is None:
        max_lost = 1000
    out = _skeleton(single['config'])
    out['total']['meta'] = single['meta']
    out['total']['image'] = single['image']
    hist = single['history']
    if len(hist) > 0:
        last = list(hist.keys())[-1]
        w_found = np.flatnonzero(hist[last]['mask'])
        w_lost = np.flatnonzero(np.invert(hist[last]['mask']))
        max_lost = min(max_lost, len(w_lost))
        order = np.arange(len(w_lost))
        stream.shuffle(order)
        w_lost = w_lost[order[:max_lost]]
        for name in hist:
            out['found']['history'][name] = {k: v[w_found] for k, v in hist[name].items()}
            out['lost']['history'][name] = {k: v[w_lost] for k, v in hist[name].items()}
    return out


def combine(parts):
    """xicsrt_raytrace.py:281-393 (keep_images=True, all components)."""
    out = _skeleton(parts[0]['config'])
    names = list(parts[0]['total']['meta'].keys())
    last = names[-1]

    for name in names:
        out['total']['meta'][name] = {}
        for key in parts[0]['total']['meta'][name]:
            out['total']['meta'][name][key] = 0
            for p in parts:
                out['total']['meta'][name][key] += p['total']['meta'][name][key]

    for name in names:
        if name in parts[0]['total']['image']:
            first = parts[0]['total']['image'][name]
            if first is None:
                out['total']['image'][name] = None
            elif all(p['total']['image'][name].shape == first.shape for p in parts):
                acc = np.zeros(first.shape)
                for p in parts:
                    acc += p['total']['image'][name]
                out['total']['image'][name] = acc
            else:
                out['total']['image'][name] = None

    if len(parts[0]['found']['history']) > 0:
        for kind in ('found', 'lost'):
            for name in names:
                out[kind]['history'][name] = {
                    key: np.concatenate([p[kind]['history'][name][key] for p in parts])
                    for key in RAY_KEYS}
    return out


def raytrace_single(config, _internal=False, stream=None):
    """xicsrt_raytrace.py:87-175."""
    config = xconfig.to_numpy(config)
    config = xconfig.get_config(config)
    seed = config['general']['random_seed']
    if stream is None:
        stream = LegacyStream(seed)

    num_iter = config['general']['number_of_iter']
    max_lost_iter = int(config['general']['history_max_lost'] / num_iter)
    if _internal:
        max_lost_iter = max_lost_iter // config['general']['number_of_runs']
    max_lost_iter = max(int(max_lost_iter), 1)

    scene = Scene(config, stream)
    g = config['general']
    parts = []
    for _ in range(num_iter):
        single = run_iteration(scene, stream, keep_history=g['keep_history'],
                               keep_images=g['keep_images'], keep_meta=g['keep_meta'])
        parts.append(sort_iteration(single, stream, max_lost=max_lost_iter))
    return combine(parts)


def _run_configs(config):
    config = xconfig.get_config(config)
    seed = config['general']['random_seed']
    runs = []
    for ii in range(config['general']['number_of_runs']):
        c = copy.deepcopy(config)
        c['general']['output_run_suffix'] = '{:04d}'.format(ii)
        if seed is not None:
            seed += ii          # cumulative, as in xicsrt_raytrace.py:61-63
        c['general']['random_seed'] = seed
        runs.append(c)
    return config, runs


def raytrace(config):
    """xicsrt_raytrace.py:28-84 (file saving and printing left out)."""
    config, runs = _run_configs(config)
    out = combine([raytrace_single(c, _internal=True) for c in runs])
    out['config']['general']['output_run_suffix'] = config['general']['output_run_suffix']
    out['config']['general']['random_seed'] = config['general']['random_seed']
    return out


def _mp_worker(c):
    return raytrace_single(c, _internal=True)


def raytrace_mp(config, processes=None):
    """xicsrt_multiprocessing.py:12-81 -- one pool task per run."""
    config, runs = _run_configs(config)
    with multiprocessing.Pool(processes) as pool:
        results = [pool.apply_async(_mp_worker, (c,)) for c in runs]
        pool.close()
        pool.join()
    out = combine([r.get() for r in results])
    out['config']['general']['output_run_suffix'] = config['general']['output_run_suffix']
    out['config']['general']['random_seed'] = config['general']['random_seed']
    return out


def trace_recorded(config, seed=None):
    """
    One unsorted iteration with every random draw recorded (for injection into
    the CUDA path).  Returns (single, stream, scene): ``single['history']``
    holds full-length per-element ray arrays in ray order; ``stream.scattered``
    turns a draw site into a full-length array.
    """
    config = xconfig.to_numpy(copy.deepcopy(config))
    config = xconfig.get_config(config)
    if seed is None:
        seed = config['general']['random_seed']
    stream = LegacyStream(seed, record=True)
    scene = Scene(config, stream)
    single = run_iteration(scene, stream, keep_history=True, keep_images=True)
    return single, stream, scene
