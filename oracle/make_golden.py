# -*- coding: utf-8 -*-
"""
oracle.make_golden -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Generates ``tests/golden/<scene>.npz`` by running the UNMODIFIED reference
(imported from /root/reference, which exists only in the build container) on
the scenes of oracle/scenes.py.  Nothing at test / bench / smoke time imports
the reference; only these fixtures travel.

Two kinds of record per scene:

  iter/<element>/<key>   full-length, unsorted per-element ray arrays of one
                         iteration, obtained exactly the way
                         ``xicsrt_raytrace._raytrace_iter`` (:178-226) does it:
                         seed, build the three Dispatchers, generate_rays,
                         trace with keep_history / keep_images.
  iter_image/<element>, iter_meta/<element>
  run/...                (selected scenes) the complete ``xicsrt.raytrace``
                         output: total meta+image, found and lost histories.

Usage:  PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden [scene ...]
"""
import os
import sys

import numpy as np

REFERENCE = '/root/reference'
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

RUN_SCENES = ('two_iter_two_runs', 'sphere_step_box')


def _import_reference():
    if not os.path.isdir(REFERENCE):
        raise RuntimeError('the reference tree is only available in the build container')
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import logging
    import xicsrt
    # rocking_type='file' is broken as shipped (xicsrt_bragg.py:40,85,87 use an
    # undefined m_log); give the module the logger it meant to have.
    from xicsrt.tools import xicsrt_bragg
    if not hasattr(xicsrt_bragg, 'm_log'):
        xicsrt_bragg.m_log = logging.getLogger('xicsrt_bragg')
    logging.getLogger('xicsrt').setLevel(logging.ERROR)
    logging.getLogger().setLevel(logging.ERROR)
    return xicsrt


def reference_iteration(xicsrt, config):
    """One unsorted iteration through the reference's own Dispatchers."""
    from xicsrt import xicsrt_config
    from xicsrt.objects._Dispatcher import Dispatcher

    config = xicsrt_config.config_to_numpy(config)
    config = xicsrt_config.get_config(config)
    np.random.seed(config['general']['random_seed'])

    filters = None
    if 'filters' in config:
        filters = Dispatcher(config, 'filters')
        filters.instantiate()
        filters.setup()
        filters.initialize()
    sources = Dispatcher(config, 'sources')
    sources.instantiate()
    sources.apply_filters(filters)
    sources.setup()
    sources.check_param()
    sources.initialize()
    optics = Dispatcher(config, 'optics')
    optics.instantiate()
    optics.apply_filters(filters)
    optics.setup()
    optics.check_param()
    optics.initialize()

    rays = sources.generate_rays(keep_history=True)
    optics.trace(rays, keep_history=True, keep_images=True)

    history = dict(sources.history)
    history.update(optics.history)
    meta = dict(sources.meta)
    meta.update(optics.meta)
    return history, dict(optics.image), meta


def flatten(prefix, tree, out):
    for key, val in tree.items():
        name = f'{prefix}/{key}'
        if isinstance(val, dict):
            flatten(name, val, out)
        elif val is None:
            out[name] = np.array(np.nan)
        else:
            out[name] = np.asarray(val)


def make(scene_name):
    from oracle import scenes
    xicsrt = _import_reference()
    out = {}

    history, image, meta = reference_iteration(xicsrt, scenes.get(scene_name))
    flatten('iter', {k: dict(v) for k, v in history.items()}, out)
    flatten('iter_image', image, out)
    flatten('iter_meta', {k: v['num_out'] for k, v in meta.items()}, out)

    if scene_name in RUN_SCENES:
        res = xicsrt.raytrace(scenes.get(scene_name))
        flatten('run/total/meta', {k: v['num_out'] for k, v in res['total']['meta'].items()}, out)
        flatten('run/total/image', res['total']['image'], out)
        flatten('run/found', {k: dict(v) for k, v in res['found']['history'].items()}, out)
        flatten('run/lost', {k: dict(v) for k, v in res['lost']['history'].items()}, out)

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, scene_name + '.npz')
    np.savez_compressed(path, **out)
    return path, sum(v.nbytes for v in out.values())


# scenes that use product-only options (the reference does not know the keys): no fixture
NO_REFERENCE = ('mesh_torus_lossless',)


def main(argv):
    from oracle import scenes
    todo = argv or [n for n in scenes.names() if n not in NO_REFERENCE]
    for name in todo:
        path, raw = make(name)
        print(f'{name:28s} {os.path.getsize(path)/1e6:7.2f} MB on disk ({raw/1e6:6.2f} MB raw)')


if __name__ == '__main__':
    main(sys.argv[1:])
