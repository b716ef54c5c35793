# -*- coding: utf-8 -*-
"""
oracle.make_io_fixtures -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Writes ``tests/golden/io/`` with result files produced by the UNMODIFIED reference's own
``xicsrt_io.save_results`` / ``save_config`` / ``save_images`` (xicsrt/xicsrt_io.py:28-118) for a small run of the
scene ``sphere``: the results as pickle, the config as pickle and json, and one TIFF per imaged optic.  tests/test_io_reference_files.py loads them with
``xicsrt_b200.io`` and checks that files written here have the same keys, dtypes and orientation.  (hdf5 is left
out: h5py is not installed in the build container, so the reference cannot write one here.)

Usage:  PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_io_fixtures
"""
import os
import shutil
import sys

REFERENCE = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'io')


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REFERENCE)
    import logging
    import xicsrt
    from xicsrt import xicsrt_io
    from oracle import scenes
    logging.getLogger('xicsrt').setLevel(logging.ERROR)
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    cfg = scenes.get('sphere')
    cfg['sources']['source']['intensity'] = 2000
    cfg['general'].update({'output_path': OUT, 'output_prefix': 'ref', 'history_max_lost': 50, 'save_images': False})
    res = xicsrt.raytrace(cfg)
    # results: pickle (the reference's json writer fails on the numpy integers of meta['num_out']; hdf5 needs h5py)
    res['config']['general']['results_ext'] = '.pkl'
    xicsrt_io.save_results(res, path=OUT, overwrite=True)
    for ext in ('.pkl', '.json'):
        res['config']['general']['config_ext'] = ext
        xicsrt_io.save_config(res['config'], path=OUT, overwrite=True)
    xicsrt_io.save_images(res, path=OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
