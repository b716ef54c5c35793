# -*- coding: utf-8 -*-
"""
Files written after a trace (host side, outside the hot path): config / results
as json, pickle or hdf5 and per-optic images as TIFF.  File naming and the
``rot90`` image orientation follow the reference (``xicsrt/xicsrt_io.py:27-146``)
so that files written here load with the reference's ``load_results`` and the
other way round.
"""
import copy
import json
import logging
import pathlib
import pickle

import numpy as np

from . import config as xconfig

log = logging.getLogger('xicsrt_b200')

_EXT_KEY = {'image': 'image_ext', 'results': 'results_ext', 'config': 'config_ext'}


def generate_filename(config, kind=None, name=None, path=None):
    """``<prefix>_<name>_<suffix>_<run_suffix><ext>`` under ``output_path`` (xicsrt_io.py:116-146)."""
    g = xconfig.get_config(config)['general']
    if kind is None:
        ext = ''
    elif kind in _EXT_KEY:
        ext = g[_EXT_KEY[kind]]
    else:
        raise Exception(f'Data kind {kind} unknown.')
    parts = (g['output_prefix'], kind if name is None else name, g['output_suffix'], g['output_run_suffix'])
    base = '_'.join(p for p in parts if p) + ext
    return str(pathlib.Path(g['output_path'] if path is None else path) / base)


def _ensure_parent(filename):
    p = pathlib.Path(filename).expanduser()
    p = p.parent if p.suffix else p
    p.mkdir(parents=True, exist_ok=True)


def _kind_of(filename):
    ext = pathlib.Path(filename).suffix
    if 'pickle' in ext or 'pkl' in ext:
        return 'pickle'
    if 'json' in ext:
        return 'json'
    if 'hdf5' in ext or 'h5' in ext:
        return 'hdf5'
    raise NotImplementedError(f'filetype: {ext} not currently supported.')


def _jsonable(obj):
    if isinstance(obj, dict):
        return {k: _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_jsonable(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.generic):
        return obj.item()
    return obj


def write_dict(data, filename, mkdir=False, overwrite=False):
    if mkdir:
        _ensure_parent(filename)
    path = pathlib.Path(filename).expanduser()
    if not overwrite and path.exists():
        raise FileExistsError("File exists. Use overwrite=True to overwrite.")
    kind = _kind_of(path)
    if kind == 'pickle':
        with open(path, 'wb') as ff:
            pickle.dump(data, ff)
    elif kind == 'json':
        with open(path, 'w') as ff:
            json.dump(_jsonable(copy.deepcopy(data)), ff, indent=2)
    else:
        _write_hdf5(data, path)


def read_dict(filename):
    path = pathlib.Path(filename).expanduser()
    kind = _kind_of(path)
    if kind == 'pickle':
        with open(path, 'rb') as ff:
            return pickle.load(ff)
    if kind == 'json':
        with open(path, 'r') as ff:
            return xconfig.to_numpy(json.load(ff))
    return _read_hdf5(path)


def _h5py():
    try:
        import h5py
    except ImportError as err:
        raise ImportError('hdf5 results need h5py, which is not installed; '
                          'set general.results_ext to ".pkl" or ".json".') from err
    return h5py


def _write_hdf5(data, path):
    h5py = _h5py()

    def put(group, tree):
        for key, val in tree.items():
            if isinstance(val, dict):
                put(group.create_group(str(key)), val)
            elif val is None:
                group.attrs[f'{key}__none'] = True
            elif isinstance(val, str):
                group.attrs[str(key)] = val
            else:
                arr = np.asarray(val)
                if arr.dtype.kind in 'OU':
                    group.attrs[str(key)] = json.dumps(_jsonable(val))
                    group.attrs[f'{key}__json'] = True
                else:
                    group.create_dataset(str(key), data=arr)
    with h5py.File(path, 'w') as ff:
        put(ff, data)


def _read_hdf5(path):
    h5py = _h5py()

    def get(group):
        out = {}
        for key, val in group.items():
            out[key] = get(val) if isinstance(val, h5py.Group) else val[()]
        for key, val in group.attrs.items():
            if key.endswith('__none'):
                out[key[:-6]] = None
            elif key.endswith('__json'):
                continue
            elif f'{key}__json' in group.attrs:
                out[key] = json.loads(val)
            else:
                out[key] = val
        return out
    with h5py.File(path, 'r') as ff:
        return get(ff)


def load_config(filename):
    return read_dict(filename)


def save_config(config, filename=None, path=None, mkdir=None, overwrite=None):
    if filename is None:
        filename = generate_filename(config, kind='config', path=path)
    elif path is not None:
        filename = str(pathlib.Path(path) / filename)
    if mkdir is None:
        mkdir = config.get('general', {}).get('make_directories', False)
    write_dict(config, filename, mkdir=mkdir, overwrite=overwrite)
    log.info('Config saved to {}'.format(filename))


def save_results(output, filename=None, path=None, mkdir=None, overwrite=None):
    config = output['config']
    if filename is None:
        filename = generate_filename(config, kind='results', path=path)
    elif path is not None:
        filename = str(pathlib.Path(path) / filename)
    if mkdir is None:
        mkdir = config['general'].get('make_directories', False)
    write_dict(output, filename, mkdir=mkdir, overwrite=overwrite)
    log.info('History saved to {}'.format(filename))


def load_results(filename=None, path=None, config=None):
    if filename is None:
        filename = generate_filename(config, kind='results', path=path)
    elif path is not None:
        filename = str(pathlib.Path(path) / filename)
    return read_dict(filename)


def save_images(output, rotate=True, path=None, mkdir=None):
    """One float TIFF per imaged optic, rotated by 90 degrees like the reference (xicsrt_io.py:92-113)."""
    from PIL import Image
    config = output['config']
    if mkdir is None:
        mkdir = config['general'].get('make_directories', False)
    for name in config['optics']:
        img = output['total']['image'].get(name)
        if img is None:
            continue
        filename = generate_filename(config, 'image', name, path=path)
        if mkdir:
            _ensure_parent(filename)
        Image.fromarray(np.rot90(img) if rotate else img).save(filename)
        log.info('Saved image: {}'.format(filename))
