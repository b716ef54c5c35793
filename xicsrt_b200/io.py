# -*- coding: utf-8 -*-
"""
Files written after a trace (host side, outside the hot path): config / results
as json, pickle or hdf5 and per-optic images as TIFF.  File naming, the ``rot90``
image orientation (``xicsrt/xicsrt_io.py:27-146``) and the hdf5 group / attribute
conventions (``xicsrt/util/mirhdf5.py``) follow the reference, so that files written
here load with the reference's ``load_results`` and the other way round
(``tests/test_io_reference_files.py`` loads files the reference itself wrote).  hdf5 needs
h5py on both sides; without it ``results_ext`` must be ``.pkl`` or ``.json``.
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


class RayArray(dict):
    """Stand-in for the reference's dict subclass of the same name (xicsrt/objects/_RayArray.py:12-96)."""


class _ReferenceUnpickler(pickle.Unpickler):
    """
    Result pickles written by the reference hold its ``RayArray`` dict subclass by class reference
    (the lost histories, xicsrt_raytrace.py:268-276): map it to a plain dict subclass so that such files load
    where the reference package is not installed.
    """

    def find_class(self, module, name):
        if module.split('.')[0] == 'xicsrt' and name == 'RayArray':
            return RayArray
        return super().find_class(module, name)


def _plain(tree):
    if isinstance(tree, dict):
        return {k: _plain(v) for k, v in tree.items()}
    return tree


def read_dict(filename):
    path = pathlib.Path(filename).expanduser()
    kind = _kind_of(path)
    if kind == 'pickle':
        with open(path, 'rb') as ff:
            return _plain(_ReferenceUnpickler(ff).load())
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


# The hdf5 layout is the one of the reference's util/mirhdf5.py (:185-245 writing, :248-330 reading), so that
# result files move between the two programs:
#   dict  -> group with attrs '_mirhdf5 python object type' = b'dict' and '_mirhdf5 dictionary order' = [key bytes]
#   list  -> group with attr  '_mirhdf5 python object type' = b'list', items under '0000', '0001', ...
#   None  -> dataset False with attr '_mirhdf5 python None' = True
#   str   -> scalar dataset with attr '_mirhdf5 python str' = True
#   other -> dataset (numpy arrays, numbers, bools)
# The functions take any object with h5py's group interface (tests drive them with an in-memory stand-in when h5py
# is not installed).
H5_TYPE, H5_ORDER, H5_NONE, H5_STR = ('_mirhdf5 python object type', '_mirhdf5 dictionary order',
                                      '_mirhdf5 python None', '_mirhdf5 python str')


def hdf5_put_item(group, key, item):
    if isinstance(item, dict):
        hdf5_put_dict(group.create_group(key), item)
    elif isinstance(item, (list, tuple)):
        sub = group.create_group(key)
        sub.attrs[H5_TYPE] = 'list'.encode()
        for ii, val in enumerate(item):
            hdf5_put_item(sub, '{:04d}'.format(ii), val)
    elif item is None:
        group[key] = False
        group[key].attrs[H5_NONE] = True
    elif isinstance(item, str):
        group.create_dataset(key, data=item)
        group[key].attrs[H5_STR] = True
    else:
        group.create_dataset(key, data=item if np.isscalar(item) else np.asarray(item))


def hdf5_put_dict(group, tree):
    group.attrs[H5_TYPE] = 'dict'.encode()
    group.attrs[H5_ORDER] = [str(key).encode() for key in tree.keys()]
    for key, val in tree.items():
        hdf5_put_item(group, str(key), val)


def hdf5_get(node, is_group):
    """node: a group or dataset; is_group(node) tells which."""
    attrs = node.attrs
    if is_group(node):
        kind = attrs[H5_TYPE] if H5_TYPE in attrs else 'dict'
        kind = kind.decode() if hasattr(kind, 'decode') else kind
        if kind == 'list':
            return [hdf5_get(node[key], is_group) for key in node.keys()]
        if kind != 'dict':
            raise Exception('Unknown group type: {}'.format(kind))
        keys = attrs[H5_ORDER] if H5_ORDER in attrs else node.keys()
        out = {}
        for key in keys:
            key = key.decode() if hasattr(key, 'decode') else key
            out[key] = hdf5_get(node[key], is_group)
        return out
    if H5_NONE in attrs:
        return None
    val = node[()]
    if H5_STR in attrs:
        val = val.decode() if hasattr(val, 'decode') else str(val)
    return val


def _write_hdf5(data, path):
    h5py = _h5py()
    if not isinstance(data, dict):
        raise Exception('Incorrect input type. Dictionary expected.')
    with h5py.File(path, 'w') as ff:
        hdf5_put_dict(ff, data)


def _read_hdf5(path):
    h5py = _h5py()
    with h5py.File(path, 'r') as ff:
        return hdf5_get(ff, lambda node: isinstance(node, h5py.Group))


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
