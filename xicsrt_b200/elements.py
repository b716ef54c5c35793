# -*- coding: utf-8 -*-
"""
Element preparation: user config dict -> (full config, derived ``param`` dict).

This is the setup-time half of the reference objects (``setup`` /
``check_param`` / ``initialize``), restated as plain functions.  No ray math
happens here; the derived values are what the scene flattener
(:mod:`xicsrt_b200.scene`) packs into the device structs.

Reference sites:
  * frames            ``xicsrt/objects/_GeometryObject.py:64-111``
  * pixel grid        ``xicsrt/optics/_TraceObject.py:102-133``
  * sphere/cyl center ``xicsrt/optics/_ShapeSphere.py:37-43``, ``_ShapeCylinder.py:36-43``
  * torus             ``xicsrt/optics/_ShapeTorus.py:54-89``
  * crystal           ``xicsrt/optics/_InteractCrystal.py:86-88``
  * sources           ``xicsrt/sources/_XicsrtSourceGeneric.py:188-196``,
                      ``_XicsrtSourceDirected.py:40-44``
  * plasma            ``xicsrt/sources/_XicsrtPlasmaGeneric.py:160-174``
"""
import copy
import logging

import numpy as np

from . import config as xconfig
from . import registry

log = logging.getLogger('xicsrt_b200')


def cross3(a, b):
    """np.cross for two 3-vectors with the same operations in the same order (bit-identical), ten times cheaper."""
    a0, a1, a2 = float(a[0]), float(a[1]), float(a[2])
    b0, b1, b2 = float(b[0]), float(b[1]), float(b[2])
    return np.array([a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0])


def default_xaxis(zaxis):
    """cross([0,0,1], zaxis) normalised, or [1,0,0] when that vanishes."""
    xaxis = cross3((0.0, 0.0, 1.0), zaxis)
    if not np.all(xaxis == 0.0):
        xaxis = xaxis / np.linalg.norm(xaxis)
    else:
        xaxis = np.array([1.0, 0.0, 0.0])
    return xaxis


def build_config(config_user, strict=True):
    """defaults(class) overlaid by the user dict; unknown keys raise if strict."""
    cfg = registry.defaults(config_user['class_name'])
    xconfig.merge(cfg, config_user, strict=strict)
    return cfg


def _check_geometry_config(cfg):
    if cfg['xaxis'] is not None:
        zaxis = np.array(cfg['zaxis'], dtype=np.float64)
        xaxis = np.array(cfg['xaxis'], dtype=np.float64)
        if not np.isclose(np.dot(zaxis, xaxis), 0.0):
            raise ValueError('zaxis and xaxis are not orthogonal.')


def _param_from_config(cfg):
    param = copy.deepcopy(cfg)
    param = xconfig.to_numpy(param)
    return param


def _setup_geometry(param):
    param['origin'] = np.array(param['origin'], dtype=np.float64)
    param['zaxis'] = np.array(param['zaxis'], dtype=np.float64)
    if param['xaxis'] is None:
        param['xaxis'] = default_xaxis(param['zaxis'])
    else:
        param['xaxis'] = np.array(param['xaxis'], dtype=np.float64)
    xaxis, zaxis = param['xaxis'], param['zaxis']
    # rows are x, y = z cross x, z
    param['orientation'] = np.array([xaxis, cross3(zaxis, xaxis), zaxis])
    return param


def _convex_pair(value):
    v = np.asarray(value).astype(bool).ravel()
    if v.size != 2:
        raise Exception(f"Cannot be parse convex config option: {value}")
    return bool(v[0]), bool(v[1])


_mesh_cache = {}


def _cache_key(cfg):
    """Hashable image of a config dict (arrays by content); None when something is not hashable."""
    import hashlib
    items = []
    for key in sorted(cfg):
        val = cfg[key]
        if isinstance(val, np.ndarray):
            items.append((key, val.shape, str(val.dtype), hashlib.sha1(np.ascontiguousarray(val)).hexdigest()))
        elif isinstance(val, (list, tuple, dict)):
            try:
                items.append((key, repr(np.asarray(val).tolist()) if not isinstance(val, dict) else repr(sorted(val.items()))))
            except Exception:
                return None
        else:
            items.append((key, repr(val)))
    return tuple(items)


def prepare_optic(config_user, strict=True):
    """
    Returns (config, param) for one optic.  ``param`` additionally carries
    ``_interact`` and ``_shape`` (kinds from the registry).

    Prepared mesh optics (Delaunay triangulations, Clough-Tocher gradients, lookup tables:
    about a second of setup) are cached by config content, so repeated runs of one scene
    pay for them once.
    """
    cfg = build_config(config_user, strict=strict)
    _check_geometry_config(cfg)
    interact, shape = registry.OPTICS[cfg['class_name']]
    key = None
    if shape.startswith('mesh'):
        key = _cache_key(xconfig.to_numpy(dict(cfg)))
        if key is not None and key in _mesh_cache:
            return copy.deepcopy(cfg), _mesh_cache[key]
        if len(_mesh_cache) > 16:
            _mesh_cache.clear()

    param = _param_from_config(cfg)
    param['_interact'] = interact
    param['_shape'] = shape
    _setup_geometry(param)

    # mesh generators run in setup(), before the pixel grid is derived
    if shape.startswith('mesh'):
        from . import mesh
        mesh.setup_mesh(param)

    # pixel grid; truthiness test as in the reference (None or 0 disables)
    if param['xsize'] and param['ysize']:
        if param['pixel_size'] is None:
            param['pixel_size'] = param['xsize'] / 100
        pixel_xsize = param['xsize'] / param['pixel_size']
        pixel_ysize = param['ysize'] / param['pixel_size']
        if (abs(pixel_xsize - np.round(pixel_xsize)) >= 1.5e-7
                or abs(pixel_ysize - np.round(pixel_ysize)) >= 1.5e-7):
            log.warning(f"Optic width ({param['xsize']:0.4f}x{param['ysize']:0.4f})"
                        f"is not a multiple of the pixel_size ({param['pixel_size']:0.4f})."
                        f"May lead to truncation of output image.")
        param['pixel_xsize'] = int(np.round(pixel_xsize))
        param['pixel_ysize'] = int(np.round(pixel_ysize))
        param['enable_image'] = True
    else:
        param['enable_image'] = False

    if shape in ('sphere', 'cylinder'):
        sign = -1 if param['convex'] else 1
        param['center'] = sign * param['radius'] * param['zaxis'] + param['origin']
    elif shape == 'torus':
        r_minor = param['radius_minor']
        r_major = param['radius_major']
        if r_minor >= r_major:
            raise Exception(r'Cannot construct geometry with radius_major <= radius_minor.')
        param['torus_minor'] = r_minor
        cvx = _convex_pair(param['convex'])
        if cvx == (False, False):
            param['root_idx'], param['torus_major'], sign = 3, r_major - r_minor, 1
        elif cvx == (False, True):
            param['root_idx'], param['torus_major'], sign = 2, r_major + r_minor, 1
        elif cvx == (True, False):
            param['root_idx'], param['torus_major'], sign = 1, r_major + r_minor, -1
        else:
            param['root_idx'], param['torus_major'], sign = 0, r_major - r_minor, -1
        param['center'] = param['origin'] + sign * r_major * param['zaxis']
    elif shape.startswith('mesh'):
        from . import mesh
        mesh.initialize_mesh(param)

    if interact in ('crystal', 'mosaic'):
        param['rocking_type'] = str.lower(param['rocking_type'])

    if key is not None:
        _mesh_cache[key] = param
    return cfg, param


def prepare_source(config_user, strict=True, poisson=None):
    """
    Returns (config, param) for a point/box source (generic/directed/focused).

    poisson : callable(lam) -> int, used when ``use_poisson`` is set.  The
              caller owns the random stream (Philox on the product path, the
              legacy numpy stream in the oracle).
    """
    cfg = build_config(config_user, strict=strict)
    _check_geometry_config(cfg)
    kind = registry.SOURCES[cfg['class_name']]
    param = _param_from_config(cfg)
    param['_kind'] = kind
    _setup_geometry(param)

    if kind.startswith('plasma'):
        _initialize_plasma(cfg, param)
        return cfg, param

    if param['use_poisson']:
        if poisson is None:
            raise ValueError('use_poisson needs a poisson sampler')
        param['intensity'] = poisson(param['intensity'])
    else:
        if param['intensity'] < 1:
            raise ValueError('intensity of less than one encountered. Turn on poisson statistics.')
    param['intensity'] = int(param['intensity'])

    if kind == 'directed' and param['direction'] is None:
        param['direction'] = param['zaxis']
    if kind == 'focused':
        if param['target'] is None:
            raise ValueError('XicsrtSourceFocused needs a target.')
        param['target'] = np.asarray(param['target'], dtype=np.float64)
    return cfg, param


def _initialize_plasma(cfg, param):
    if param['max_rays'] is not None:
        param['max_rays'] = int(param['max_rays'])
    param['volume'] = cfg['xsize'] * cfg['ysize'] * cfg['zsize']
    if param['bundle_count'] is None:
        param['bundle_count'] = param['volume'] / param['bundle_volume']
    param['bundle_count'] = int(np.round(param['bundle_count']))
    if param['bundle_count'] < 1:
        raise Exception(f'Bundle volume is larger than the plasma volume.')
    if param['bundle_count'] > param['max_bundles']:
        raise ValueError(
            f"Current settings will produce too many bundles ({param['bundle_count']:0.2e}). "
            f"Increase the bundle_volume, explicitly set bundle_count or increase max_bundles.")
    if param['bundle_type'] == 'point':
        param['voxel_size'] = 0.0
    elif param['bundle_type'] == 'voxel':
        param['voxel_size'] = param['bundle_volume'] ** (1 / 3)
    else:
        raise Exception(f"bundle_type {param['bundle_type']} unknown.")
    if param['target'] is not None:
        param['target'] = np.asarray(param['target'], dtype=np.float64)


def prepare_filter(config_user, strict=True):
    cfg = build_config(config_user, strict=strict)
    _check_geometry_config(cfg)
    param = _param_from_config(cfg)
    param['_kind'] = registry.FILTERS[cfg['class_name']]
    _setup_geometry(param)
    return cfg, param
