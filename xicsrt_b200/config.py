# -*- coding: utf-8 -*-
"""
Config handling for the B200 raytrace path.

This mirrors the *semantics* of the reference's config system so that the
same user dictionaries are accepted and the fully-defaulted config comes back
in ``output['config']``:

* top-level sections and ``general`` defaults: reference
  ``xicsrt/xicsrt_config.py:29-205``
* recursive merge with strict unknown-key errors: reference
  ``xicsrt/xicsrt_config.py:294-364``
* list -> ndarray conversion of numeric config values: reference
  ``xicsrt/tools/xicsrt_misc.py:18-51``

Only dictionary plumbing lives here; nothing in this file touches rays.
"""
import copy

import numpy as np

__version__ = '0.8.13+b200'


def general_defaults():
    """The ``general`` section (reference ``xicsrt_config.py:169-197``)."""
    g = dict()
    g['version'] = __version__
    g['number_of_iter'] = 1
    g['number_of_runs'] = 1
    g['random_seed'] = None
    g['pathlist'] = []
    g['pathlist_default'] = []
    g['strict_config_check'] = True

    g['output_path'] = None
    g['output_prefix'] = 'xicsrt'
    g['output_suffix'] = None
    g['output_run_suffix'] = None
    g['image_ext'] = '.tif'
    g['results_ext'] = '.hdf5'
    g['config_ext'] = '.json'
    g['make_directories'] = False

    g['keep_meta'] = True
    g['keep_images'] = True
    g['keep_history'] = True

    g['history_max_lost'] = 10000

    g['save_config'] = False
    g['save_images'] = False
    g['save_results'] = False

    g['print_results'] = True
    return g


def default_config():
    return {
        'general': general_defaults(),
        'sources': dict(),
        'optics': dict(),
        'filters': dict(),
        'scenario': dict(),
    }


def merge(base, new, strict=True, update=False, ignore_none=False):
    """
    Recursively overwrite ``base`` with ``new`` (in place, returns base).

    strict      unknown key in ``new`` raises (same message as the reference).
    update      when not strict, unknown keys are kept instead of dropped.
    ignore_none ``None`` values in ``new`` do not overwrite.
    """
    if new is None:
        return base
    for key, val in new.items():
        if key not in base:
            if strict:
                raise Exception("User option not recognized: {}".format(key))
            if update:
                base[key] = val
            continue
        if isinstance(base[key], dict) and isinstance(val, dict):
            merge(base[key], val, strict=strict, update=update, ignore_none=ignore_none)
        elif ignore_none and val is None:
            continue
        else:
            base[key] = val
    return base


def get_config(config_user=None):
    """Full config = defaults overlaid (non-strict, keeping extras) by the user dict."""
    return merge(default_config(), config_user, strict=False, update=True)


def to_numpy(obj):
    """
    Convert non-empty numeric lists nested in dicts/lists to ndarrays (a new
    container is returned, like the reference's non-inplace conversion).
    String lists are left alone, object lists are descended into.
    """
    if isinstance(obj, dict):
        out = dict(obj)
        keys = list(out.keys())
    elif isinstance(obj, list):
        out = list(obj)
        keys = range(len(out))
    else:
        raise TypeError('Object must be either a dict or a list.')

    for key in keys:
        val = out[key]
        if isinstance(val, list):
            if not val:
                continue
            arr = np.array(val)
            if arr.dtype.char == 'U':
                continue
            if arr.dtype.char == 'O':
                out[key] = to_numpy(val)
            else:
                out[key] = arr
        elif isinstance(val, dict):
            out[key] = to_numpy(val)
    return out


def from_numpy(obj):
    """Inverse of :func:`to_numpy` (ndarray -> list), for json export."""
    if isinstance(obj, dict):
        out = dict(obj)
        keys = list(out.keys())
    elif isinstance(obj, list):
        out = list(obj)
        keys = range(len(out))
    else:
        raise TypeError('Object must be either a dict or a list.')
    for key in keys:
        val = out[key]
        if isinstance(val, np.ndarray):
            out[key] = val.tolist()
        elif isinstance(val, (dict, list)):
            out[key] = from_numpy(val)
    return out


def deepcopy(config):
    return copy.deepcopy(config)
