# -*- coding: utf-8 -*-
"""
Plasma (extended) sources: the per-iteration bundle table.

The reference scatters ``bundle_count`` bundles in a box, evaluates emissivity /
temperature / velocity per bundle and then builds one ``XicsrtSourceFocused`` per
bundle in a Python loop (``xicsrt/sources/_XicsrtPlasmaGeneric.py:176-382``, 0.5 ms
per bundle).  Here the same quantities are computed as whole-array operations and
packed into the device table the kernel indexes by ray id: ``XrtBundle[n]`` plus the
inclusive prefix sum of the rays per bundle (a ray finds its bundle by binary search).

Random numbers on this (host) side -- bundle centres and Poisson ray counts -- come from
the caller's generator: a numpy Philox keyed by ``(seed, iteration)`` on the product path,
the legacy MT19937 stream in the parity tests.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import voigt


def solid_angle_isotropic(spread):
    """xicsrt/tools/xicsrt_spread.py:112-128 -- 4 pi sin^2(theta / 2), array form."""
    return 4 * np.pi * np.sin(np.asarray(spread, dtype=np.float64) / 2)**2


def sightline_mask(fparam, origin):
    """xicsrt/filters/_XicsrtBundleFilterSightline.py:31-56 on bundle centres."""
    if fparam['_kind'] == 'none':
        return np.ones(len(origin), dtype=np.bool_)
    axis = np.asarray(fparam['zaxis'], dtype=np.float64)
    l0 = np.asarray(fparam['origin'], dtype=np.float64) - origin
    along = np.outer(np.einsum('j,ij->i', axis, l0), axis)
    perp = l0 - along
    dist = np.sqrt(np.einsum('ij,ij->i', perp, perp))
    return fparam['radius'] >= dist


def rho_toroidal(param, points):
    """
    Normalised minor radius of _XicsrtPlasmaToroidal.py:34-42 for many points: toroidal
    coordinates about ``torus_origin`` (xicsrt_math.py:211-244), then sqrt(r^2 / minor_radius)
    -- the division by minor_radius (not its square) is the reference's and is kept.
    """
    p = points - np.asarray(param['torus_origin'], dtype=np.float64)
    d = np.sqrt(p[:, 0]**2 + p[:, 1]**2) - param['major_radius']
    r = np.sqrt(np.power(p[:, 2], 2) + np.power(d, 2))
    return np.sqrt(r**2 / param['minor_radius'])


def _profile(fname, rho):
    data = np.loadtxt(fname, dtype=np.float64)
    return np.interp(rho, data[:, 0], data[:, 1], left=0.0, right=0.0)


def bundle_properties(param, filters, rng):
    """
    setup_bundles + bundle_filter + bundle_generate (_XicsrtPlasmaGeneric.py:176-250 and the
    Cubic / Toroidal / ToroidalDatafile overrides) -> dict of per-bundle arrays.
    """
    nb = int(param['bundle_count'])
    off = np.zeros((nb, 3))
    off[:, 0] = rng.uniform(-1 * param['xsize'] / 2, param['xsize'] / 2, nb)
    off[:, 1] = rng.uniform(-1 * param['ysize'] / 2, param['ysize'] / 2, nb)
    off[:, 2] = rng.uniform(-1 * param['zsize'] / 2, param['zsize'] / 2, nb)
    origin = np.einsum('ij,ki->kj', param['orientation'], off) + param['origin']

    if param['spread_radius'] is not None:
        dist = np.linalg.norm(origin - param['target'], axis=1)
        spread = np.arctan(param['spread_radius'] / dist)
    else:
        spread = np.full(nb, float(np.atleast_1d(param['spread'])[0]))

    b = {'origin': origin, 'spread': spread, 'solid_angle': solid_angle_isotropic(spread),
         'temperature': np.ones(nb), 'emissivity': np.ones(nb), 'velocity': np.zeros((nb, 3)),
         'mask': np.ones(nb, dtype=np.bool_)}
    for f in filters:
        b['mask'] &= sightline_mask(f, origin)

    kind = param['_kind']
    if kind == 'plasma_cubic':
        b['temperature'][:] = param['temperature']
        b['emissivity'][:] = param['emissivity']
    elif kind in ('plasma_toroidal', 'plasma_datafile'):
        m = b['mask']
        rho = rho_toroidal(param, origin[m])
        if kind == 'plasma_datafile':
            temp = _profile(param['temperature_file'], rho)
            emis = _profile(param['emissivity_file'], rho)
        else:
            temp, emis = param['temperature'], param['emissivity']
        b['temperature'][m] = temp * param['temperature_scale']
        b['emissivity'][m] = emis * param['emissivity_scale']
        b['velocity'][m] = param['velocity'] * param['velocity_scale']
        m &= np.isfinite(b['temperature'])
    return b


def bundle_intensity(param, b):
    """Expected photons per bundle (_XicsrtPlasmaGeneric.py:301-319)."""
    inten = b['emissivity'] * param['time_resolution'] * param['bundle_volume'] * b['solid_angle'] / (4 * np.pi)
    return inten * (param['volume'] / (param['bundle_count'] * param['bundle_volume']))


def bundle_counts(param, b, rng):
    """Rays per bundle: Poisson draw or truncation, as each per-bundle source does at initialize."""
    m = b['mask']
    inten = bundle_intensity(param, b)
    predicted = int(np.sum(inten[m]))
    if param['max_rays'] and predicted > param['max_rays']:
        raise ValueError(
            f"Current settings will produce too many rays ({predicted:0.2e}). "
            f"Please reduce integration time or adjust other parameters.")
    counts = np.zeros(len(m), dtype=np.int64)
    if param['use_poisson']:
        counts[m] = rng.poisson_array(inten[m])
    else:
        if np.any(inten[m] < 1):
            raise ValueError('intensity of less than one encountered. Turn on poisson statistics.')
        counts[m] = inten[m].astype(np.int64)
    return counts


def build_bundles(param, filters, rng):
    """
    The device bundle table of one iteration.  Returns {'table', 'end', 'n_rays', 'props', 'counts'}.
    Bundles that emit no ray are left out of the table.
    """
    if str(param['angular_dist']).lower() != 'isotropic':
        raise NotImplementedError('plasma sources on the device support angular_dist="isotropic" only')
    if float(param['linewidth']) != 0.0 and str(param['wavelength_dist']).lower() == 'voigt':
        raise NotImplementedError('plasma sources with a natural linewidth need per-bundle Voigt tables')
    if param['target'] is None:
        raise ValueError('plasma sources need a target')
    b = bundle_properties(param, filters, rng)
    counts = bundle_counts(param, b, rng)
    keep = counts > 0
    n = int(np.sum(keep))
    if n == 0:
        raise ValueError('No rays generated. Check plasma input parameters')

    lam0 = float(param['wavelength'])
    temp = b['temperature'][keep]
    sigma = np.where(temp > 0, voigt.doppler_sigma(np.abs(temp), param['mass_number'], lam0), 0.0)
    if str(param['wavelength_dist']).lower() != 'voigt':
        sigma = np.zeros(n)

    rec = np.zeros(n, dtype=[('origin', 'f8', 3), ('cos_spread', 'f8'), ('wave_sigma', 'f8'), ('velocity_c', 'f8', 3)])
    rec['origin'] = b['origin'][keep]
    rec['cos_spread'] = np.cos(b['spread'][keep])
    rec['wave_sigma'] = sigma
    rec['velocity_c'] = b['velocity'][keep] / voigt.C_LIGHT
    assert rec.dtype.itemsize == C.sizeof(L.XrtBundle)
    table = np.ascontiguousarray(rec)
    end = np.cumsum(counts[keep]).astype(np.uint64)
    return {'table': table, 'end': end, 'n_rays': int(end[-1]), 'props': b, 'counts': counts}
