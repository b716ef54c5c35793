# -*- coding: utf-8 -*-
"""
oracle.sources -- TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy restatement of ray generation: box/gaussian origins, emission cones,
wavelength lines, plasma bundles.  Draw order follows the reference so that a
:class:`oracle.stream.LegacyStream` reproduces its rays exactly.
"""
import numpy as np

from xicsrt_b200 import elements, voigt

from oracle import vecs


# ---------------------------------------------------------------------------
# cone distributions -- reference xicsrt/tools/xicsrt_spread.py

def _one(spread):
    s = np.atleast_1d(np.asarray(spread, dtype=np.float64))
    if s.size != 1:
        raise Exception('Spread must be a scalar or one element array.')
    return s


def _four(spread):
    s = np.atleast_1d(np.asarray(spread, dtype=np.float64))
    if s.size == 1:
        return [-s[0], s[0], -s[0], s[0]]
    if s.size == 2:
        return [-s[0], s[0], -s[1], s[1]]
    if s.size == 4:
        return [s[0], s[1], s[2], s[3]]
    raise Exception('Spread must have 1, 2 or 3 elements. See docstring.')


def cone_isotropic(spread, n, stream, site='src.cone'):
    """xicsrt_spread.py:80-110 -- z ~ U(cos t, 1), phi ~ U(0, 2pi)."""
    theta = _one(spread)
    z = stream.uniform(np.cos(theta), 1, n, site=site and site + '.0')
    phi = stream.uniform(0, 2 * np.pi, n, site=site and site + '.1')
    out = np.empty((n, 3))
    rho = np.sqrt(1 - z**2)
    out[:, 0] = rho * np.cos(phi)
    out[:, 1] = rho * np.sin(phi)
    out[:, 2] = z
    return out


def cone_isotropic_xy(spread, n, stream):
    """xicsrt_spread.py:130-196 -- rejection from the enclosing circular cone."""
    th = _four(spread)
    tx = np.max(np.abs(th[0:2]))
    ty = np.max(np.abs(th[2:]))
    theta_max = np.arcsin(np.sqrt(np.sin(tx)**2 + np.sin(ty)**2))

    out = np.empty((n, 3))
    filled = 0
    while filled < n:
        v = cone_isotropic(theta_max, n, stream, site=None)
        sx = v[:, 0] / np.sqrt(v[:, 0]**2 + v[:, 2]**2)
        sy = v[:, 1] / np.sqrt(v[:, 1]**2 + v[:, 2]**2)
        ok = (sx > np.sin(th[0])) & (sx <= np.sin(th[1])) & (sy > np.sin(th[2])) & (sy <= np.sin(th[3]))
        take = min(int(np.sum(ok)), n - filled)
        out[filled:filled + take] = v[ok][:take]
        filled += take
    return out


def _from_angles(a0, a1):
    out = np.empty((len(a0), 3))
    out[:, 0] = np.cos(a1) * np.sin(a0)
    out[:, 1] = np.sin(a1) * np.sin(a0)
    out[:, 2] = np.cos(a0)
    return out


def cone_flat(spread, n, stream):
    """xicsrt_spread.py:213-245 -- r = sqrt(U(0, tan t)), angle ~ U(0, 2pi)."""
    theta = _one(spread)
    r = np.sqrt(stream.uniform(0, np.tan(theta), n, site='src.cone.0'))
    a1 = stream.uniform(0, 2 * np.pi, n, site='src.cone.1')
    return _from_angles(np.arctan(r), a1)


def cone_flat_xy(spread, n, stream):
    """xicsrt_spread.py:247-294 -- x, y uniform on the z=1 plane."""
    rng = np.tan(_four(spread))
    x = stream.uniform(rng[0], rng[1], n, site='src.cone.0')
    y = stream.uniform(rng[2], rng[3], n, site='src.cone.1')
    return _from_angles(np.arctan(np.sqrt(x**2 + y**2)), np.arctan2(y, x))


def cone(spread, n, name, stream):
    name = 'isotropic' if name is None else name.lower()
    if name == 'isotropic':
        return cone_isotropic(spread, n, stream)
    if name == 'isotropic_xy':
        return cone_isotropic_xy(spread, n, stream)
    if name == 'flat':
        return cone_flat(spread, n, stream)
    if name == 'flat_xy':
        return cone_flat_xy(spread, n, stream)
    if name == 'gaussian':
        # xicsrt_spread.py:55 names an undefined function; the reference
        # raises NameError here.
        raise NotImplementedError('angular_dist "gaussian" is not implemented in the reference.')
    raise Exception(f'Distribution "{name}" is not known.')


def solid_angle_isotropic(spread):
    """xicsrt_spread.py:112-128."""
    theta = _one(spread)
    return 4 * np.pi * np.sin(theta[0] / 2)**2


# ---------------------------------------------------------------------------
# one box source -- reference xicsrt/sources/_XicsrtSourceGeneric.py:198-393

def box_origins(param, n, stream):
    """:229-255 -- three uniforms (drawn even for zero sizes) or one mvn."""
    if param['spatial_dist'] == 'uniform':
        dx = stream.uniform(-1 * param['xsize'] / 2, param['xsize'] / 2, n, site='src.origin.0')
        dy = stream.uniform(-1 * param['ysize'] / 2, param['ysize'] / 2, n, site='src.origin.1')
        dz = stream.uniform(-1 * param['zsize'] / 2, param['zsize'] / 2, n, site='src.origin.2')
    elif param['spatial_dist'] == 'gaussian':
        k = 2 * np.sqrt(2 * np.log(2))
        cov = np.diag([param['xsize']**2, param['ysize']**2, param['zsize']**2]) / k**2
        dx, dy, dz = stream.mvn([0, 0, 0], cov, n, site='src.origin.g').T
    else:
        raise NotImplementedError(f"spatial_dist: {param['spatial_dist']} not implemented.")
    R = param['orientation']
    return (param['origin'] + np.outer(dx, R[0]) + np.outer(dy, R[1]) + np.outer(dz, R[2]))


def cone_axes(param, origin):
    """:262-266, _XicsrtSourceDirected.py:46-50, _XicsrtSourceFocused.py:40-44."""
    n = len(origin)
    kind = param['_kind']
    if kind == 'focused':
        a = param['target'] - origin
    else:
        a = np.empty((n, 3))
        a[:] = param['direction'] if kind == 'directed' else param['zaxis']
    return a / np.linalg.norm(a, axis=1)[:, None]


def rotate_cone(param, axis, local):
    """:268-293 -- basis (o_2, o_1, axis); o_1 = axis x xaxis + axis x zaxis."""
    o1 = np.cross(axis, param['xaxis']) + np.cross(axis, param['zaxis'])
    o1 /= np.linalg.norm(o1, axis=1)[:, None]
    o2 = np.cross(axis, o1)
    o2 /= np.linalg.norm(o2, axis=1)[:, None]
    return local[:, 0:1] * o2 + local[:, 1:2] * o1 + local[:, 2:3] * axis


def wavelengths(param, direction, n, stream):
    """:295-367 plus xicsrt_voigt.py:119-130."""
    model = voigt.wavelength_model(param)
    if model['mode'] == 'const':
        lam = np.ones(n, dtype=np.float64) * model['wavelength']
    elif model['mode'] == 'uniform':
        lam = stream.uniform(model['lo'], model['hi'], n, site='src.wave')
    elif model['mode'] == 'normal':
        lam = stream.normal(model['wavelength'], model['sigma'], n, site='src.wave')
    else:
        cdf, x = model['cdf'], model['x']
        y = stream.uniform(np.min(cdf), np.max(cdf), n, site='src.wave')
        lam = np.interp(y, cdf, x)
        lam += model['wavelength']
        # the reference bumps a zero temperature to 1 eV for good (:339)
        if float(param['temperature']) == 0.0:
            param['temperature'] = param['temperature'] + 1.0

    vel = np.asarray(param['velocity'], dtype=np.float64)
    if not np.all(vel == 0.0):
        lam *= 1 - (np.einsum('j,ij->i', vel, direction) / voigt.C_LIGHT)
    return lam


def box_source(param, stream, filters=()):
    n = param['intensity']
    rays = {}
    rays['origin'] = box_origins(param, n, stream)
    axis = cone_axes(param, rays['origin'])
    local = cone(param['spread'], n, param['angular_dist'], stream)
    rays['direction'] = rotate_cone(param, axis, local)
    rays['wavelength'] = wavelengths(param, rays['direction'], n, stream)
    rays['weight'] = np.ones(n, dtype=np.float64)
    rays['mask'] = np.ones(n, dtype=np.bool_)
    for f in filters:
        rays = sightline(f, rays)
    return rays


# ---------------------------------------------------------------------------
# bundle filter -- reference xicsrt/filters/_XicsrtBundleFilterSightline.py:31-56

def sightline(fparam, bundle):
    if fparam['_kind'] == 'none':
        return bundle
    axis = np.asarray(fparam['zaxis'], dtype=np.float64)
    l0 = np.asarray(fparam['origin'], dtype=np.float64) - bundle['origin']
    along = np.outer(np.einsum('j,ij->i', axis, l0), axis)
    perp = l0 - along
    dist = np.sqrt(np.einsum('ij,ij->i', perp, perp))
    bundle['mask'] &= (fparam['radius'] >= dist)
    return bundle


# ---------------------------------------------------------------------------
# plasma -- reference xicsrt/sources/_XicsrtPlasmaGeneric.py:176-393

def tor_from_car(p, major_radius):
    """xicsrt/tools/xicsrt_math.py:211-244 (single point)."""
    d = np.linalg.norm(p[0:2]) - major_radius
    out = np.empty(3)
    out[2] = np.arctan2(p[1], p[0])
    out[1] = np.arctan2(p[2], d)
    out[0] = np.sqrt(np.power(p[2], 2) + np.power(d, 2))
    return out


def plasma_rho(param, p):
    """_XicsrtPlasmaToroidal.py:34-42 -- note r**2 / minor_radius (sic)."""
    flx = tor_from_car(p - param['torus_origin'], param['major_radius'])
    flx[0] = flx[0]**2
    flx[0] /= param['minor_radius']
    return np.sqrt(flx[0])


def plasma_bundles(param, stream, filters=()):
    """setup_bundles + bundle_filter + bundle_generate -> bundle table."""
    nb = param['bundle_count']
    b = {
        'origin': np.zeros([nb, 3]), 'temperature': np.ones([nb]),
        'emissivity': np.ones([nb]), 'velocity': np.zeros([nb, 3]),
        'mask': np.ones([nb], dtype=np.bool_), 'spread': np.zeros([nb]),
        'solid_angle': np.zeros([nb]),
    }
    off = np.zeros((nb, 3))
    off[:, 0] = stream.uniform(-1 * param['xsize'] / 2, param['xsize'] / 2, nb, site='plasma.center.0')
    off[:, 1] = stream.uniform(-1 * param['ysize'] / 2, param['ysize'] / 2, nb, site='plasma.center.1')
    off[:, 2] = stream.uniform(-1 * param['zsize'] / 2, param['zsize'] / 2, nb, site='plasma.center.2')
    b['origin'][:] = vecs.to_external(param['orientation'], off) + param['origin']

    if param['spread_radius'] is not None:
        dist = np.linalg.norm(b['origin'] - param['target'], axis=1)
        spread = np.arctan(param['spread_radius'] / dist)
    else:
        spread = param['spread']
    b['spread'][:] = spread
    for i in range(nb):
        b['solid_angle'][i] = solid_angle_isotropic(b['spread'][i])

    for f in filters:
        b = sightline(f, b)

    kind = param['_kind']
    if kind == 'plasma_cubic':
        b['temperature'][:] = param['temperature']
        b['emissivity'][:] = param['emissivity']
    elif kind in ('plasma_toroidal', 'plasma_datafile'):
        m = b['mask']
        pts = b['origin'][m]
        rho = np.zeros(len(pts))
        for i in range(len(pts)):
            rho[i] = plasma_rho(param, pts[i, :])
        if kind == 'plasma_datafile':
            def profile(fname):
                data = np.loadtxt(fname, dtype=np.float64)
                return np.interp(rho, data[:, 0], data[:, 1], left=0.0, right=0.0)
            temp = profile(param['temperature_file'])
            emis = profile(param['emissivity_file'])
        else:
            temp = param['temperature']
            emis = param['emissivity']
        b['temperature'][m] = temp * param['temperature_scale']
        b['emissivity'][m] = emis * param['emissivity_scale']
        b['velocity'][m] = param['velocity'] * param['velocity_scale']
        m &= np.isfinite(b['temperature'])
    return b


def bundle_intensity(param, b):
    """create_sources :301-319 -- expected photons per bundle (vector form)."""
    inten = (b['emissivity'] * param['time_resolution'] * param['bundle_volume']
             * b['solid_angle'] / (4 * np.pi))
    inten = inten * (param['volume'] / (param['bundle_count'] * param['bundle_volume']))
    return inten


def plasma_source(param, stream, filters=()):
    b = plasma_bundles(param, stream, filters)
    m = b['mask']
    inten = bundle_intensity(param, b)
    predicted = int(np.sum(inten[m]))
    if param['max_rays'] and predicted > param['max_rays']:
        raise ValueError(
            f"Current settings will produce too many rays ({predicted:0.2e}). "
            f"Please reduce integration time or adjust other parameters.")

    parts = []
    for i in range(param['bundle_count']):
        if not m[i]:
            continue
        cfg = {
            'class_name': 'XicsrtSourceFocused',
            'origin': b['origin'][i], 'temperature': b['temperature'][i],
            'velocity': b['velocity'][i], 'spread': b['spread'][i],
            'intensity': inten[i],
            'xsize': param['voxel_size'], 'ysize': param['voxel_size'],
            'zsize': param['voxel_size'], 'zaxis': param['zaxis'],
            'xaxis': param['xaxis'], 'target': param['target'],
            'mass_number': param['mass_number'],
            'wavelength_dist': param['wavelength_dist'],
            'wavelength': param['wavelength'],
            'wavelength_range': param['wavelength_range'],
            'linewidth': param['linewidth'],
            'angular_dist': param['angular_dist'],
            'use_poisson': param['use_poisson'],
        }
        _, sp = elements.prepare_source(cfg, poisson=lambda lam: stream.poisson(lam, site='plasma.count'))
        parts.append(box_source(sp, stream))

    total = int(np.sum([len(p['mask']) for p in parts])) if parts else 0
    if total == 0:
        raise ValueError('No rays generated. Check plasma input parameters')
    rays = {k: np.concatenate([p[k] for p in parts]) for k in ('origin', 'direction', 'wavelength', 'weight', 'mask')}
    rays['_bundle_counts'] = np.array([len(p['mask']) for p in parts])
    return rays, b
