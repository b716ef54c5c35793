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


N_TABLE = 1000          # bins of the reference's Voigt table (xicsrt_voigt.py:47)


def cone_kind(param):
    """angular_dist of the per-bundle sources -> XRT_CONE_* (gaussian is a NameError in the reference)."""
    name = 'isotropic' if param['angular_dist'] is None else str(param['angular_dist']).lower()
    if name == 'gaussian':
        raise NotImplementedError('angular_dist "gaussian" is not implemented in the reference.')
    if name not in L.CONE:
        raise Exception(f'Distribution "{name}" is not known.')
    return name


def cone_parameter(name, spread):
    """What XrtBundle.cos_spread holds for a bundle's scalar spread (see include/xrt.h)."""
    spread = np.asarray(spread, dtype=np.float64)
    return {'isotropic': np.cos, 'isotropic_xy': np.sin, 'flat': np.tan, 'flat_xy': np.tan}[name](spread)


def line_model(param):
    """
    'const' | 'uniform' | 'normal' | 'table' for the per-bundle sources: the branch order of
    _XicsrtSourceGeneric.py:295-354 with a per-bundle temperature.  With a natural linewidth
    every bundle samples its own tabulated Voigt profile.
    """
    wtype = str(param['wavelength_dist']).lower()
    if wtype == 'monochrome':
        return 'const'
    if wtype == 'uniform':
        return 'uniform'
    if wtype != 'voigt':
        raise Exception(f'Wavelength distribution {wtype} unknown')
    return 'table' if float(param['linewidth']) != 0.0 else 'normal'


def bundle_sigma(param, temperature):
    """Doppler sigma per bundle; with a natural linewidth T == 0 becomes 1 eV (:333-339)."""
    model = line_model(param)
    temp = np.asarray(temperature, dtype=np.float64)
    if model not in ('normal', 'table'):
        return np.zeros(len(temp))
    if model == 'table':
        temp = np.where(temp == 0.0, 1.0, temp)
    return np.where(temp > 0, voigt.doppler_sigma(np.abs(temp), param['mass_number'], float(param['wavelength'])), 0.0)


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
    cone = cone_kind(param)
    model = line_model(param)
    if param['target'] is None:
        raise ValueError('plasma sources need a target')
    b = bundle_properties(param, filters, rng)
    counts = bundle_counts(param, b, rng)
    keep = counts > 0
    n = int(np.sum(keep))
    if n == 0:
        raise ValueError('No rays generated. Check plasma input parameters')

    sigma = bundle_sigma(param, b['temperature'][keep])

    rec = np.zeros(n, dtype=[('origin', 'f8', 3), ('cos_spread', 'f8'), ('wave_sigma', 'f8'), ('velocity_c', 'f8', 3)])
    rec['origin'] = b['origin'][keep]
    rec['cos_spread'] = cone_parameter(cone, b['spread'][keep])
    rec['wave_sigma'] = sigma
    rec['velocity_c'] = b['velocity'][keep] / voigt.C_LIGHT
    assert rec.dtype.itemsize == C.sizeof(L.XrtBundle)
    table = np.ascontiguousarray(rec)
    end = np.cumsum(counts[keep]).astype(np.uint64)
    out = {'table': table, 'end': end, 'n_rays': int(end[-1]), 'props': b, 'counts': counts}
    if model == 'table':
        # one table per emitting bundle, as each per-bundle source builds (host restatement with scipy's wofz)
        gamma = float(voigt.natural_gamma(float(param['linewidth']), float(param['wavelength'])))
        tabs = [voigt.cdf_table(gamma, float(sg), gridsize=N_TABLE) for sg in sigma]
        out['voigt_x'] = np.ascontiguousarray([t[0] for t in tabs], dtype=np.float64)
        out['voigt_cdf'] = np.ascontiguousarray([t[1] for t in tabs], dtype=np.float64)
    return out


# ---------------------------------------------------------------------------
# device path: the same table built by libxrt (xrt_bundles_generate)

def plasma_desc(param, filters, profiles=None, inject_u=None):
    """
    ``XrtPlasmaDesc`` of a prepared plasma source.  ``profiles`` = dict of device tensors
    (t_rho, t_val, e_rho, e_val) for the datafile class; ``inject_u`` = optional device tensor
    [3, n] of centre uniforms (parity tests).
    """
    model = line_model(param)
    if param['target'] is None:
        raise ValueError('plasma sources need a target')
    d = L.XrtPlasmaDesc()
    d.kind = L.PLASMA[param['_kind']]
    d.use_poisson = 1 if param['use_poisson'] else 0
    d.use_spread_radius = 1 if param['spread_radius'] is not None else 0
    d.thermal_line = {'normal': 1, 'table': 2}.get(model, 0)
    d.cone = L.CONE[cone_kind(param)]
    for i in range(3):
        d.origin[i] = float(param['origin'][i])
        d.target[i] = float(param['target'][i])
    for i, v in enumerate(np.asarray(param['orientation'], dtype=np.float64).ravel()):
        d.orient[i] = float(v)
    d.size[0], d.size[1], d.size[2] = float(param['xsize']), float(param['ysize']), float(param['zsize'])
    d.spread = 0.0 if param['spread'] is None else float(np.atleast_1d(param['spread'])[0])
    d.spread_radius = 0.0 if param['spread_radius'] is None else float(param['spread_radius'])
    kind = param['_kind']
    if kind in ('plasma_cubic', 'plasma_toroidal'):
        d.temperature = float(param['temperature'])
        d.emissivity = float(param['emissivity'])
    if kind in ('plasma_toroidal', 'plasma_datafile'):
        vel = np.broadcast_to(np.asarray(param['velocity'], dtype=np.float64), (3,))
        d.velocity[0], d.velocity[1], d.velocity[2] = (float(v) for v in vel)
        d.temperature_scale = float(param['temperature_scale'])
        d.emissivity_scale = float(param['emissivity_scale'])
        d.velocity_scale = float(param['velocity_scale'])
        d.major_radius, d.minor_radius = float(param['major_radius']), float(param['minor_radius'])
        for i in range(3):
            d.torus_origin[i] = float(np.asarray(param['torus_origin'], dtype=np.float64)[i])
    if kind == 'plasma_datafile':
        d.n_profile_t, d.n_profile_e = int(profiles['t_rho'].numel()), int(profiles['e_rho'].numel())
        d.profile_t_rho, d.profile_t_val = profiles['t_rho'].data_ptr(), profiles['t_val'].data_ptr()
        d.profile_e_rho, d.profile_e_val = profiles['e_rho'].data_ptr(), profiles['e_val'].data_ptr()
    d.intensity_factor = float(param['time_resolution'] * param['bundle_volume'] / (4 * np.pi)
                               * (param['volume'] / (param['bundle_count'] * param['bundle_volume'])))
    d.sigma_factor = float(voigt.doppler_sigma(1.0, param['mass_number'], float(param['wavelength'])))
    d.inv_c = 1.0 / voigt.C_LIGHT
    sight = [f for f in filters if f['_kind'] == 'sightline']
    if len(sight) > L.MAX_SIGHTLINES:
        raise NotImplementedError(f'more than {L.MAX_SIGHTLINES} sightline filters on one source')
    d.n_sightlines = len(sight)
    for k, f in enumerate(sight):
        for i in range(3):
            d.sightlines[k].origin[i] = float(f['origin'][i])
            d.sightlines[k].axis[i] = float(f['zaxis'][i])
        d.sightlines[k].radius = float(f['radius'])
    if inject_u is not None:
        d.inject_u = inject_u.data_ptr()
    return d


def load_profiles(param, torch, device):
    """rho -> value tables of the datafile class as device tensors (two-column text files)."""
    out = {}
    for tag, key in (('t', 'temperature_file'), ('e', 'emissivity_file')):
        data = np.loadtxt(param[key], dtype=np.float64)
        out[f'{tag}_rho'] = torch.from_numpy(np.ascontiguousarray(data[:, 0])).to(device)
        out[f'{tag}_val'] = torch.from_numpy(np.ascontiguousarray(data[:, 1])).to(device)
    return out


class DeviceBundles:
    """The bundle table of one iteration, resident on the device (torch tensors)."""

    def __init__(self, torch, device, param, filters, lib):
        self.torch, self.device, self.param, self.lib = torch, device, param, lib
        self.n = int(param['bundle_count'])
        self.profiles = load_profiles(param, torch, device) if param['_kind'] == 'plasma_datafile' else None
        self.filters = filters
        self.table = torch.empty((self.n, 8), dtype=torch.float64, device=device)     # XrtBundle = 8 doubles
        self.intensity = torch.empty(self.n, dtype=torch.float64, device=device)
        self.counts = torch.empty(self.n, dtype=torch.int64, device=device)
        self.end = None
        assert C.sizeof(L.XrtBundle) == 64
        # natural linewidth: one Voigt inverse-CDF table per bundle, built on the device each iteration
        self.voigt_x = self.voigt_cdf = None
        if line_model(param) == 'table':
            need = 2 * 8 * N_TABLE * self.n
            free = torch.cuda.mem_get_info(device)[0]
            if need > 0.8 * free:
                raise MemoryError(f'per-bundle Voigt tables need {need / 2**30:.1f} GiB for {self.n} bundles '
                                  f'({free / 2**30:.1f} GiB free): reduce bundle_count')
            self.gamma = float(voigt.natural_gamma(float(param['linewidth']), float(param['wavelength'])))
            self.voigt_x = torch.empty((self.n, N_TABLE), dtype=torch.float64, device=device)
            self.voigt_cdf = torch.empty((self.n, N_TABLE), dtype=torch.float64, device=device)

    def generate(self, seed, stream_id, inject_u=None):
        """Build the table for (seed, stream_id); returns the total number of rays."""
        torch = self.torch
        desc = plasma_desc(self.param, self.filters, self.profiles, inject_u)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            L.check(self.lib.xrt_bundles_generate(C.byref(desc), int(seed), int(stream_id), self.n,
                                                  self.table.data_ptr(), self.intensity.data_ptr(),
                                                  self.counts.data_ptr(), stream))
            if self.voigt_x is not None:
                L.check(self.lib.xrt_bundle_voigt_tables(self.table.data_ptr(), self.counts.data_ptr(), self.n,
                                                         self.gamma, N_TABLE, self.voigt_x.data_ptr(),
                                                         self.voigt_cdf.data_ptr(), stream))
        kept = self.intensity >= 0
        inten = torch.where(kept, self.intensity, torch.zeros_like(self.intensity))
        # one device->host read for the three numbers the host needs
        self.end = torch.cumsum(self.counts, 0)
        lowest = torch.where(kept, self.intensity, torch.full_like(self.intensity, float('inf'))).min()
        stats = torch.stack([inten.sum(), lowest, self.end[-1].to(torch.float64)]).cpu().numpy()
        predicted, lowest, total = int(stats[0]), float(stats[1]), int(stats[2])
        if self.param['max_rays'] and predicted > self.param['max_rays']:
            raise ValueError(
                f"Current settings will produce too many rays ({predicted:0.2e}). "
                f"Please reduce integration time or adjust other parameters.")
        if not self.param['use_poisson'] and lowest < 1:
            raise ValueError('intensity of less than one encountered. Turn on poisson statistics.')
        if total == 0:
            raise ValueError('No rays generated. Check plasma input parameters')
        return total
