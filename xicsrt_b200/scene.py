# -*- coding: utf-8 -*-
"""
Scene flattening: prepared element ``param`` dicts -> ``XrtSceneDesc`` (the POD
the kernels read from the constant bank) and the device scene handle.

Everything here is setup-time; the derived numbers are computed with the same
expressions the reference uses at its ``initialize`` time or at the top of its
per-call numerics (cited per field) so that the kernel constants are bit-equal
to the reference's temporaries.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import elements, rocking, voigt


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _set(arr, values):
    values = np.asarray(values, dtype=np.float64).ravel()
    for i, v in enumerate(values):
        arr[i] = float(v)


class _Keep:
    """Owns the numpy buffers the descriptor points into until the scene is uploaded."""

    def __init__(self):
        self.items = []

    def f64(self, a):
        a = _f64(a)
        self.items.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def arr(self, a, dtype, ctype):
        a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
        self.items.append(a)
        return a.ctypes.data_as(C.POINTER(ctype))

    def obj(self, o):
        self.items.append(o)
        return o

    def nbytes(self):
        """Bytes of the tables xrt_scene_create copies to the device (numpy buffers and ctypes tables)."""
        total = 0
        for it in self.items:
            if isinstance(it, np.ndarray):
                total += int(it.nbytes)
            else:
                try:
                    total += int(C.sizeof(it))
                except TypeError:
                    pass
        return total


# ---------------------------------------------------------------------------
# source

def _one(spread):
    s = np.atleast_1d(np.asarray(spread, dtype=np.float64))
    if s.size != 1:
        raise Exception('Spread must be a scalar or one element array.')
    return float(s[0])


def _four(spread):
    s = np.atleast_1d(np.asarray(spread, dtype=np.float64))
    if s.size == 1:
        return [-s[0], s[0], -s[0], s[0]]
    if s.size == 2:
        return [-s[0], s[0], -s[1], s[1]]
    if s.size == 4:
        return [s[0], s[1], s[2], s[3]]
    raise Exception('Spread must have 1, 2 or 3 elements. See docstring.')


def _norm3(v):
    """np.linalg.norm(v[None, :], axis=1)[0] for a 3-vector: the same sum order, bit-identical."""
    return np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def cone_basis(axis, xaxis, zaxis):
    """Rows (o_2, o_1, axis) of the cone frame, _XicsrtSourceGeneric.py:282-292 (norms over axis 1 there)."""
    from .elements import cross3
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / _norm3(axis)
    o1 = cross3(axis, xaxis) + cross3(axis, zaxis)
    o1 /= _norm3(o1)
    o2 = cross3(axis, o1)
    o2 /= _norm3(o2)
    return np.stack([o2, o1, axis])


def _fill_cone(src, name, spread):
    name = 'isotropic' if name is None else str(name).lower()
    if name == 'gaussian':
        # xicsrt_spread.py:55 calls an undefined function: NameError in the reference
        raise NotImplementedError('angular_dist "gaussian" is not implemented in the reference.')
    if name not in L.CONE:
        raise Exception(f'Distribution "{name}" is not known.')
    src.cone = L.CONE[name]
    if name == 'isotropic':       # xicsrt_spread.py:80-110
        _set(src.cone_par, [np.cos(_one(spread))])
    elif name == 'flat':          # :213-245
        _set(src.cone_par, [np.tan(_one(spread))])
    elif name == 'flat_xy':       # :247-294
        _set(src.cone_par, np.tan(_four(spread)))
    else:                         # isotropic_xy :130-196
        th = _four(spread)
        tx = np.max(np.abs(th[0:2]))
        ty = np.max(np.abs(th[2:]))
        theta_max = np.arcsin(np.sqrt(np.sin(tx)**2 + np.sin(ty)**2))
        _set(src.cone_par, np.sin(th))
        src.cone_cos_max = float(np.cos(theta_max))


def _fill_wavelength(src, param, keep):
    model = voigt.wavelength_model(param)
    src.wave = L.WAVE[model['mode']]
    if model['mode'] == 'const':
        _set(src.wave_par, [model['wavelength']])
    elif model['mode'] == 'uniform':
        _set(src.wave_par, [model['lo'], model['hi']])
    elif model['mode'] == 'normal':
        _set(src.wave_par, [model['wavelength'], model['sigma']])
    else:
        cdf, x = model['cdf'], model['x']
        _set(src.wave_par, [model['wavelength'], np.min(cdf), np.max(cdf)])
        src.n_table = len(cdf)
        src.table_cdf = keep.f64(cdf)
        src.table_x = keep.f64(x)
    return model


def fill_source(src, param, filters, keep, bundles=None):
    """Box sources (generic / directed / focused) and plasma bundle tables."""
    kind = param['_kind']
    _set(src.origin, param['origin'])
    _set(src.orient, param['orientation'])
    _set(src.velocity_c, [0.0, 0.0, 0.0])

    spatial = 'uniform' if kind.startswith('plasma') else str(param['spatial_dist']).lower()
    if spatial not in L.SPATIAL:
        raise NotImplementedError(f"spatial_dist: {param['spatial_dist']} not implemented.")
    src.spatial = L.SPATIAL[spatial]

    if kind.startswith('plasma'):
        src.kind = L.SRC_BUNDLES
        _set(src.target, param['target'])
        from . import plasma as xplasma
        # per-bundle cone parameter (XrtBundle.cos_spread) and Doppler sigma come from the bundle table
        src.cone = L.CONE[xplasma.cone_kind(param)]
        line = xplasma.line_model(param)
        if line == 'table':
            # natural linewidth: every bundle samples its own table (xrt_bundle_voigt_tables)
            src.wave = L.WAVE['table']
            _set(src.wave_par, [float(param['wavelength'])])
            src.n_table = xplasma.N_TABLE
        else:
            wparam = dict(param)
            wparam['temperature'] = 1.0
            _fill_wavelength(src, wparam, keep)
        src.voxel_size = float(param['voxel_size'])
        if bundles is not None:          # host-built table (tests); the driver builds it on the device
            src.n_bundles = len(bundles['end'])
            table = keep.obj(np.ascontiguousarray(bundles['table']))
            src.bundles = C.cast(table.ctypes.data, C.POINTER(L.XrtBundle))
            src.bundle_end = keep.arr(bundles['end'], np.uint64, C.c_uint64)
            if line == 'table':
                src.bundle_x = keep.f64(bundles['voigt_x'].ravel())
                src.bundle_cdf = keep.f64(bundles['voigt_cdf'].ravel())
        _set(src.extent, [param['voxel_size']] * 3)
    else:
        if spatial == 'uniform':
            _set(src.extent, [param['xsize'], param['ysize'], param['zsize']])
        else:
            k = 2 * np.sqrt(2 * np.log(2))
            _set(src.extent, [param['xsize'] / k, param['ysize'] / k, param['zsize'] / k])
        _fill_cone(src, param['angular_dist'], param['spread'])
        _fill_wavelength(src, param, keep)
        vel = np.asarray(param['velocity'], dtype=np.float64)
        _set(src.velocity_c, vel / voigt.C_LIGHT)
        if kind == 'focused':
            src.kind = L.SRC_FOCUSED
            _set(src.target, param['target'])
        else:
            src.kind = L.SRC_FIXED_AXIS
            axis = param['direction'] if kind == 'directed' else param['zaxis']
            _set(src.axis_basis, cone_basis(axis, param['xaxis'], param['zaxis']))

    sight = [f for f in filters if f['_kind'] == 'sightline']
    if kind.startswith('plasma'):
        sight = []     # bundle filters act on the bundle table (host side), not on rays
    if len(sight) > L.MAX_SIGHTLINES:
        raise NotImplementedError(f'more than {L.MAX_SIGHTLINES} sightline filters on one source')
    src.n_sightlines = len(sight)
    for i, f in enumerate(sight):
        _set(src.sightlines[i].origin, f['origin'])
        _set(src.sightlines[i].axis, f['zaxis'])
        src.sightlines[i].radius = float(f['radius'])


# ---------------------------------------------------------------------------
# optics

def aperture_list(aperture):
    if aperture is None:
        return []
    if isinstance(aperture, dict):
        return [aperture]
    return list(np.atleast_1d(np.asarray(aperture, dtype=object)))


def _fill_apertures(param, keep):
    aps = aperture_list(param['aperture'])
    if not aps:
        return None, 0
    table = (L.XrtAperture * len(aps))()
    for i, ap in enumerate(aps):
        shape = (ap.get('shape') or 'none').lower()
        logic = (ap.get('logic') or 'and').lower()
        if shape not in L.AP_SHAPE:
            raise Exception(f'Aperture shape: "{shape}" is not implemented.')
        if logic not in L.AP_LOGIC:
            raise Exception(f'Aperture logic "{logic}" is not known.')
        org = ap.get('origin')
        org = np.array([0.0, 0.0]) if org is None else np.atleast_1d(np.asarray(org, dtype=np.float64))
        table[i].shape = L.AP_SHAPE[shape]
        table[i].logic = L.AP_LOGIC[logic]
        _set(table[i].origin, org[0:2])
        if 'size' in ap and shape != 'none':
            size = np.atleast_1d(np.asarray(ap['size'], dtype=np.float64))
            need = 2 if shape in ('rectangle', 'ellipse') else 1
            if shape != 'triangle':
                if size.size < need:
                    raise Exception(f'Aperture "{shape}" needs {need} size value(s).')
                _set(table[i].size, size[0:need])
        if shape == 'triangle':
            v = np.asarray(ap['vertices'], dtype=np.float64)
            _set(table[i].vert, (v[0:3, 0:2] + org[None, 0:2]).ravel())
    keep.obj(table)
    return C.cast(table, C.POINTER(L.XrtAperture)), len(aps)


def fill_optic(op, param, keep, image_offset):
    """One optic; returns the number of image pixels it owns."""
    shape, interact = param['_shape'], param['_interact']
    op.shape = L.SHAPE['mesh' if shape.startswith('mesh') else shape]
    op.interact = L.INTERACT[interact]

    flags = 0
    if param['trace_local']:
        flags |= L.F_TRACE_LOCAL
    if param['check_size']:
        flags |= L.F_CHECK_SIZE
    if param['check_aperture']:
        flags |= L.F_CHECK_APERTURE
    for ax, (key, flag) in enumerate((('xsize', L.F_HAS_XSIZE), ('ysize', L.F_HAS_YSIZE), ('zsize', L.F_HAS_ZSIZE))):
        if param[key] is not None:
            flags |= flag
            op.half_size[ax] = float(param[key]) / 2     # strict |x| < size/2, _TraceObject.py:194-214
    _set(op.origin, param['origin'])
    _set(op.orient, param['orientation'])

    if shape in ('sphere', 'cylinder'):
        _set(op.center, param['center'])
        op.radius = float(param['radius'])
        if param['convex']:
            flags |= L.F_CONVEX
    elif shape == 'torus':
        _set(op.center, param['center'])
        op.torus_major = float(param['torus_major'])
        op.torus_minor = float(param['torus_minor'])
        op.root_idx = int(param['root_idx'])
    elif shape.startswith('mesh'):
        from . import mesh
        op.mesh, mflags = mesh.fill_mesh(param, keep)
        flags |= mflags

    if interact in ('crystal', 'mosaic'):
        if param['check_bragg'] is not False:          # identity test, as _InteractCrystal.py:122
            flags |= L.F_CHECK_BRAGG
        op.two_d = 2 * float(param['crystal_spacing'])
        op.inv_two_d = 1.0 / op.two_d if op.two_d != 0.0 else float('inf')
        op.reflectivity = float(param['reflectivity'])
        op.rocking_mix = float(param['rocking_mix'])
        rtype = param['rocking_type']
        if flags & L.F_CHECK_BRAGG:
            if 'step' in rtype:
                op.rocking_type = L.ROCK['step']
                op.rocking_fwhm = float(param['rocking_fwhm'])
            elif 'gauss' in rtype:
                op.rocking_type = L.ROCK['gauss']
                op.rocking_fwhm = float(param['rocking_fwhm'])
                sigma = param['rocking_fwhm'] / (2 * np.sqrt(2 * np.log(2)))
                op.rock_two_sigma2 = float(2 * sigma**2)
                op.rock_inv_two_sigma2 = 1.0 / op.rock_two_sigma2 if op.rock_two_sigma2 != 0.0 else float('inf')
            elif 'file' in rtype:
                op.rocking_type = L.ROCK['table']
                tab = rocking.load_table(param['rocking_file'], param['rocking_filetype'])
                op.n_rock = len(tab['dtheta'])
                op.rock_dtheta = keep.f64(tab['dtheta'])
                op.rock_s = keep.f64(tab['reflect_s'])
                op.rock_p = keep.f64(tab['reflect_p'])
            else:
                raise Exception('Rocking curve type not understood: {}'.format(rtype))
    if interact == 'mosaic':
        op.mosaic_depth = int(param['mosaic_depth'])
        op.mosaic_spread = float(param['mosaic_spread'])
        hwhm = param['mosaic_spread'] / 2.0
        op.mosaic_sin_sigma = float(np.sin(hwhm / np.sqrt(2 * np.log(2))))   # xicsrt_spread.py:318-323
        if param['mosaic_cutoff'] is not None:
            flags |= L.F_MOSAIC_CUTOFF
            sig = param['mosaic_spread'] / (2 * np.sqrt(2 * np.log(2)))
            op.mosaic_angle_cut = float(np.sqrt(-1 * np.log(param['mosaic_cutoff']) * 2 * sig**2))

    if param['check_aperture']:
        op.apertures, op.n_aperture = _fill_apertures(param, keep)

    npix = 0
    if param['enable_image']:
        flags |= L.F_IMAGE
        op.npix[0] = int(param['pixel_xsize'])
        op.npix[1] = int(param['pixel_ysize'])
        op.pixel_size = float(param['pixel_size'])
        op.image_offset = int(image_offset)
        npix = op.npix[0] * op.npix[1]
    op.flags = flags
    return npix


# ---------------------------------------------------------------------------

class SceneLayout:
    """Names, ray count and image layout of a flattened scene (host-side bookkeeping)."""

    def __init__(self):
        self.source_name = None
        self.optic_names = []
        self.n_rays = 0
        self.images = {}        # optic name -> (offset, nx, ny) or None
        self.n_pixels = 0

    @property
    def element_names(self):
        return [self.source_name] + list(self.optic_names)


def flatten(source_name, source_param, source_filters, optics, bundles=None):
    """
    Build the descriptor.  ``optics`` is an ordered {name: param} dict.
    Returns (desc, layout, keep); ``keep`` must outlive the xrt_scene_create call.
    """
    if len(optics) > L.MAX_OPTICS:
        raise NotImplementedError(f'more than {L.MAX_OPTICS} optics in one scene')
    keep = _Keep()
    desc = L.XrtSceneDesc()
    desc.version = L.XRT_VERSION
    desc.n_optics = len(optics)
    layout = SceneLayout()
    layout.source_name = source_name
    fill_source(desc.source, source_param, source_filters, keep, bundles=bundles)
    if bundles is not None:
        layout.n_rays = int(bundles['end'][-1])
    else:
        layout.n_rays = int(source_param.get('intensity', 0))     # plasma: set per iteration by the driver
    offset = 0
    for k, (name, param) in enumerate(optics.items()):
        npix = fill_optic(desc.optics[k], param, keep, offset)
        layout.optic_names.append(name)
        layout.images[name] = (offset, int(param['pixel_xsize']), int(param['pixel_ysize'])) if npix else None
        offset += npix
    layout.n_pixels = offset
    return desc, layout, keep


class DeviceScene:
    """RAII handle of an uploaded scene (xrt_scene_create / xrt_scene_destroy)."""

    def __init__(self, desc, layout):
        self.lib = L.load()
        self.layout = layout
        handle = C.c_void_p()
        L.check(self.lib.xrt_scene_create(C.byref(desc), C.byref(handle)))
        self.handle = handle

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.xrt_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bundles(self, table, end, n_rays=0):
        """Point a plasma scene at a device-resident bundle table (torch tensors, kept alive by the caller)."""
        L.check(self.lib.xrt_scene_set_bundles(self.handle, table.data_ptr(), end.data_ptr(), int(end.numel()),
                                               int(n_rays)))

    def set_bundle_tables(self, x, cdf):
        """Per-bundle wavelength tables [n_bundles, n_table] of a plasma with a natural linewidth."""
        L.check(self.lib.xrt_scene_set_bundle_tables(self.handle, x.data_ptr(), cdf.data_ptr(), int(x.shape[1])))

    def launch_info(self):
        g, b, r, p = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        L.check(self.lib.xrt_launch_info(self.handle, C.byref(g), C.byref(b), C.byref(r), C.byref(p)))
        info = {'grid': g.value, 'block': b.value, 'registers': r.value, 'blocks_per_sm': p.value}
        m = C.c_int32()
        L.check(self.lib.xrt_launch_info_cull(self.handle, C.byref(m), C.byref(g), C.byref(r), C.byref(p)))
        # FP32 broad phase in front of the fused kernel: None when it does not apply to this scene
        info['broad_phase'] = None if m.value < 0 or m.value > 3 else {
            'source_kind': ('point', 'box', 'focused', 'bundles')[m.value], 'grid': g.value, 'registers': r.value,
            'blocks_per_sm': p.value}
        # sorted mesh path (csrc/xrt_meshsort.cuh) instead of the fused kernel for launches above 2^21 rays
        info['mesh_sort'] = None if m.value != 4 else {
            'coarse': {'grid': g.value, 'registers': r.value & 0xffff, 'blocks_per_sm': p.value & 0xff},
            'refine_registers': r.value >> 16, 'bins': p.value >> 8}
        # FP32 broad phase of a mosaic crystal's crystallite scan (k_mosaic32) in front of the fused kernel
        info['mosaic_broad_phase'] = None if m.value != 5 else {'grid': g.value, 'registers': r.value, 'blocks_per_sm': p.value}
        return info


def prepare(config, poisson=None):
    """
    Element preparation for one run: returns (config_out, source_name, source_param,
    source_filters, optics) with the fully defaulted configs written back into
    ``config`` the way the reference's Dispatchers do (xicsrt_raytrace.py:123-149).
    """
    strict = config['general']['strict_config_check']
    filters = {}
    cfg_filters = {}
    for name, c in (config.get('filters') or {}).items():
        cfg_filters[name], filters[name] = elements.prepare_filter(c, strict=strict)
    if 'filters' in config:
        config['filters'] = cfg_filters

    if len(config['sources']) == 0:
        raise Exception('No ray sources defined.')
    if len(config['sources']) != 1:
        raise NotImplementedError('Multiple ray sources are not currently supported.')
    cfg_sources = {}
    source_name = source_param = None
    for name, c in config['sources'].items():
        cfg_sources[name], source_param = elements.prepare_source(c, strict=strict, poisson=poisson)
        source_name = name
    config['sources'] = cfg_sources
    wanted = source_param.get('filters')
    source_filters = [filters[f] for f in filters if wanted is not None and f in wanted]

    optics = {}
    cfg_optics = {}
    for name, c in config['optics'].items():
        cfg_optics[name], optics[name] = elements.prepare_optic(c, strict=strict)
    config['optics'] = cfg_optics
    return config, source_name, source_param, source_filters, optics
