# -*- coding: utf-8 -*-
"""
ctypes binding of ``libxrt.so`` -- a field-for-field mirror of include/xrt.h.

There is no CPU path behind these calls: if the library cannot be loaded the
import of the caller fails with the loader's error, and every entry point
that needs a device returns ``XRT_ECUDA`` (raised here as :class:`XrtError`)
when none is present.
"""
import ctypes as C
import os

XRT_VERSION = 1
MAX_OPTICS = 16
MAX_SIGHTLINES = 4

OK, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM = 0, -1, -2, -3, -4

SHAPE = {'plane': 0, 'sphere': 1, 'cylinder': 2, 'torus': 3, 'mesh': 4}
INTERACT = {'none': 0, 'mirror': 1, 'crystal': 2, 'mosaic': 3}
ROCK = {'step': 0, 'gauss': 1, 'table': 2}
F_TRACE_LOCAL, F_CHECK_SIZE, F_CHECK_APERTURE, F_CHECK_BRAGG = 1 << 0, 1 << 1, 1 << 2, 1 << 3
F_CONVEX, F_HAS_XSIZE, F_HAS_YSIZE, F_HAS_ZSIZE = 1 << 4, 1 << 5, 1 << 6, 1 << 7
F_IMAGE, F_MOSAIC_CUTOFF, F_MESH_REFINE, F_MESH_INTERP = 1 << 8, 1 << 9, 1 << 10, 1 << 11
F_MESH_LOSSLESS = 1 << 12
AP_SHAPE = {'none': 0, 'circle': 1, 'square': 2, 'rectangle': 3, 'ellipse': 4, 'triangle': 5}
AP_LOGIC = {'and': 0, 'not': 1, 'or': 2, 'nand': 3, 'nor': 4, 'xor': 5, 'xnor': 6}
SRC_FIXED_AXIS, SRC_FOCUSED, SRC_BUNDLES = 0, 1, 2
SPATIAL = {'uniform': 0, 'gaussian': 1}
CONE = {'isotropic': 0, 'isotropic_xy': 1, 'flat': 2, 'flat_xy': 3}
WAVE = {'const': 0, 'uniform': 1, 'normal': 2, 'table': 3}

_pd = C.POINTER(C.c_double)
_pi32 = C.POINTER(C.c_int32)
_pu8 = C.POINTER(C.c_uint8)
_pu64 = C.POINTER(C.c_uint64)


class XrtAperture(C.Structure):
    _fields_ = [('shape', C.c_int32), ('logic', C.c_int32), ('origin', C.c_double * 2),
                ('size', C.c_double * 2), ('vert', C.c_double * 6)]


class XrtMesh(C.Structure):
    _fields_ = [('n_points', C.c_int32), ('n_faces', C.c_int32),
                ('points', _pd), ('faces', _pi32), ('face_normals', _pd), ('face_geom', _pd), ('face_area', _pd), ('face_rec', _pd),
                ('n_coarse_points', C.c_int32), ('n_coarse_faces', C.c_int32),
                ('coarse_points', _pd), ('coarse_faces', _pi32), ('coarse_geom', _pd),
                ('point_faces', _pi32), ('point_faces_mask', _pu8), ('vertex_faces', _pi32),
                ('n_tri', C.c_int32), ('pad0', C.c_int32),
                ('ct_coef', _pd), ('tri_transform', _pd),
                ('grid_nx', C.c_int32), ('grid_ny', C.c_int32),
                ('grid_x0', C.c_double), ('grid_y0', C.c_double),
                ('grid_inv_dx', C.c_double), ('grid_inv_dy', C.c_double),
                ('grid_start', _pi32), ('grid_items', _pi32),
                ('vgrid_start', _pi32), ('vgrid_items', _pi32), ('vgrid_xyz', _pd),
                ('nb_start', _pi32), ('nb_rec', _pd), ('tri_rec', _pd), ('vertex_face_rec', _pd),
                ('fgrid_nx', C.c_int32), ('fgrid_ny', C.c_int32),
                ('fgrid_x0', C.c_double), ('fgrid_y0', C.c_double), ('fgrid_inv_dx', C.c_double), ('fgrid_inv_dy', C.c_double),
                ('fgrid_z_min', C.c_double), ('fgrid_z_max', C.c_double),
                ('fgrid_start', _pi32), ('fgrid_items', _pi32)]


class XrtOpticDesc(C.Structure):
    _fields_ = [('shape', C.c_int32), ('interact', C.c_int32), ('rocking_type', C.c_int32), ('flags', C.c_uint32),
                ('origin', C.c_double * 3), ('orient', C.c_double * 9), ('half_size', C.c_double * 3),
                ('center', C.c_double * 3), ('radius', C.c_double),
                ('torus_major', C.c_double), ('torus_minor', C.c_double),
                ('root_idx', C.c_int32), ('mosaic_depth', C.c_int32),
                ('two_d', C.c_double), ('inv_two_d', C.c_double), ('reflectivity', C.c_double), ('rocking_fwhm', C.c_double),
                ('rock_two_sigma2', C.c_double), ('rock_inv_two_sigma2', C.c_double), ('rocking_mix', C.c_double),
                ('mosaic_spread', C.c_double), ('mosaic_sin_sigma', C.c_double), ('mosaic_angle_cut', C.c_double),
                ('n_aperture', C.c_int32), ('n_rock', C.c_int32),
                ('apertures', C.POINTER(XrtAperture)),
                ('rock_dtheta', _pd), ('rock_s', _pd), ('rock_p', _pd),
                ('mesh', C.POINTER(XrtMesh)),
                ('npix', C.c_int32 * 2), ('pixel_size', C.c_double), ('image_offset', C.c_uint64),
                ('cull_t2', C.c_double), ('cull_err', C.c_double), ('cull_inv_r', C.c_double),
                ('mosaic_scan', C.c_int32), ('pad2', C.c_int32), ('mosaic_t2', C.c_double), ('mosaic_err', C.c_double)]


class XrtSightline(C.Structure):
    _fields_ = [('origin', C.c_double * 3), ('axis', C.c_double * 3), ('radius', C.c_double)]


class XrtBundle(C.Structure):
    _fields_ = [('origin', C.c_double * 3), ('cos_spread', C.c_double), ('wave_sigma', C.c_double),
                ('velocity_c', C.c_double * 3)]


PLASMA = {'plasma_generic': 0, 'plasma_cubic': 1, 'plasma_toroidal': 2, 'plasma_datafile': 3}


class XrtPlasmaDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('use_poisson', C.c_int32), ('use_spread_radius', C.c_int32),
                ('n_sightlines', C.c_int32), ('n_profile_t', C.c_int32), ('n_profile_e', C.c_int32),
                ('thermal_line', C.c_int32), ('cone', C.c_int32),
                ('origin', C.c_double * 3), ('orient', C.c_double * 9), ('size', C.c_double * 3),
                ('target', C.c_double * 3), ('spread', C.c_double), ('spread_radius', C.c_double),
                ('temperature', C.c_double), ('emissivity', C.c_double), ('velocity', C.c_double * 3),
                ('temperature_scale', C.c_double), ('emissivity_scale', C.c_double), ('velocity_scale', C.c_double),
                ('major_radius', C.c_double), ('minor_radius', C.c_double), ('torus_origin', C.c_double * 3),
                ('intensity_factor', C.c_double), ('sigma_factor', C.c_double), ('inv_c', C.c_double),
                ('profile_t_rho', C.c_void_p), ('profile_t_val', C.c_void_p),
                ('profile_e_rho', C.c_void_p), ('profile_e_val', C.c_void_p),
                ('inject_u', C.c_void_p),
                ('sightlines', XrtSightline * MAX_SIGHTLINES)]


class XrtSourceDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('spatial', C.c_int32), ('cone', C.c_int32), ('wave', C.c_int32),
                ('origin', C.c_double * 3), ('orient', C.c_double * 9), ('extent', C.c_double * 3),
                ('axis_basis', C.c_double * 9), ('target', C.c_double * 3),
                ('cone_par', C.c_double * 4), ('cone_cos_max', C.c_double),
                ('wave_par', C.c_double * 4), ('velocity_c', C.c_double * 3),
                ('n_table', C.c_int32), ('n_sightlines', C.c_int32),
                ('table_cdf', _pd), ('table_x', _pd),
                ('sightlines', XrtSightline * MAX_SIGHTLINES),
                ('n_bundles', C.c_uint64), ('bundles', C.POINTER(XrtBundle)), ('bundle_end', _pu64),
                ('voxel_size', C.c_double), ('bundle_x', _pd), ('bundle_cdf', _pd),
                ('bundle_hint', C.c_void_p), ('bundle_hint_shift', C.c_int32), ('pad1', C.c_int32)]


class XrtSceneDesc(C.Structure):
    _fields_ = [('version', C.c_int32), ('n_optics', C.c_int32), ('source', XrtSourceDesc),
                ('optics', XrtOpticDesc * MAX_OPTICS), ('kn32', C.c_float * 32)]


class XrtOutputs(C.Structure):
    _fields_ = [('counts', C.c_void_p), ('images', C.c_void_p),
                ('found_ids', C.c_void_p), ('found_count', C.c_void_p), ('found_capacity', C.c_uint64),
                ('lost_ids', C.c_void_p), ('lost_keys', C.c_void_p), ('lost_count', C.c_void_p),
                ('lost_capacity', C.c_uint64), ('lost_threshold', C.c_uint64),
                ('found_bits', C.c_void_p), ('lost_bits', C.c_void_p), ('bits_begin', C.c_uint64)]


class XrtHistory(C.Structure):
    _fields_ = [('rays', C.c_void_p), ('mask', C.c_void_p), ('capacity', C.c_uint64), ('layout', C.c_int32), ('pad0', C.c_int32)]


HIST_PLANES, HIST_ROWS = 0, 1


class XrtRaysIn(C.Structure):
    _fields_ = [('origin', C.c_void_p), ('direction', C.c_void_p), ('wavelength', C.c_void_p),
                ('mask', C.c_void_p)]


class XrtInject(C.Structure):
    _fields_ = [('u', C.c_void_p * MAX_OPTICS), ('xy', C.c_void_p * MAX_OPTICS)]


class XrtSourceInject(C.Structure):
    _fields_ = [('origin', C.c_void_p), ('cone', C.c_void_p), ('wave', C.c_void_p)]


class XrtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libxrt error {code}: {message}')
        self.code = code


LIB_PATH = os.environ.get('XRT_LIB_PATH') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libxrt.so')

# every symbol include/xrt.h declares: (restype, argtypes)
_u64, _vp = C.c_uint64, C.c_void_p
SYMBOLS = {
    'xrt_version': (C.c_int, []),
    'xrt_last_error': (C.c_char_p, []),
    'xrt_scene_create': (C.c_int, [C.POINTER(XrtSceneDesc), C.POINTER(_vp)]),
    'xrt_scene_destroy': (C.c_int, [_vp]),
    'xrt_trace': (C.c_int, [_vp, _u64, _u64, _u64, _u64, C.POINTER(XrtOutputs), _vp]),
    'xrt_trace_history': (C.c_int, [_vp, _u64, _u64, _vp, _u64, _u64, C.POINTER(XrtHistory), _vp]),
    'xrt_trace_injected': (C.c_int, [_vp, C.POINTER(XrtRaysIn), C.POINTER(XrtInject), _u64,
                                     C.POINTER(XrtOutputs), C.POINTER(XrtHistory), _vp]),
    'xrt_source_injected': (C.c_int, [_vp, C.POINTER(XrtSourceInject), _u64, C.POINTER(XrtHistory), _vp]),
    'xrt_source_generate': (C.c_int, [_vp, _u64, _u64, _u64, _u64, C.POINTER(XrtHistory), _vp]),
    'xrt_bundles_generate': (C.c_int, [C.POINTER(XrtPlasmaDesc), _u64, _u64, _u64, _vp, _vp, _vp, _vp]),
    'xrt_scene_set_bundles': (C.c_int, [_vp, _vp, _vp, _u64, _u64]),
    'xrt_bundle_voigt_tables': (C.c_int, [_vp, _vp, _u64, C.c_double, C.c_int32, _vp, _vp, _vp]),
    'xrt_scene_set_bundle_tables': (C.c_int, [_vp, _vp, _vp, C.c_int32]),
    'xrt_bits_to_ids': (C.c_int, [_vp, _u64, _u64, _vp, _u64, _vp, _vp]),
    'xrt_lost_select': (C.c_int, [_u64, _u64, _vp, _u64, _u64, _vp, _vp, _vp]),
    'xrt_fp64_burn': (C.c_int, [_u64, _vp, C.POINTER(C.c_double), _vp]),
    'xrt_launch_info': (C.c_int, [_vp, _pi32, _pi32, _pi32, _pi32]),
    'xrt_launch_info_cull': (C.c_int, [_vp, _pi32, _pi32, _pi32, _pi32]),
}

_lib = None


def load():
    """dlopen libxrt.so (built by ``python -m xicsrt_b200.build``); raises if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing. Build it with "python -m xicsrt_b200.build" (needs nvcc). '
                'xicsrt_b200 has no CPU fallback.')
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.xrt_version() != XRT_VERSION:
            raise ImportError(f'libxrt.so version {lib.xrt_version()} != binding version {XRT_VERSION}')
        _lib = lib
    return _lib


def check(code):
    if code != OK:
        raise XrtError(code, load().xrt_last_error().decode('utf-8', 'replace'))
