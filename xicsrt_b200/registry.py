# -*- coding: utf-8 -*-
"""
The ``class_name`` registry: which names the path accepts, what defaults each
carries and which (interaction, shape) pair or source kind it denotes.

The reference discovers classes by globbing ``_Xicsrt*.py`` files
(``xicsrt/objects/_Dispatcher.py:63-111``) and every optic is an empty
``class X(Interact*, Shape*)`` mixin (``xicsrt/optics/_XicsrtOptic*.py``).
Here the same surface is a table; default-config key order follows the
reference's MRO (ConfigObject -> GeometryObject -> TraceObject -> Shape* ->
Interact*), so ``output['config']`` prints the same way.
"""
import numpy as np

# ---------------------------------------------------------------------------
# default-config fragments (file:line of the reference default_config methods)

def _d_config_object(class_name):
    # xicsrt/objects/_ConfigObject.py:42-55
    return {'class_name': class_name,
            'yo_mama': 'Is a beautiful person and she loves you.'}


def _d_geometry():
    # xicsrt/objects/_GeometryObject.py:37-62
    return {'origin': np.array([0.0, 0.0, 0.0]),
            'zaxis': np.array([0.0, 0.0, 1.0]),
            'xaxis': None}


def _d_trace():
    # xicsrt/optics/_TraceObject.py:80-100
    return {'xsize': None, 'ysize': None, 'zsize': None, 'pixel_size': None,
            'trace_local': False, 'check_size': True, 'check_aperture': True,
            'aperture': None, 'filters': []}


def _d_shape(shape):
    if shape == 'plane':
        return {}
    if shape in ('sphere', 'cylinder'):
        # _ShapeSphere.py:32-35, _ShapeCylinder.py:31-34
        return {'radius': 1.0, 'convex': False}
    if shape == 'torus':
        # _ShapeTorus.py:47-52
        return {'radius_major': 1.0, 'radius_minor': 0.2, 'convex': [False, False]}
    mesh = {'mesh_points': None, 'mesh_normals': None, 'mesh_faces': None,
            'mesh_coarse_points': None, 'mesh_coarse_normals': None,
            'mesh_coarse_faces': None, 'mesh_interpolate': None,
            'mesh_refine': None,  # _ShapeMesh.py:96-109
            # not in the reference: find the fine face by the full Moeller-Trumbore test (accelerated by a face grid)
            # instead of the lossy coarse -> nearest-vertex pre-selection the reference documents in
            # _ShapeMesh.py:52-79 and TODO:3-8; results are those of mesh_refine=False
            'mesh_lossless': False}
    if shape == 'mesh':
        return mesh
    if shape == 'mesh_sphere':
        # _ShapeMeshSphere.py:50-58
        mesh.update({'radius': 1.0, 'mesh_size': (11, 11),
                     'mesh_coarse_size': (5, 5), 'trace_local': True})
        return mesh
    if shape == 'mesh_cylinder':
        # _ShapeMeshCylinder.py:48-61
        mesh.update({'mesh_refine': True, 'mesh_size': (11, 11),
                     'mesh_coarse_size': (5, 5), 'mesh_xsize': None,
                     'mesh_ysize': None, 'radius': 1.0, 'trace_local': True})
        return mesh
    if shape == 'mesh_torus':
        # _ShapeMeshTorus.py:66-82
        mesh.update({'mesh_refine': True, 'mesh_size': (11, 11),
                     'mesh_coarse_size': (5, 5), 'mesh_xsize': None,
                     'mesh_ysize': None, 'radius_major': 1.0,
                     'radius_minor': 0.2, 'convex': [False, False],
                     'normal_method': 'analytic', 'trace_local': True})
        return mesh
    raise KeyError(shape)


def _d_interact(kind):
    if kind in ('none', 'mirror'):
        return {}
    crystal = {'crystal_spacing': 0.0, 'reflectivity': 1.0, 'check_bragg': True,
               'rocking_type': 'gaussian', 'rocking_fwhm': None,
               'rocking_file': None, 'rocking_filetype': None,
               'rocking_mix': 0.5}  # _InteractCrystal.py:71-84
    if kind == 'crystal':
        return crystal
    if kind == 'mosaic':
        # _InteractMosaicCrystal.py:47-51
        crystal.update({'mosaic_spread': 0.0, 'mosaic_depth': 15,
                        'mosaic_cutoff': None})
        return crystal
    raise KeyError(kind)


def _d_source():
    # xicsrt/sources/_XicsrtSourceGeneric.py:157-186
    return {'xsize': 0.0, 'ysize': 0.0, 'zsize': 0.0,
            'intensity': 0.0, 'use_poisson': False,
            'spatial_dist': 'uniform', 'angular_dist': 'isotropic',
            'spread': np.pi, 'wavelength_dist': 'voigt', 'wavelength': 1.0,
            'mass_number': 1.0, 'linewidth': 0.0, 'temperature': 0.0,
            'velocity': np.array([0.0, 0.0, 0.0]),
            'wavelength_range': np.array([0.0, 0.0]), 'filters': []}


def _d_plasma():
    # xicsrt/sources/_XicsrtPlasmaGeneric.py:128-158
    return {'xsize': 0.0, 'ysize': 0.0, 'zsize': 0.0,
            'angular_dist': 'isotropic', 'spread': None, 'spread_radius': None,
            'target': None, 'use_poisson': False, 'wavelength_dist': 'voigt',
            'wavelength': 1.0, 'wavelength_range': None, 'mass_number': 1.0,
            'linewidth': 0.0, 'emissivity': 0.0, 'temperature': 0.0,
            'velocity': 0.0, 'time_resolution': 1e-3, 'bundle_type': 'voxel',
            'bundle_volume': 1e-6, 'bundle_count': None,
            'max_rays': int(1e7), 'max_bundles': int(1e7), 'filters': []}


# ---------------------------------------------------------------------------
# the registry proper

# class_name -> (interaction kind, shape kind); reference file _<name>.py:15-17
OPTICS = {
    'XicsrtOpticAperture': ('none', 'plane'),
    'XicsrtOpticDetector': ('none', 'plane'),
    'XicsrtOpticPlanarMirror': ('mirror', 'plane'),
    'XicsrtOpticSphericalMirror': ('mirror', 'sphere'),
    'XicsrtOpticCylindricalMirror': ('mirror', 'cylinder'),
    'XicsrtOpticMeshMirror': ('mirror', 'mesh'),
    'XicsrtOpticPlanarCrystal': ('crystal', 'plane'),
    'XicsrtOpticSphericalCrystal': ('crystal', 'sphere'),
    'XicsrtOpticCylindricalCrystal': ('crystal', 'cylinder'),
    'XicsrtOpticToroidalCrystal': ('crystal', 'torus'),
    'XicsrtOpticMeshCrystal': ('crystal', 'mesh'),
    'XicsrtOpticMeshSphericalCrystal': ('crystal', 'mesh_sphere'),
    'XicsrtOpticMeshCylindricalCrystal': ('crystal', 'mesh_cylinder'),
    'XicsrtOpticMeshToroidalCrystal': ('crystal', 'mesh_torus'),
    'XicsrtOpticPlanarMosaicCrystal': ('mosaic', 'plane'),
    'XicsrtOpticSphericalMosaicCrystal': ('mosaic', 'sphere'),
    'XicsrtOpticMeshMosaicCrystal': ('mosaic', 'mesh'),
}

# class_name -> source kind
SOURCES = {
    'XicsrtSourceGeneric': 'generic',
    'XicsrtSourceDirected': 'directed',
    'XicsrtSourceFocused': 'focused',
    'XicsrtPlasmaGeneric': 'plasma_generic',
    'XicsrtPlasmaCubic': 'plasma_cubic',
    'XicsrtPlasmaToroidal': 'plasma_toroidal',
    'XicsrtPlasmaToroidalDatafile': 'plasma_datafile',
}

FILTERS = {
    'XicsrtBundleFilter': 'none',
    'XicsrtBundleFilterSightline': 'sightline',
}


def find(class_name):
    """Return (section, kind) for a class name or raise like the reference."""
    if class_name in OPTICS:
        return 'optics', OPTICS[class_name]
    if class_name in SOURCES:
        return 'sources', SOURCES[class_name]
    if class_name in FILTERS:
        return 'filters', FILTERS[class_name]
    if class_name == 'XicsrtPlasmaCylindrical':
        # reference _XicsrtPlasmaCylindrical.py:22-23 declares itself broken
        raise NotImplementedError('XicsrtPlasmaCylindrical is broken in the reference and not provided.')
    raise Exception('Could not find {} in available objects.'.format(class_name))


def defaults(class_name):
    """The default config dict of one class, keys in the reference's MRO order."""
    section, kind = find(class_name)
    cfg = _d_config_object(class_name)
    cfg.update(_d_geometry())
    if section == 'optics':
        interact, shape = kind
        cfg.update(_d_trace())
        cfg.update(_d_shape(shape))
        cfg.update(_d_interact(interact))
    elif section == 'sources':
        if kind.startswith('plasma'):
            cfg.update(_d_plasma())
            if kind in ('plasma_toroidal', 'plasma_datafile'):
                # _XicsrtPlasmaToroidal.py:24-32
                cfg.update({'major_radius': 0.0, 'minor_radius': 0.0,
                            'torus_origin': np.array([0.0, 0.0, 0.0]),
                            'emissivity_scale': 1.0, 'temperature_scale': 1.0,
                            'velocity_scale': 1.0})
            if kind == 'plasma_datafile':
                # _XicsrtPlasmaToroidalDatafile.py:23-28
                cfg.update({'emissivity_file': None, 'temperature_file': None,
                            'velocity_file': None})
        else:
            cfg.update(_d_source())
            if kind == 'directed':
                cfg['direction'] = None   # _XicsrtSourceDirected.py:37
            elif kind == 'focused':
                cfg['target'] = None      # _XicsrtSourceFocused.py:32
    elif section == 'filters':
        if kind == 'sightline':
            cfg['radius'] = None          # _XicsrtBundleFilterSightline.py:29
    return cfg
