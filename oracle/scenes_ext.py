# -*- coding: utf-8 -*-
"""
oracle.scenes_ext -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Scenes for the plasma (extended) sources and the mesh optics.
"""
import os

import numpy as np

from oracle.scenes import assemble, crystal_G, detector_G

PROFILES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'profiles')


def _plasma(class_name, **kw):
    p = {'class_name': class_name, 'origin': [0.0, 0.0, 0.0], 'zaxis': [0.0, 0.0, 1.0],
         'xsize': 0.02, 'ysize': 0.03, 'zsize': 0.01, 'target': [0.0, 0.0, 0.80374151],
         'spread': np.radians(6.0), 'bundle_type': 'voxel', 'bundle_volume': 1e-9, 'bundle_count': 40,
         'time_resolution': 1.0, 'wavelength': 3.9492, 'mass_number': 39.948, 'linewidth': 0.0,
         'temperature': 1000.0, 'emissivity': 1e10, 'max_rays': int(1e7)}
    p.update(kw)
    return p


def plasma_cubic(seed=21):
    """XicsrtPlasmaCubic, fixed ray counts per bundle (no Poisson), config-5 style."""
    return assemble(_plasma('XicsrtPlasmaCubic'),
                    {'crystal': crystal_G(radius=1.0, rocking_fwhm=2000e-6), 'detector': detector_G()}, seed)


def plasma_cubic_poisson(seed=22):
    return assemble(_plasma('XicsrtPlasmaCubic', use_poisson=True, emissivity=3e9, spread_radius=0.08, spread=None),
                    {'crystal': crystal_G(radius=1.0, rocking_fwhm=2000e-6), 'detector': detector_G()}, seed)


def plasma_toroidal(seed=23):
    """Toroidal plasma with a flow velocity and a sightline bundle filter."""
    src = _plasma('XicsrtPlasmaToroidal', major_radius=0.5, minor_radius=0.05,
                  torus_origin=[-0.5, 0.0, 0.0], velocity=np.array([0.0, 2.0e4, 1.0e4]),
                  emissivity=2e10, bundle_count=60, filters=['sight'])
    filters = {'sight': {'class_name': 'XicsrtBundleFilterSightline', 'origin': np.array([0.0, 0.0, 0.0]),
                         'zaxis': np.array([0.0, 0.0, 1.0]), 'radius': 0.012}}
    return assemble(src, {'crystal': crystal_G(radius=1.0, rocking_fwhm=2000e-6), 'detector': detector_G()},
                    seed, filters=filters)


def plasma_datafile(seed=24):
    src = _plasma('XicsrtPlasmaToroidalDatafile', major_radius=0.5, minor_radius=0.05,
                  torus_origin=[-0.5, 0.0, 0.0], bundle_count=50, time_resolution=1e-5,
                  temperature_file=os.path.join(PROFILES, 'temperature.txt'),
                  emissivity_file=os.path.join(PROFILES, 'emissivity.txt'), use_poisson=True)
    return assemble(src, {'crystal': crystal_G(radius=1.0, rocking_fwhm=2000e-6), 'detector': detector_G()}, seed)


def plasma_voigt(seed=25):
    """Natural linewidth > 0 on a plasma whose temperature varies from bundle to bundle: every
    per-bundle source of the reference builds its own Voigt table."""
    src = _plasma('XicsrtPlasmaToroidalDatafile', major_radius=0.5, minor_radius=0.05,
                  torus_origin=[-0.5, 0.0, 0.0], bundle_count=30, time_resolution=4e-5,
                  temperature_file=os.path.join(PROFILES, 'temperature.txt'),
                  emissivity_file=os.path.join(PROFILES, 'emissivity.txt'), linewidth=1e14)
    return assemble(src, {'crystal': crystal_G(radius=1.0, rocking_fwhm=2000e-6), 'detector': detector_G()}, seed)


def plasma_cone(angular_dist, seed):
    """Per-bundle sources with a non-isotropic emission cone (scalar spread per bundle)."""
    src = _plasma('XicsrtPlasmaCubic', angular_dist=angular_dist, spread_radius=0.09, spread=None, bundle_count=30,
                  emissivity=2e11)
    return assemble(src, {'crystal': crystal_G(radius=1.0, rocking_fwhm=2000e-6), 'detector': detector_G()}, seed)


def _mesh_optic(class_name, **kw):
    c = crystal_G(class_name, check_bragg=False, rocking_fwhm=2000e-6)
    c.update(kw)
    return c


def mesh_torus(n=4000, seed=31, mesh_size=(21, 21), **kw):
    """config 4: XicsrtOpticMeshToroidalCrystal, coarse -> fine refinement + Clough-Tocher interpolation."""
    from oracle.scenes import source_G
    c = _mesh_optic('XicsrtOpticMeshToroidalCrystal', radius_major=1.0, radius_minor=0.2,
                    mesh_size=mesh_size, mesh_coarse_size=(5, 5), **kw)
    return assemble(source_G(n), {'crystal': c, 'detector': detector_G(xsize=1.0, ysize=1.0, pixel_size=0.01)}, seed)


def mesh_sphere(n=4000, seed=32):
    from oracle.scenes import source_G
    c = _mesh_optic('XicsrtOpticMeshSphericalCrystal', radius=1.0, mesh_size=(15, 15), mesh_coarse_size=(4, 4),
                    check_bragg=True)
    return assemble(source_G(n), {'crystal': c, 'detector': detector_G()}, seed)


def mesh_cylinder(n=4000, seed=33):
    from oracle.scenes import source_G
    c = _mesh_optic('XicsrtOpticMeshCylindricalCrystal', radius=1.0, mesh_size=(13, 17), mesh_coarse_size=(5, 5))
    return assemble(source_G(n), {'crystal': c, 'detector': detector_G(xsize=1.0, ysize=1.0, pixel_size=0.01)}, seed)


def _user_mesh(nx=9, ny=7):
    """A hand-made saddle surface in local coordinates with analytic normals."""
    x = np.linspace(-0.1, 0.1, nx)
    y = np.linspace(-0.1, 0.1, ny)
    xx, yy = np.meshgrid(x, y, indexing='ij')
    zz = 0.4 * xx**2 - 0.3 * yy**2 + 0.05 * xx * yy
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    nrm = np.stack([-(0.8 * xx + 0.05 * yy).ravel(), -(-0.6 * yy + 0.05 * xx).ravel(), np.ones(xx.size)], axis=1)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    return pts, nrm


def mesh_user_flat(n=4000, seed=34):
    """XicsrtOpticMeshMirror with user points only: all faces tested, flat face normals, no interpolation."""
    from oracle.scenes import source_G
    pts, _ = _user_mesh()
    m = {'class_name': 'XicsrtOpticMeshMirror', 'origin': [0.0, 0.0, 0.80374151],
         'zaxis': [0.0, 0.59497864, -0.80374151], 'xsize': 0.2, 'ysize': 0.2, 'trace_local': True,
         'mesh_points': pts}
    return assemble(source_G(n), {'mirror': m, 'detector': detector_G(xsize=1.0, ysize=1.0, pixel_size=0.01)}, seed)


def mesh_user_interp(n=4000, seed=35):
    """XicsrtOpticMeshCrystal with user points + normals (interpolation on, no coarse mesh)."""
    from oracle.scenes import source_G
    pts, nrm = _user_mesh(11, 11)
    c = _mesh_optic('XicsrtOpticMeshCrystal', trace_local=True, mesh_points=pts, mesh_normals=nrm)
    return assemble(source_G(n), {'crystal': c, 'detector': detector_G(xsize=1.0, ysize=1.0, pixel_size=0.01)}, seed)


def mesh_mosaic(n=3000, seed=36):
    """XicsrtOpticMeshMosaicCrystal on user points + normals."""
    from oracle.scenes import source_G
    pts, nrm = _user_mesh(11, 11)
    c = _mesh_optic('XicsrtOpticMeshMosaicCrystal', trace_local=True, mesh_points=pts, mesh_normals=nrm,
                    check_bragg=True, mosaic_spread=np.radians(0.4), mosaic_depth=4, rocking_fwhm=2000e-6)
    return assemble(source_G(n), {'crystal': c, 'detector': detector_G(xsize=1.0, ysize=1.0, pixel_size=0.01)}, seed)


EXTRA = {
    'plasma_cubic': plasma_cubic,
    'plasma_cubic_poisson': plasma_cubic_poisson,
    'plasma_toroidal': plasma_toroidal,
    'plasma_datafile': plasma_datafile,
    'plasma_voigt': plasma_voigt,
    'plasma_flat': lambda: plasma_cone('flat', 26),
    'plasma_flat_xy': lambda: plasma_cone('flat_xy', 27),
    'plasma_isotropic_xy': lambda: plasma_cone('isotropic_xy', 28),
    'mesh_torus': mesh_torus,
    # config 4 of BASELINE.json at its stated size: 41 x 41 fine (1681 points / 3200 faces), 5 x 5 coarse
    'mesh_torus_41': lambda: mesh_torus(n=6000, seed=38, mesh_size=(41, 41)),
    # the full Moeller-Trumbore test on 3200 fine faces (mesh_refine off): the face-grid path of the kernels
    'mesh_torus_norefine': lambda: mesh_torus(n=5000, seed=39, mesh_size=(41, 41), mesh_refine=False),
    # product option mesh_lossless: a refining mesh traced with the full test; same rays as the scene above
    'mesh_torus_lossless': lambda: mesh_torus(n=5000, seed=39, mesh_size=(41, 41), mesh_lossless=True),
    # refinement without interpolation: flat normals of the fine face that was hit (_ShapeMesh.py:428-432)
    'mesh_torus_flat_normals': lambda: mesh_torus(n=5000, seed=40, mesh_size=(17, 23), mesh_interpolate=False),
    'mesh_torus_convex': lambda: mesh_torus(seed=37, convex=[True, False], mesh_size=(15, 15)),
    'mesh_sphere': mesh_sphere,
    'mesh_cylinder': mesh_cylinder,
    'mesh_user_flat': mesh_user_flat,
    'mesh_user_interp': mesh_user_interp,
    'mesh_mosaic': mesh_mosaic,
}
