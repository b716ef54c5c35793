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


EXTRA = {
    'plasma_cubic': plasma_cubic,
    'plasma_cubic_poisson': plasma_cubic_poisson,
    'plasma_toroidal': plasma_toroidal,
    'plasma_datafile': plasma_datafile,
}
