# -*- coding: utf-8 -*-
"""
oracle.scenes -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Named test scenes.  Geometry **G** is the spherical-crystal spectrometer of
the reference's ``testing/integrated_test_01.ipynb`` cell 2 (identical to
``examples/example_01/example_01.py:18-62``); the variants swap the optic
class, the source distributions or add apertures so that every branch of the
hot path is exercised.  ``make_golden.py`` runs the unmodified reference on
these; the tests run the oracle and the CUDA path on them.
"""
import copy

import numpy as np


def _general(seed=0, history=True):
    return {'number_of_iter': 1, 'number_of_runs': 1, 'random_seed': seed,
            'print_results': False, 'keep_history': history}


def source_G(n, **kw):
    s = {'class_name': 'XicsrtSourceDirected', 'intensity': n,
         'wavelength': 3.9492, 'spread': np.radians(10.0),
         'temperature': 1000.0, 'mass_number': 39.948, 'linewidth': 0.0,
         'xsize': 0.0, 'ysize': 0.0, 'zsize': 0.0}
    s.update(kw)
    return s


def crystal_G(class_name='XicsrtOpticSphericalCrystal', **kw):
    c = {'class_name': class_name, 'check_size': True,
         'origin': [0.0, 0.0, 0.80374151],
         'zaxis': [0.0, 0.59497864, -0.80374151],
         'xsize': 0.2, 'ysize': 0.2,
         'crystal_spacing': 2.45676, 'rocking_type': 'gaussian',
         'rocking_fwhm': 48.070e-6}
    c.update(kw)
    return c


def detector_G(**kw):
    d = {'class_name': 'XicsrtOpticDetector',
         'origin': [0.0, 0.76871290, 0.56904832],
         'zaxis': [0.0, -0.95641806, 0.29200084],
         'xsize': 0.4, 'ysize': 0.2}
    d.update(kw)
    return d


def assemble(source, optics, seed=0, filters=None, **general):
    cfg = {'general': _general(seed), 'sources': {'source': source}, 'optics': dict(optics)}
    cfg['general'].update(general)
    if filters:
        cfg['filters'] = filters
    return cfg


# ---------------------------------------------------------------------------

def sphere(n=10000, seed=0):
    """config 1 of BASELINE.json: directed source, Gaussian line, spherical crystal, detector."""
    return assemble(source_G(n), {'crystal': crystal_G(radius=1.0), 'detector': detector_G()}, seed)


def sphere_voigt(n=5000, seed=1):
    """config 1b: true Voigt line (natural linewidth > 0) -> tabulated CDF sampler."""
    return assemble(source_G(n, linewidth=1e14),
                    {'crystal': crystal_G(radius=1.0), 'detector': detector_G()}, seed)


def sphere_step_box(n=5000, seed=2):
    """step rocking curve; extended box source with a focused cone; Doppler shift; uniform line."""
    src = source_G(n, class_name='XicsrtSourceFocused', target=[0.0, 0.0, 0.80374151],
                   xsize=0.01, ysize=0.02, zsize=0.005, spread=np.radians(8.0),
                   wavelength_dist='uniform', wavelength_range=[3.9480, 3.9504],
                   velocity=[0.0, 1.0e5, 5.0e4])
    return assemble(src, {'crystal': crystal_G(radius=1.0, rocking_type='step', rocking_fwhm=400e-6,
                                               reflectivity=0.8),
                          'detector': detector_G()}, seed)


def plane_mirror(n=4000, seed=3):
    """planar mirror + gaussian-shaped source + flat cone."""
    src = source_G(n, class_name='XicsrtSourceGeneric', spatial_dist='gaussian',
                   xsize=0.004, ysize=0.002, zsize=0.001, angular_dist='flat',
                   spread=np.radians(6.0), wavelength_dist='monochrome')
    return assemble(src, {'mirror': _mirror('XicsrtOpticPlanarMirror'),
                          'detector': detector_G()}, seed)


def plane_crystal_xy(n=4000, seed=4):
    """planar crystal with rectangular isotropic cone (rejection sampler) and check_bragg off."""
    src = source_G(n, angular_dist='isotropic_xy', spread=[np.radians(4.0), np.radians(7.0)])
    return assemble(src, {'crystal': crystal_G('XicsrtOpticPlanarCrystal', check_bragg=False),
                          'detector': detector_G()}, seed)


def cylinder(n=5000, seed=5):
    src = source_G(n, angular_dist='flat_xy', spread=[-0.12, 0.10, -0.09, 0.11])
    return assemble(src, {'crystal': crystal_G('XicsrtOpticCylindricalCrystal', radius=1.0,
                                               rocking_fwhm=2000e-6),
                          'detector': detector_G()}, seed)


def cylinder_mirror_convex(n=4000, seed=6):
    return assemble(source_G(n), {'mirror': _mirror('XicsrtOpticCylindricalMirror', radius=2.0, convex=True),
                                  'detector': detector_G(xsize=2.0, ysize=2.0, pixel_size=0.02)}, seed)


def sphere_mirror_convex(n=4000, seed=7):
    return assemble(source_G(n), {'mirror': _mirror('XicsrtOpticSphericalMirror', radius=1.5, convex=True),
                                  'detector': detector_G(xsize=2.0, ysize=2.0, pixel_size=0.02)}, seed)


def _mirror(class_name, **kw):
    m = {'class_name': class_name, 'check_size': True,
         'origin': [0.0, 0.0, 0.80374151], 'zaxis': [0.0, 0.59497864, -0.80374151],
         'xsize': 0.2, 'ysize': 0.2}
    m.update(kw)
    return m


def torus(n=5000, seed=8, convex=(False, False), check_bragg=False):
    return assemble(source_G(n),
                    {'crystal': crystal_G('XicsrtOpticToroidalCrystal', radius_major=1.0, radius_minor=0.2,
                                          convex=list(convex), check_bragg=check_bragg,
                                          rocking_fwhm=2000e-6),
                     'detector': detector_G(xsize=1.0, ysize=1.0, pixel_size=0.01)}, seed)


def mosaic_sphere(n=4000, seed=9, cutoff=None, depth=15):
    """config 3: spherical HOPG-like mosaic crystal."""
    return assemble(source_G(n),
                    {'crystal': crystal_G('XicsrtOpticSphericalMosaicCrystal', radius=1.0,
                                          mosaic_spread=np.radians(0.4), mosaic_depth=depth,
                                          rocking_fwhm=200e-6, mosaic_cutoff=cutoff),
                     'detector': detector_G()}, seed)


def mosaic_plane(n=4000, seed=10):
    return assemble(source_G(n),
                    {'crystal': crystal_G('XicsrtOpticPlanarMosaicCrystal',
                                          mosaic_spread=np.radians(0.4), mosaic_depth=6,
                                          rocking_fwhm=200e-6, mosaic_cutoff=1e-8),
                     'detector': detector_G()}, seed)


def apertures(n=5000, seed=11):
    """example_02-style composite aperture in front of the crystal, all logic ops and shapes."""
    ap = [
        {'shape': 'circle', 'size': [0.09], 'logic': 'and'},
        {'shape': 'ellipse', 'size': [0.03, 0.015], 'origin': [0.02, 0.0], 'logic': 'not'},
        {'shape': 'rectangle', 'size': [0.02, 0.05], 'origin': [-0.05, 0.01], 'logic': 'xor'},
        {'shape': 'square', 'size': [0.01], 'origin': [0.0, -0.06], 'logic': 'or'},
        {'shape': 'triangle', 'vertices': [[0.0, 0.0], [0.03, 0.0], [0.0, 0.04]],
         'origin': [-0.01, 0.03], 'logic': 'xnor'},
        {'shape': 'circle', 'size': [0.004], 'origin': [0.07, 0.07], 'logic': 'nor'},
        {'shape': 'circle', 'size': [0.2], 'logic': 'nand'},
        {'shape': 'circle', 'size': [0.095], 'logic': 'nand'},
    ]
    aperture = {'class_name': 'XicsrtOpticAperture', 'origin': [0.0, 0.0, 0.5],
                'zaxis': [0.0, 0.0, -1.0], 'aperture': ap}
    return assemble(source_G(n),
                    {'aperture': aperture,
                     'crystal': crystal_G(radius=1.0, check_bragg=False,
                                          aperture={'shape': 'circle', 'size': [0.08]}),
                     'detector': detector_G()}, seed)


def local_frames(n=4000, seed=12):
    """trace_local on analytic optics + explicit xaxis + zsize check."""
    c = crystal_G('XicsrtOpticPlanarCrystal', check_bragg=False, trace_local=True,
                  xaxis=[1.0, 0.0, 0.0], zsize=0.01)
    d = detector_G(trace_local=True)
    return assemble(source_G(n), {'crystal': c, 'detector': d}, seed)


def sphere_rocking_file(n=20000, seed=19):
    """rocking_type='file' (_InteractCrystal.py:151-178): XOP diff_pat table, sigma / pi mix, reflectivity < 1."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'profiles',
                        'diff_pat.dat')
    return assemble(source_G(n), {'crystal': crystal_G(radius=1.0, rocking_type='file', rocking_file=path,
                                                       rocking_mix=0.3, reflectivity=0.9),
                                  'detector': detector_G()}, seed)


def two_iter_two_runs(n=3000, seed=13):
    cfg = sphere(n, seed)
    cfg['general']['number_of_iter'] = 2
    cfg['general']['number_of_runs'] = 2
    cfg['general']['history_max_lost'] = 400
    return cfg


ANALYTIC = {
    'sphere': sphere, 'sphere_voigt': sphere_voigt, 'sphere_step_box': sphere_step_box,
    'plane_mirror': plane_mirror, 'plane_crystal_xy': plane_crystal_xy,
    'cylinder': cylinder, 'cylinder_mirror_convex': cylinder_mirror_convex,
    'sphere_mirror_convex': sphere_mirror_convex,
    'torus_ff': lambda: torus(convex=(False, False)),
    'torus_ft': lambda: torus(seed=14, convex=(False, True)),
    'torus_tf': lambda: torus(seed=15, convex=(True, False)),
    'torus_tt': lambda: torus(seed=16, convex=(True, True)),
    'torus_bragg': lambda: torus(n=20000, seed=17, check_bragg=True),
    'mosaic_sphere': mosaic_sphere,
    'mosaic_sphere_cutoff': lambda: mosaic_sphere(seed=18, cutoff=1e-8, depth=5),
    'mosaic_plane': mosaic_plane,
    'apertures': apertures, 'local_frames': local_frames,
    'sphere_rocking_file': sphere_rocking_file,
    'two_iter_two_runs': two_iter_two_runs,
}


def get(name):
    from oracle import scenes_ext
    table = dict(ANALYTIC)
    table.update(scenes_ext.EXTRA)
    return copy.deepcopy(table[name]())


def names():
    from oracle import scenes_ext
    return list(ANALYTIC) + list(scenes_ext.EXTRA)
