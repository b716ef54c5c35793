# -*- coding: utf-8 -*-
"""
Parity of the BENCHMARKED kernel at the scale it is benchmarked (``-m gpu``).

The bit-exact oracle tests (test_gpu_parity.py) drive the straight replay kernel through
``xrt_trace_injected``; the fused kernel that bench.py times adds work-skipping stages in front
of the same ray code: an FP32 broad phase, a two-level conservative Bragg pre-test, table
``sincos``, lazy / deferred wavelengths and shared-memory queues.  Each of them is allowed to
*skip* work only, never to change a result.  This file proves that at 1e9 rays per launch:

  1. for every geometry below and three seeds, the history-off launch (exactly what bench.py
     times) gives the same counters and images, and the history-on launch the same found-id
     set, with the broad phase off (XRT_NO_BROAD32) and with every pre-test off (XRT_NO_CULL);
     the geometries include the enable thresholds of the broad phase (|C - O|^2 ~ 4 R^2,
     sin(theta_B) ~ 0.1), the planar limit R = 1e5 of the reference's
     testing/integrated_test_02.ipynb, reflectivity < 1, a step curve, box / focused / moving
     sources and a plasma (the generic fast path);
  2. the found rays of a 1e8-ray run have the distributions of the oracle's own 1e8-ray run
     (``oracle.raytrace_mp``): binomial z on the per-element counts, two-sample chi-square on the
     detector image, two-sample KS on the found rays' local x, y and wavelength
     (reference side >= 1e8 rays).

A conservative bound that is violated with probability 1e-7 per ray shows up ~100 times in 1e9
rays; the 3e6-ray versions of these tests in test_gpu_statistics.py could not see it.
"""
import copy
import os

import numpy as np
import pytest

import oracle
from oracle import scenes

pytestmark = pytest.mark.gpu

N_SCALE = int(os.environ.get('XRT_TEST_SCALE_RAYS', 1_000_000_000))
SEEDS = (11, 2025, 90210)
ENVS = ({}, {'XRT_NO_BROAD32': '1'}, {'XRT_NO_CULL': '1'})
LAMBDA = 3.9492


@pytest.fixture(scope='module')
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def rowland(sin_b=0.80374151, radius=1.0, source_distance=None, detector_distance=None, spread_deg=10.0, size=0.2,
            n=N_SCALE, **crystal_kw):
    """
    Point source at the origin looking along +z at a concave spherical crystal whose vertex is met at the Bragg
    angle asin(sin_b); detector square on the reflected central ray at the same distance.  With the defaults this
    is geometry G of SURVEY.md section 8d (source on the Rowland circle: distance = R sin(theta_B)).
    """
    cos_b = float(np.sqrt(1.0 - sin_b * sin_b))
    s = radius * sin_b if source_distance is None else source_distance
    refl = np.array([0.0, 2.0 * sin_b * cos_b, 1.0 - 2.0 * sin_b * sin_b])
    s_det = radius * sin_b if detector_distance is None else detector_distance
    crystal = {'class_name': 'XicsrtOpticSphericalCrystal', 'check_size': True, 'origin': [0.0, 0.0, s],
               'zaxis': [0.0, cos_b, -sin_b], 'xsize': size, 'ysize': size, 'radius': radius,
               'crystal_spacing': LAMBDA / (2.0 * sin_b), 'rocking_type': 'gaussian', 'rocking_fwhm': 48.070e-6}
    crystal.update(crystal_kw)
    detector = {'class_name': 'XicsrtOpticDetector', 'origin': (np.array([0.0, 0.0, s]) + s_det * refl).tolist(),
                'zaxis': (-refl).tolist(), 'xsize': 2.0 * size + 0.4 * s_det, 'ysize': 2.0 * size + 0.4 * s_det,
                'pixel_size': (2.0 * size + 0.4 * s_det) / 100.0}
    source = scenes.source_G(n, spread=float(np.radians(spread_deg)))
    return scenes.assemble(source, {'crystal': crystal, 'detector': detector}, 0, keep_history=False)


def _geometries():
    import bench
    g = {}
    g['config2'] = bench.spectrometer(N_SCALE)
    g['config2_reflectivity_0.37'] = bench.spectrometer(N_SCALE)
    g['config2_reflectivity_0.37']['optics']['crystal']['reflectivity'] = 0.37
    g['config2_step_curve'] = bench.spectrometer(N_SCALE)
    g['config2_step_curve']['optics']['crystal'].update({'rocking_type': 'step', 'reflectivity': 0.9})
    # |C - O|^2 = s^2 + R^2 - 2 s R sin(theta_B); the broad phase is enabled up to 4 R^2: s = 2.7120 -> 3.9950 R^2
    # (on), s = 2.7160 -> 4.0103 R^2 (off: FP64 stage A for every ray)
    g['broad_phase_threshold_inside'] = rowland(source_distance=2.7120, spread_deg=4.0)
    g['broad_phase_threshold_outside'] = rowland(source_distance=2.7160, spread_deg=4.0)
    # grazing incidence: sin(theta_B) just above / below the 0.1 enable threshold of the broad phase
    g['sin_bragg_0.1002'] = rowland(sin_b=0.1002, spread_deg=12.0)
    g['sin_bragg_0.0998'] = rowland(sin_b=0.0998, spread_deg=12.0)
    # planar limit of testing/integrated_test_02.ipynb: radius 1e5 (point source, and its 0.1 m box source)
    g['planar_limit_r1e5'] = rowland(radius=1e5, source_distance=0.80374151, detector_distance=0.80374151, spread_deg=5.0)
    c = rowland(radius=1e5, source_distance=0.80374151, detector_distance=0.80374151, spread_deg=5.0)
    c['sources']['source'].update({'xsize': 0.10, 'ysize': 0.10})
    g['planar_limit_r1e5_box_source'] = c
    # wide cone: most rays miss the crystal, many miss the sphere (NaN paths of the broad phase)
    g['wide_cone_75deg'] = bench.spectrometer(N_SCALE)
    g['wide_cone_75deg']['sources']['source']['spread'] = float(np.radians(75.0))
    # broad line, narrow lossy curve
    c = bench.spectrometer(N_SCALE)
    c['sources']['source']['temperature'] = 40000.0
    c['optics']['crystal'].update({'rocking_fwhm': 9e-6, 'reflectivity': 0.5})
    g['broad_line_narrow_curve'] = c
    # the generic fast path: box source, focused source, moving source
    c = bench.spectrometer(N_SCALE)
    c['sources']['source'].update({'xsize': 1e-3, 'ysize': 1e-3, 'zsize': 1e-3})
    g['box_source_1mm'] = c
    c = bench.spectrometer(N_SCALE)
    c['sources']['source'].update({'class_name': 'XicsrtSourceFocused', 'target': [0.0, 0.0, 0.80374151],
                                   'xsize': 0.02, 'ysize': 0.02, 'zsize': 0.02, 'spread': float(np.radians(8.0))})
    g['focused_box_source_2cm'] = c
    c = bench.spectrometer(N_SCALE)
    c['sources']['source']['velocity'] = [0.0, 3.0e4, 1.0e5]
    g['doppler_shifted_line'] = c
    g['config5_plasma'] = bench.workload_config('config5', N_SCALE)
    # mosaic crystals (stage S: per-layer FP32 pre-test of the crystallite loop); 1e8 rays as BASELINE.json's config 3
    n3 = max(N_SCALE // 10, 1000)
    g['config3_mosaic'] = bench.workload_config('config3', n3)
    c = bench.workload_config('config3', n3)
    c['optics']['crystal'].update({'mosaic_cutoff': 1e-8, 'reflectivity': 0.6, 'mosaic_depth': 7})
    g['config3_mosaic_cutoff_lossy'] = c
    c = bench.workload_config('config3', n3)
    c['optics']['crystal'].update({'rocking_type': 'step', 'rocking_fwhm': 400e-6, 'mosaic_spread': float(np.radians(0.1))})
    g['config3_mosaic_step_curve'] = c
    c = bench.workload_config('config3', n3)
    c['optics']['crystal'].pop('radius')
    c['optics']['crystal']['class_name'] = 'XicsrtOpticPlanarMosaicCrystal'
    g['config3_mosaic_planar'] = c
    # the other source kinds of the mosaic broad phase (k_mosaic32<box>, <focused>), a lossy shallow crystal, and a run of
    # more than 2^27 rays (three launches of the broad phase: 27-bit id offsets beside the layer tag)
    c = bench.workload_config('config3', n3)
    c['sources']['source'].update({'xsize': 2e-3, 'ysize': 1e-3, 'zsize': 3e-3})
    g['config3_mosaic_box_source'] = c
    c = bench.workload_config('config3', n3)
    c['sources']['source'].update({'class_name': 'XicsrtSourceFocused', 'target': [0.0, 0.0, 0.80374151],
                                   'xsize': 0.02, 'ysize': 0.02, 'zsize': 0.02, 'spread': float(np.radians(8.0))})
    g['config3_mosaic_focused_source'] = c
    c = bench.workload_config('config3', n3)
    c['optics']['crystal'].update({'reflectivity': 0.55, 'mosaic_depth': 4, 'mosaic_spread': float(np.radians(0.8))})
    g['config3_mosaic_lossy_depth4'] = c
    g['config3_mosaic_3e8'] = bench.workload_config('config3', 3 * n3)
    # the mosaic broad phase at the enable thresholds it shares with k_cull32 and in the planar limit
    mosaic = {'class_name': 'XicsrtOpticSphericalMosaicCrystal', 'mosaic_spread': float(np.radians(0.4)), 'mosaic_depth': 15,
              'rocking_fwhm': 200e-6}
    g['mosaic_threshold_inside'] = rowland(source_distance=2.7120, spread_deg=4.0, n=n3, **mosaic)
    g['mosaic_sin_bragg_0.1002'] = rowland(sin_b=0.1002, spread_deg=12.0, n=n3, **mosaic)
    g['mosaic_planar_limit_r1e5'] = rowland(radius=1e5, source_distance=0.80374151, detector_distance=0.80374151,
                                            spread_deg=5.0, n=n3, **mosaic)
    return g


GEOMETRIES = ['config2', 'config2_reflectivity_0.37', 'config2_step_curve', 'broad_phase_threshold_inside',
              'broad_phase_threshold_outside', 'sin_bragg_0.1002', 'sin_bragg_0.0998', 'planar_limit_r1e5',
              'planar_limit_r1e5_box_source', 'wide_cone_75deg', 'broad_line_narrow_curve', 'box_source_1mm',
              'focused_box_source_2cm', 'doppler_shifted_line', 'config5_plasma', 'config3_mosaic',
              'config3_mosaic_cutoff_lossy', 'config3_mosaic_step_curve', 'config3_mosaic_planar',
              'config3_mosaic_box_source', 'config3_mosaic_focused_source', 'config3_mosaic_lossy_depth4',
              'config3_mosaic_3e8', 'mosaic_threshold_inside', 'mosaic_sin_bragg_0.1002', 'mosaic_planar_limit_r1e5']


MOSAIC32_SCENES = {'config3_mosaic', 'config3_mosaic_step_curve', 'config3_mosaic_box_source', 'config3_mosaic_focused_source',
                   'config3_mosaic_lossy_depth4', 'config3_mosaic_3e8', 'mosaic_threshold_inside', 'mosaic_sin_bragg_0.1002',
                   'mosaic_planar_limit_r1e5'}


@pytest.mark.timeout(900)
@pytest.mark.parametrize('name', GEOMETRIES)
def test_work_skipping_stages_change_no_result_at_bench_scale(torch, name, monkeypatch):
    from xicsrt_b200 import _driver, config as xconfig
    cfg = _geometries()[name]
    cfg['general']['keep_history'] = False
    results = []
    for env in ENVS:
        for key in ('XRT_NO_BROAD32', 'XRT_NO_CULL'):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        per_seed = []
        for seed in SEEDS:
            tracer = _driver.Tracer(xconfig.get_config(xconfig.to_numpy(copy.deepcopy(cfg))), seed=seed)
            if not env and name in MOSAIC32_SCENES:          # the plan under test is the one that runs
                assert tracer.scene.launch_info()['mosaic_broad_phase'] is not None, name
            if env and name in MOSAIC32_SCENES:
                assert tracer.scene.launch_info()['mosaic_broad_phase'] is None, name
            tracer.trace(1)                                  # history off: the launch bench.py times
            packed_off = tracer.packed.clone()
            found, lost = tracer.select_ids(1, 64)           # history on: found / lost lists
            assert torch.equal(tracer.packed, packed_off), f'{name}: history-on launch counts differ from history-off'
            per_seed.append((tracer.n_rays, packed_off, found.clone()))
            tracer.close()
        results.append(per_seed)
    for s, seed in enumerate(SEEDS):
        n0, packed0, found0 = results[0][s]
        assert n0 >= 0.09 * N_SCALE
        n_src, n_det = int(packed0[0]), int(packed0[2])
        assert n_src == n0
        assert int(found0.numel()) == n_det
        if N_SCALE >= 10**8:
            assert n_det > 1000, f'{name}: only {n_det} rays detected -- the geometry does not exercise the Bragg path'
        for e, env in enumerate(ENVS[1:], start=1):
            n1, packed1, found1 = results[e][s]
            assert n1 == n0
            assert torch.equal(packed1, packed0), f'{name} seed {seed}: counters / images differ with {env}'
            assert torch.equal(found1, found0), f'{name} seed {seed}: found-id set differs with {env}'


N_MESH = max(N_SCALE // 10, 1000)


def _mesh_geometries():
    import bench
    g = {}
    g['config4'] = bench.workload_config('config4', N_MESH)                    # point source: dot-product pre-selection
    c = bench.workload_config('config4', N_MESH)
    c['sources']['source'].update({'xsize': 5e-3, 'ysize': 5e-3, 'zsize': 5e-3})
    g['config4_box_source'] = c                                               # staged face operands
    c = bench.workload_config('config4', N_MESH)
    c['optics']['crystal'].update({'check_bragg': True, 'rocking_fwhm': 2e-3, 'mesh_size': (24, 31),
                                   'mesh_coarse_size': (4, 5)})
    g['config4_bragg_24x31'] = c
    c = bench.workload_config('config4', N_MESH)
    c['optics']['crystal'].update({'trace_local': True, 'mesh_interpolate': False})
    g['config4_local_flat_normals'] = c
    g['config4_3e8'] = bench.workload_config('config4', 3 * N_MESH)          # three launches of the sorted path (27-bit offsets)
    # wavelengths drawn with the ray: a Doppler-shifted source and the config 5 plasma in front of the mesh (Bragg test on)
    c = bench.workload_config('config4', N_MESH)
    c['sources']['source']['velocity'] = [0.0, 3.0e4, 1.0e5]
    c['optics']['crystal'].update({'check_bragg': True, 'rocking_fwhm': 2e-3})
    g['config4_doppler_bragg'] = c
    c = bench.workload_config('config5', N_MESH)
    c['optics']['crystal'] = copy.deepcopy(bench.workload_config('config4', N_MESH)['optics']['crystal'])
    c['optics']['crystal'].update({'check_bragg': True, 'rocking_fwhm': 2e-3})
    g['config4_plasma_source'] = c
    return g


@pytest.mark.timeout(900)
@pytest.mark.parametrize('name', ['config4', 'config4_box_source', 'config4_bragg_24x31', 'config4_local_flat_normals',
                                  'config4_3e8', 'config4_doppler_bragg', 'config4_plasma_source'])
def test_sorted_mesh_path_changes_no_result(torch, name, monkeypatch):
    """The sorted mesh path (k_mesh_coarse -> counting sort by hit location -> k_trace in sorted mode) only changes the
    order in which rays are refined: same counters, images and found-id sets as the single-kernel path at 1e8 rays."""
    from xicsrt_b200 import _driver, config as xconfig
    cfg = _mesh_geometries()[name]
    cfg['general']['keep_history'] = False
    results = []
    for env in ({}, {'XRT_NO_MESH_SORT': '1'}):
        monkeypatch.delenv('XRT_NO_MESH_SORT', raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        per_seed = []
        for seed in SEEDS[:2]:
            tracer = _driver.Tracer(xconfig.get_config(xconfig.to_numpy(copy.deepcopy(cfg))), seed=seed)
            assert (tracer.scene.launch_info()['mesh_sort'] is not None) == (not env), name
            tracer.trace(1)
            packed_off = tracer.packed.clone()
            found, lost = tracer.select_ids(1, 64)
            assert torch.equal(tracer.packed, packed_off), f'{name}: history-on launch counts differ from history-off'
            per_seed.append((packed_off, found.clone(), lost.clone()))
            tracer.close()
        results.append(per_seed)
    for s, seed in enumerate(SEEDS[:2]):
        packed0, found0, lost0 = results[0][s]
        packed1, found1, lost1 = results[1][s]
        assert int(packed0[0]) in (N_MESH, 3 * N_MESH) or name == 'config4_plasma_source'       # Poisson total
        assert int(found0.numel()) == int(packed0[2]) > N_MESH // 1000
        assert torch.equal(packed1, packed0), f'{name} seed {seed}: counters / images differ without the sort'
        assert torch.equal(found1, found0), f'{name} seed {seed}: found-id set differs without the sort'
        assert torch.equal(torch.sort(lost1)[0], torch.sort(lost0)[0]), f'{name} seed {seed}: lost sample differs'


def _local_xy(res, elem):
    from xicsrt_b200 import elements
    from oracle import vecs
    _, param = elements.prepare_optic(res['config']['optics'][elem])
    h = res['found']['history'][elem]
    return vecs.point_to_local(param, h['origin']), h['wavelength']


@pytest.mark.timeout(1800)
def test_found_ray_distributions_match_oracle_at_1e8(torch):
    """RNG-driven end-to-end statistics with >= 1e8 rays on BOTH sides (config 1 / 2 geometry, history on)."""
    from scipy import stats
    import xicsrt_b200
    from test_gpu_statistics import chi2_two_sample, binomial_z
    runs, per_run = 100, 1_000_000
    n = runs * per_run
    cfg = scenes.get('sphere')
    cfg['sources']['source']['intensity'] = per_run
    cfg['general'].update({'number_of_runs': runs, 'keep_history': True, 'history_max_lost': 100, 'random_seed': 3})
    ref = oracle.raytrace_mp(copy.deepcopy(cfg), processes=min(os.cpu_count() or 1, 64))
    cfg = scenes.get('sphere')
    cfg['sources']['source']['intensity'] = n
    cfg['general'].update({'keep_history': True, 'history_max_lost': 100, 'random_seed': 77})
    got = xicsrt_b200.raytrace(cfg)

    assert got['total']['meta']['source']['num_out'] == n == ref['total']['meta']['source']['num_out']
    for elem in ('crystal', 'detector'):
        k1, k2 = got['total']['meta'][elem]['num_out'], ref['total']['meta'][elem]['num_out']
        z = binomial_z(k1, n, k2, n)
        assert abs(z) < 4.5, f'{elem}: {k1} vs {k2} of {n}, z = {z:.2f}'
    for elem in ('crystal', 'detector'):
        chi2, dof, p = chi2_two_sample(got['total']['image'][elem], ref['total']['image'][elem])
        assert p > 1e-4, f'{elem} image: chi2 = {chi2:.1f} for {dof} dof, p = {p:.2e}'
    n_found = got['total']['meta']['detector']['num_out']
    assert len(got['found']['history']['detector']['mask']) == n_found > 1_000_000
    for elem in ('crystal', 'detector'):
        xg, wg = _local_xy(got, elem)
        xr, wr = _local_xy(ref, elem)
        for label, a, b in (('x', xg[:, 0], xr[:, 0]), ('y', xg[:, 1], xr[:, 1]), ('wavelength', wg, wr)):
            ks = stats.ks_2samp(a, b)
            assert ks.pvalue > 1e-4, f'{elem} {label}: KS D = {ks.statistic:.2e}, p = {ks.pvalue:.2e}'
    # direction cosines of the found rays at the source as well (the cone sampler seen through the Bragg selection)
    dg, dr = got['found']['history']['source']['direction'], ref['found']['history']['source']['direction']
    for k in range(3):
        ks = stats.ks_2samp(dg[:, k], dr[:, k])
        assert ks.pvalue > 1e-4, f'source direction[{k}]: KS p = {ks.pvalue:.2e}'
