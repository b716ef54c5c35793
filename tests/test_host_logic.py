# -*- coding: utf-8 -*-
"""
Host-side logic that needs no GPU: config semantics, the class_name registry,
scene flattening, ray-id sharding, result combination and result files.
"""
import copy
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import scenes
from xicsrt_b200 import _driver, _lib as L, config as xconfig, elements, io as xio, registry, scene as xscene


def test_registry_covers_the_reference_class_names():
    optics = ['XicsrtOpticAperture', 'XicsrtOpticDetector', 'XicsrtOpticPlanarMirror', 'XicsrtOpticSphericalMirror',
              'XicsrtOpticCylindricalMirror', 'XicsrtOpticMeshMirror', 'XicsrtOpticPlanarCrystal',
              'XicsrtOpticSphericalCrystal', 'XicsrtOpticCylindricalCrystal', 'XicsrtOpticToroidalCrystal',
              'XicsrtOpticMeshCrystal', 'XicsrtOpticMeshSphericalCrystal', 'XicsrtOpticMeshCylindricalCrystal',
              'XicsrtOpticMeshToroidalCrystal', 'XicsrtOpticPlanarMosaicCrystal',
              'XicsrtOpticSphericalMosaicCrystal', 'XicsrtOpticMeshMosaicCrystal']
    sources = ['XicsrtSourceGeneric', 'XicsrtSourceDirected', 'XicsrtSourceFocused', 'XicsrtPlasmaGeneric',
               'XicsrtPlasmaCubic', 'XicsrtPlasmaToroidal', 'XicsrtPlasmaToroidalDatafile']
    for name in optics:
        assert registry.find(name)[0] == 'optics'
        assert registry.defaults(name)['class_name'] == name
    for name in sources:
        assert registry.find(name)[0] == 'sources'
    for name in ('XicsrtBundleFilter', 'XicsrtBundleFilterSightline'):
        assert registry.find(name)[0] == 'filters'
    with pytest.raises(NotImplementedError):
        registry.find('XicsrtPlasmaCylindrical')
    with pytest.raises(Exception, match='Could not find'):
        registry.find('XicsrtOpticCrystalSpherical')     # the stale name of examples/example_01.py:39


def test_strict_config_rejects_unknown_keys():
    cfg = scenes.get('sphere')
    cfg['optics']['crystal']['radius_of_curvature'] = 2.0
    with pytest.raises(Exception, match='User option not recognized'):
        oracle.raytrace(cfg)
    cfg['general']['strict_config_check'] = False
    full = xconfig.get_config(xconfig.to_numpy(cfg))
    xscene.prepare(full)


def test_default_axes_and_orthogonality_check():
    _, p = elements.prepare_optic({'class_name': 'XicsrtOpticDetector', 'zaxis': [0.0, 0.0, 1.0]})
    assert np.array_equal(p['xaxis'], [1.0, 0.0, 0.0])
    _, p = elements.prepare_optic({'class_name': 'XicsrtOpticDetector', 'zaxis': [0.0, 1.0, 0.0]})
    assert np.allclose(p['xaxis'], [-1.0, 0.0, 0.0])
    assert np.allclose(p['orientation'][1], np.cross(p['zaxis'], p['xaxis']))
    with pytest.raises(ValueError, match='not orthogonal'):
        elements.prepare_optic({'class_name': 'XicsrtOpticDetector', 'zaxis': [0, 0, 1], 'xaxis': [0, 0.1, 1]})


def flat(name):
    cfg = xconfig.get_config(xconfig.to_numpy(scenes.get(name)))
    _, sname, sparam, sfilters, optics = xscene.prepare(cfg)
    return xscene.flatten(sname, sparam, sfilters, optics) + (optics,)


def test_flatten_spectrometer_constants():
    desc, layout, keep, optics = flat('sphere')
    assert desc.n_optics == 2 and layout.n_rays == 10000
    c, d = desc.optics[0], desc.optics[1]
    assert c.shape == L.SHAPE['sphere'] and c.interact == L.INTERACT['crystal']
    assert c.flags & L.F_CHECK_BRAGG and c.flags & L.F_IMAGE and not (c.flags & L.F_CONVEX)
    assert c.two_d == 2 * 2.45676
    sigma = 48.070e-6 / (2 * np.sqrt(2 * np.log(2)))
    assert c.rock_two_sigma2 == 2 * sigma**2
    assert np.allclose(list(c.center), np.array([0.0, 0.0, 0.80374151]) + np.array([0.0, 0.59497864, -0.80374151]))
    assert (c.npix[0], c.npix[1], d.npix[0], d.npix[1]) == (100, 100, 100, 50)
    assert c.image_offset == 0 and d.image_offset == 10000 and layout.n_pixels == 15000
    assert list(c.half_size)[:2] == [0.1, 0.1]
    src = desc.source
    assert src.kind == L.SRC_FIXED_AXIS and src.wave == L.WAVE['normal'] and src.cone == L.CONE['isotropic']
    assert src.cone_par[0] == np.cos(np.radians(10.0))
    assert abs(src.wave_par[1] - 6.474e-4) < 1e-6          # Doppler sigma of SURVEY.md section 8d
    basis = np.array(list(src.axis_basis)).reshape(3, 3)
    assert np.allclose(basis @ basis.T, np.eye(3), atol=1e-14)


def test_flatten_variants():
    desc, _, keep, _ = flat('sphere_voigt')
    assert desc.source.wave == L.WAVE['table'] and desc.source.n_table == 1000
    desc, _, keep, _ = flat('sphere_step_box')
    assert desc.source.kind == L.SRC_FOCUSED and desc.source.wave == L.WAVE['uniform']
    assert desc.optics[0].rocking_type == L.ROCK['step'] and desc.optics[0].reflectivity == 0.8
    assert any(v != 0.0 for v in desc.source.velocity_c)
    desc, _, keep, _ = flat('torus_ft')
    assert desc.optics[0].root_idx == 2 and desc.optics[0].torus_major == pytest.approx(1.2)
    desc, _, keep, _ = flat('mosaic_sphere_cutoff')
    assert desc.optics[0].flags & L.F_MOSAIC_CUTOFF and desc.optics[0].mosaic_depth == 5
    desc, _, keep, _ = flat('apertures')
    assert desc.optics[0].n_aperture == 8 and desc.optics[1].n_aperture == 1
    ap = desc.optics[0].apertures
    assert ap[1].shape == L.AP_SHAPE['ellipse'] and ap[1].logic == L.AP_LOGIC['not']
    assert ap[4].shape == L.AP_SHAPE['triangle'] and list(ap[4].vert)[:2] == [-0.01, 0.03]
    desc, _, keep, _ = flat('local_frames')
    assert desc.optics[0].flags & L.F_TRACE_LOCAL and desc.optics[0].flags & L.F_HAS_ZSIZE


def test_check_bragg_uses_identity_like_the_reference():
    cfg = scenes.get('sphere')
    cfg['optics']['crystal']['check_bragg'] = False
    desc, *_ = xscene.flatten(*xscene.prepare(xconfig.get_config(xconfig.to_numpy(cfg)))[1:])
    assert not (desc.optics[0].flags & L.F_CHECK_BRAGG)


@pytest.mark.parametrize('n,world', [(10, 1), (10, 3), (1000000007, 8), (5, 8), (0, 4)])
def test_shard_ranges_partition_the_ray_ids(n, world):
    pieces = [_driver.shard_range(n, r, world) for r in range(world)]
    assert pieces[0][0] == 0
    for (b0, c0), (b1, _) in zip(pieces, pieces[1:]):
        assert b0 + c0 == b1
    assert pieces[-1][0] + pieces[-1][1] == n
    counts = [c for _, c in pieces]
    assert max(counts) - min(counts) <= 1


def test_combine_matches_oracle_combine():
    cfg = scenes.get('two_iter_two_runs')
    res = oracle.raytrace(cfg)
    # split it back into per-run results and recombine with the product's combine
    runs = [oracle.raytrace_single(c, _internal=True) for c in oracle.driver._run_configs(scenes.get('two_iter_two_runs'))[1]]
    got = _driver.combine_raytrace(runs)
    for name in res['total']['meta']:
        assert got['total']['meta'][name]['num_out'] == res['total']['meta'][name]['num_out']
        if res['total']['image'].get(name) is not None:
            assert np.array_equal(got['total']['image'][name], res['total']['image'][name])
        for kind in ('found', 'lost'):
            for key in _driver.RAY_KEYS:
                assert np.array_equal(got[kind]['history'][name][key], res[kind]['history'][name][key], equal_nan=True)


def test_run_seeds_are_cumulative():
    cfg = xconfig.get_config(scenes.get('sphere'))
    cfg['general']['number_of_runs'] = 4
    cfg['general']['random_seed'] = 5
    _, runs = oracle.driver._run_configs(cfg)
    assert [r['general']['random_seed'] for r in runs] == [5, 6, 8, 11]
    assert [r['general']['output_run_suffix'] for r in runs] == ['0000', '0001', '0002', '0003']


@pytest.mark.parametrize('ext', ['.json', '.pkl'])
def test_results_files_round_trip(tmp_path, ext):
    cfg = scenes.get('sphere')
    cfg['sources']['source']['intensity'] = 500
    res = oracle.raytrace(cfg)
    res['config']['general'].update({'output_path': str(tmp_path), 'results_ext': ext, 'output_suffix': 'x'})
    xio.save_results(res)
    path = xio.generate_filename(res['config'], 'results')
    assert os.path.basename(path) == 'xicsrt_results_x' + ext
    back = xio.load_results(config=res['config'])
    assert back['total']['meta']['detector']['num_out'] == res['total']['meta']['detector']['num_out']
    assert np.array_equal(np.asarray(back['total']['image']['detector']), res['total']['image']['detector'])
    assert np.array_equal(np.asarray(back['found']['history']['crystal']['origin']),
                          res['found']['history']['crystal']['origin'])
    with pytest.raises(FileExistsError):
        xio.save_results(res)


def test_images_are_saved_rotated(tmp_path):
    from PIL import Image
    cfg = scenes.get('sphere')
    cfg['sources']['source']['intensity'] = 2000
    res = oracle.raytrace(cfg)
    res['config']['general'].update({'output_path': str(tmp_path)})
    xio.save_images(res)
    img = np.array(Image.open(tmp_path / 'xicsrt_detector.tif'))
    assert img.shape == (50, 100)
    assert np.array_equal(img, np.rot90(res['total']['image']['detector']))


def test_fp32_broad_phase_error_budget():
    """
    The FP32 broad phase of the spectrometer variant (spectro_cull32 in csrc/xrt.cu) is conservative only if its
    single-precision sin(theta_i) = thc / R stays within the margin kn32[21] - cull_err = 2e-5 max(1, |C-O|^2/R^2) of
    the FP64 value.  Here the same arithmetic is restated in numpy float32 (correctly rounded sqrt / sin / cos; the MUFU
    units add at most 2^-21 absolute to sin / cos and one ulp to sqrt) on the benchmark geometry and on the stress
    geometries of the GPU test, with uniforms that include the cone axis and the cone edge; for chords within 0.05 of
    sin(theta_B) -- a wrong rejection is only possible within ~2e-4 of it -- the worst error must use less than a
    fifth of the margin.
    """
    import bench
    from xicsrt_b200 import config as xconfig, scene as xscene
    f32 = np.float32
    rng = np.random.default_rng(11)
    n = 400000
    a = np.concatenate([rng.random(n), 1.0 - 10.0**rng.uniform(-12, -1, n // 4), 10.0**rng.uniform(-12, -1, n // 4)])
    b = rng.random(len(a))
    variants = {'default': {}, 'scaled': {'scale': 3.0}, 'wide': {'spread': np.radians(75.0)},
                'off_rowland': {'origin': [0.01, -0.02, 0.15]}}
    for name, mod in variants.items():
        cfg = bench.spectrometer(1000)
        if 'scale' in mod:
            for o in cfg['optics'].values():
                o['origin'] = [mod['scale'] * v for v in o['origin']]
            cfg['optics']['crystal']['radius'] = mod['scale']
        if 'spread' in mod:
            cfg['sources']['source']['spread'] = float(mod['spread'])
        if 'origin' in mod:
            cfg['sources']['source']['origin'] = mod['origin']
        _, sname, sp, sf, optics = xscene.prepare(xconfig.get_config(xconfig.to_numpy(cfg)))
        cp = optics['crystal']
        basis = xscene.cone_basis(sp['direction'] if sp.get('direction') is not None else sp['zaxis'], sp['xaxis'], sp['zaxis'])
        cs0 = np.cos(np.atleast_1d(sp['spread'])[0])
        Lc = np.asarray(cp['center'], dtype=np.float64) - np.asarray(sp['origin'], dtype=np.float64)
        R = float(cp['radius'])
        # ---- FP64, as spectro_stage_a
        z = cs0 + (1.0 - cs0) * a
        rho = np.sqrt(1.0 - z * z)
        d = (rho * np.cos(2 * np.pi * b))[:, None] * basis[0] + (rho * np.sin(2 * np.pi * b))[:, None] * basis[1] + z[:, None] * basis[2]
        tca = d @ Lc
        thc2 = R * R - (Lc @ Lc - tca * tca)
        hit = thc2 > 0
        sI64 = np.sqrt(np.where(hit, thc2, 1.0)) / R
        # ---- FP32, as spectro_cull32 (1 - a to 2^-24 relative, 24-bit azimuth, w (2 - w), constants rounded to float)
        na = np.floor((1.0 - a) * 2.0**32)                # top 32 bits of 1 - a; rays with na < 256 are left to FP64
        one_minus_a = ((na + 0.5) * 2.0**-32).astype(f32)
        b24 = (np.floor(b * 2**23) / 2**23).astype(f32)   # 23-bit azimuth (float in [1, 2) bit trick)
        w = (f32(1.0 - cs0) * one_minus_a).astype(f32)
        z32 = (f32(1) - w).astype(f32)
        rho32 = np.sqrt((w * (f32(2) - w)).astype(f32)).astype(f32)
        ang = (f32(6.283185307179586) * (b24 - f32(0.5))).astype(f32)
        lx, ly = (-rho32 * np.cos(ang).astype(f32)).astype(f32), (-rho32 * np.sin(ang).astype(f32)).astype(f32)
        m32 = (basis @ Lc).astype(f32)                    # point source: tca = l . (basis L), host-computed in FP64
        tca32 = (lx * m32[0] + ly * m32[1] + z32 * m32[2]).astype(f32)
        d2 = (f32(Lc @ Lc) - tca32 * tca32).astype(f32)
        t2 = (f32(R * R) - d2).astype(f32)
        sI32 = (np.sqrt(np.where(t2 > 0, t2, f32(1))).astype(f32) * f32(1.0 / R)).astype(f32)
        # a wrong rejection needs |sI32 - sI| > margin where sI is within ~T of sin(theta_B): only chords near the
        # Bragg angle matter (elsewhere the gap is hundreds of margins wide)
        sB0 = float(sp['wavelength']) / (2.0 * float(cp['crystal_spacing']))
        both = hit & (t2 > 0) & (np.abs(sI64 - sB0) < 0.05) & (na >= 256)
        err = np.abs(sI32.astype(np.float64) - sI64)[both].max()
        mufu = 2.0**-21 * 2.0 * np.abs(Lc).sum() / R + 2.4e-7          # sin / cos through rho <= 1 into tca, thc; sqrt ulps
        margin = 2e-5 * max(1.0, (Lc @ Lc) / (R * R))
        assert Lc @ Lc <= 4 * R * R
        assert err + mufu < margin / 5, (name, err, mufu, margin)


def test_fp32_mosaic_broad_phase_error_budget():
    """
    The mosaic broad phase (k_mosaic32 in csrc/xrt_kernels.cuh) rejects a crystallite layer when its single-precision
    sin(theta_i) = |x D.r_0 + y D.r_1 + D.n| / sqrt(x^2 + y^2 + 1) is farther from sin(theta_B) than the rocking curve
    allows; the frame (n, r_0, r_1) of mosaic_normal at the intersection point and the three dot products come from the
    FP32 ray (n = (L - t D) / R, r_0 = unit(n_y, n_z - n_x, -n_y), r_1 = unit(n x r_0)).  The margin is that of
    k_cull32, 2e-5 max(1, |C - O|^2 / R^2), plus stage S's 2e-6.  Restated here in numpy float32 for config 3 and two
    stress variants: the error of sin(theta_i) must stay within a fifth of the margin for every crystallite offset.
    """
    import bench
    from xicsrt_b200 import config as xconfig, scene as xscene
    f32 = np.float32
    rng = np.random.default_rng(23)
    n = 300000
    a = np.concatenate([rng.random(n), 1.0 - 10.0**rng.uniform(-12, -1, n // 4), 10.0**rng.uniform(-12, -1, n // 4)])
    b = rng.random(len(a))
    for name, mod in {'config3': {}, 'scaled': {'scale': 3.0}, 'off_rowland': {'origin': [0.01, -0.02, 0.15]}}.items():
        cfg = bench.workload_config('config3', 1000)
        if 'scale' in mod:
            for o in cfg['optics'].values():
                o['origin'] = [mod['scale'] * v for v in o['origin']]
            cfg['optics']['crystal']['radius'] = mod['scale']
        if 'origin' in mod:
            cfg['sources']['source']['origin'] = mod['origin']
        _, sname, sp, sf, optics = xscene.prepare(xconfig.get_config(xconfig.to_numpy(cfg)))
        cp = optics['crystal']
        basis = xscene.cone_basis(sp['direction'] if sp.get('direction') is not None else sp['zaxis'], sp['xaxis'], sp['zaxis'])
        cs0 = np.cos(np.atleast_1d(sp['spread'])[0])
        Lc = np.asarray(cp['center'], dtype=np.float64) - np.asarray(sp['origin'], dtype=np.float64)
        R = float(cp['radius'])
        sig = np.sin(float(cp['mosaic_spread']) / (2.0 * np.sqrt(2.0 * np.log(2.0)))) if 'mosaic_spread' in cp else 3e-3
        # crystallite offsets: Gaussian with the mosaic width, plus far tails
        xy = np.concatenate([rng.normal(0.0, sig, (len(a) // 2, 2)), rng.normal(0.0, 5 * sig, (len(a) - len(a) // 2, 2))])

        def frame(d, t, L, inv_r, dt):
            nrm = ((L - t[:, None] * d) * inv_r).astype(dt)
            r0 = np.stack([nrm[:, 1], nrm[:, 2] - nrm[:, 0], -nrm[:, 1]], axis=1).astype(dt)
            r0 = (r0 / np.sqrt((r0 * r0).sum(axis=1, dtype=dt))[:, None]).astype(dt)
            r1 = np.cross(nrm, r0).astype(dt)
            r1 = (r1 / np.sqrt((r1 * r1).sum(axis=1, dtype=dt))[:, None]).astype(dt)
            return (d * r0).sum(axis=1, dtype=dt), (d * r1).sum(axis=1, dtype=dt), (d * nrm).sum(axis=1, dtype=dt), r0

        # ---- FP64 (generate_geometry, hit_sphere, mosaic_normal)
        z = cs0 + (1.0 - cs0) * a
        rho = np.sqrt(1.0 - z * z)
        d = (rho * np.cos(2 * np.pi * b))[:, None] * basis[0] + (rho * np.sin(2 * np.pi * b))[:, None] * basis[1] + z[:, None] * basis[2]
        tca = d @ Lc
        thc2 = R * R - (Lc @ Lc - tca * tca)
        hit = thc2 > 0
        t = tca + np.sqrt(np.where(hit, thc2, 1.0))
        dr0, dr1, dn, r0_64 = frame(d, t, Lc[None, :], 1.0 / R, np.float64)
        sI64 = np.abs(xy[:, 0] * dr0 + xy[:, 1] * dr1 + dn) / np.sqrt(xy[:, 0]**2 + xy[:, 1]**2 + 1.0)
        # ---- FP32 (cull32_ray<point, FULL> + the frame of k_mosaic32)
        na = np.floor((1.0 - a) * 2.0**32)
        one_minus_a = ((na + 0.5) * 2.0**-32).astype(f32)
        b24 = (np.floor(b * 2**23) / 2**23).astype(f32)
        w = (f32(1.0 - cs0) * one_minus_a).astype(f32)
        z32 = (f32(1) - w).astype(f32)
        rho32 = np.sqrt((w * (f32(2) - w)).astype(f32)).astype(f32)
        ang = (f32(6.283185307179586) * (b24 - f32(0.5))).astype(f32)
        lx, ly = (-rho32 * np.cos(ang).astype(f32)).astype(f32), (-rho32 * np.sin(ang).astype(f32)).astype(f32)
        b32 = basis.astype(f32)
        d32 = (lx[:, None] * b32[0] + ly[:, None] * b32[1] + z32[:, None] * b32[2]).astype(f32)
        m32 = (basis @ Lc).astype(f32)
        tca32 = (lx * m32[0] + ly * m32[1] + z32 * m32[2]).astype(f32)
        t2 = (f32(R * R) - (f32(Lc @ Lc) - tca32 * tca32).astype(f32)).astype(f32)
        t32 = (tca32 + np.sqrt(np.where(t2 > 0, t2, f32(1))).astype(f32)).astype(f32)
        dr0_32, dr1_32, dn_32, r0_32 = frame(d32, t32, Lc.astype(f32)[None, :], f32(1.0 / R), f32)
        x32, y32 = xy[:, 0].astype(f32), xy[:, 1].astype(f32)
        sI32 = (np.abs(x32 * dr0_32 + y32 * dr1_32 + dn_32) / np.sqrt(x32 * x32 + y32 * y32 + f32(1))).astype(f32)
        # the kernel leaves rays with a nearly degenerate frame (|r_0 before normalisation|^2 <= 0.01) to the FP64 path
        nrm64 = (Lc[None, :] - t[:, None] * d) / R
        a2 = 2 * nrm64[:, 1]**2 + (nrm64[:, 2] - nrm64[:, 0])**2
        both = hit & (t2 > 0) & (na >= 256) & (a2 > 0.01)
        assert both.sum() > 0.5 * len(a)
        err = np.abs(sI32.astype(np.float64) - sI64)[both].max()
        mufu = 2.0**-21 * 2.0 * np.abs(Lc).sum() / R + 1e-6      # sin / cos / sqrt / rsqrt of the MUFU units
        margin = 2e-5 * max(1.0, (Lc @ Lc) / (R * R)) + 2e-6
        assert err + mufu < margin / 5, (name, err, mufu, margin)      # measured: 8-9 % of the margin


def test_fp32_broad_phase_error_budget_extended_sources():
    """
    The same budget for the generalised broad phase (cull32_ray<CULL_BOX / CULL_FOCUSED / CULL_BUNDLES> in
    csrc/xrt_kernels.cuh): per-ray origin inside a box or plasma voxel, cone axis towards a target with the basis
    o_1 = unit(axis x (xaxis + zaxis)), o_2 = axis x o_1 built in float32, Doppler factor 1 - v.D / c.  The margin
    added per ray is 2e-5 max(1, |C - O|^2 / R^2); sin(theta_i) and sin(theta_B) together must stay within a fifth
    of it for chords near the Bragg angle.
    """
    f32 = np.float32
    rng = np.random.default_rng(5)
    n = 300000
    a = np.concatenate([rng.random(n), 1.0 - 10.0**rng.uniform(-12, -1, n // 4), 10.0**rng.uniform(-12, -1, n // 4)])
    b = rng.random(len(a))
    m = len(a)
    C = np.array([0.0, 0.59497864, 0.0])                       # centre of curvature of geometry G (R = 1)
    R, lam0, two_d = 1.0, 3.9492, 2 * 2.45676
    target = np.array([0.0, 0.0, 0.80374151])
    xa, za = np.array([1.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0])
    cases = {
        'box_1mm_fixed_axis': dict(org=np.zeros(3), ext=1e-3, spread=np.radians(10.0), focused=False, vel=np.zeros(3)),
        'box_10cm_fixed_axis': dict(org=np.zeros(3), ext=0.1, spread=np.radians(10.0), focused=False, vel=np.zeros(3)),
        'focused_2cm': dict(org=np.zeros(3), ext=0.02, spread=np.radians(8.0), focused=True, vel=np.zeros(3)),
        'plasma_voxels_10cm_moving': dict(org=None, ext=1e-3, spread=np.radians(2.0), focused=True,
                                          vel=np.array([1e-4, -3e-4, 2e-4])),
        # source 2.5 m in front of the crystal on the central ray: |C - O|^2 = 3.2 R^2 (the margin scales with it)
        'far_source_1.8R': dict(org=np.array([0.0, 0.0, 0.80374151 - 2.5]), ext=0.01,
                                spread=np.radians(3.0), focused=True, vel=np.zeros(3)),
    }
    for name, cs in cases.items():
        org = cs['org']
        if org is None:                                         # one voxel origin per ray inside the 10 cm plasma cube
            org = rng.uniform(-0.05, 0.05, (m, 3))
        org = np.broadcast_to(org, (m, 3))
        u = rng.random((m, 3))
        off = cs['ext'] * (u - 0.5)                              # source axes = identity here
        O = org + off
        cs0 = np.cos(cs['spread'])
        # ---- FP64 (generate_geometry + hit_sphere)
        z = cs0 + (1.0 - cs0) * a
        rho = np.sqrt(1.0 - z * z)
        lx, ly = rho * np.cos(2 * np.pi * b), rho * np.sin(2 * np.pi * b)
        if cs['focused']:
            ax = target - O
            ax /= np.linalg.norm(ax, axis=1)[:, None]
            o1 = np.cross(ax, xa) + np.cross(ax, za)
            o1 /= np.linalg.norm(o1, axis=1)[:, None]
            o2 = np.cross(ax, o1)
            o2 /= np.linalg.norm(o2, axis=1)[:, None]
        else:
            ax = np.broadcast_to(za, (m, 3))
            o1 = np.cross(za, xa) + np.cross(za, za)
            o1 = np.broadcast_to(o1 / np.linalg.norm(o1), (m, 3))
            o2 = np.cross(ax, o1)
        d = lx[:, None] * o2 + ly[:, None] * o1 + z[:, None] * ax
        L = C - O
        tca = np.einsum('ij,ij->i', L, d)
        ll = np.einsum('ij,ij->i', L, L)
        thc2 = R * R - (ll - tca * tca)
        hit = thc2 > 0
        sI64 = np.sqrt(np.where(hit, thc2, 1.0)) / R
        dop64 = 1.0 - d @ cs['vel']
        sB64 = lam0 * dop64 / two_d
        # ---- FP32 (cull32_ray): differences C - org, target - org formed in FP64 once, then float
        Lb, Tb = (C - org).astype(f32), (target - org).astype(f32)
        o32 = (f32(cs['ext']) * ((np.floor(u * 2**23) / 2**23).astype(f32) - f32(0.5))).astype(f32)
        L32 = (Lb - o32).astype(f32)
        T32 = (Tb - o32).astype(f32)
        na = np.floor((1.0 - a) * 2.0**32)
        one_minus_a = ((na + 0.5) * 2.0**-32).astype(f32)
        b24 = (np.floor(b * 2**23) / 2**23).astype(f32)
        w = (f32(1.0 - cs0) * one_minus_a).astype(f32)
        z32 = (f32(1) - w).astype(f32)
        rho32 = np.sqrt((w * (f32(2) - w)).astype(f32)).astype(f32)
        ang = (f32(6.283185307179586) * (b24 - f32(0.5))).astype(f32)
        lx32, ly32 = (-rho32 * np.cos(ang).astype(f32)).astype(f32), (-rho32 * np.sin(ang).astype(f32)).astype(f32)
        if cs['focused']:
            it = (f32(1) / np.sqrt(np.einsum('ij,ij->i', T32, T32).astype(f32))).astype(f32)
            ax32 = (T32 * it[:, None]).astype(f32)
            p32 = np.cross(ax32, (xa + za).astype(f32)).astype(f32)
            ip = (f32(1) / np.sqrt(np.einsum('ij,ij->i', p32, p32).astype(f32))).astype(f32)
            p32 = (p32 * ip[:, None]).astype(f32)
            q32 = np.cross(ax32, p32).astype(f32)
        else:
            ax32, p32, q32 = ax.astype(f32), o1.astype(f32), o2.astype(f32)
        d32 = (lx32[:, None] * q32 + ly32[:, None] * p32 + z32[:, None] * ax32).astype(f32)
        tca32 = np.einsum('ij,ij->i', L32, d32).astype(f32)
        ll32 = np.einsum('ij,ij->i', L32, L32).astype(f32)
        d2 = (ll32 - tca32 * tca32).astype(f32)
        t2 = (f32(R * R) - d2).astype(f32)
        sI32 = (np.sqrt(np.where(t2 > 0, t2, f32(1))).astype(f32) * f32(1.0 / R)).astype(f32)
        dop32 = (f32(1) - (d32 @ cs['vel'].astype(f32)).astype(f32)).astype(f32)
        sB32 = (f32(lam0) * dop32 * f32(1.0 / two_d)).astype(f32)
        near = hit & (t2 > 0) & (np.abs(sI64 - sB64) < 0.05) & (ll <= 4 * R * R) & (na >= 256)
        assert near.sum() > 1000, name
        err = (np.abs(sI32.astype(np.float64) - sI64) + np.abs(sB32.astype(np.float64) - sB64))[near]
        q = ll[near] / (R * R)
        margin = 2e-5 * np.maximum(1.0, q)
        mufu = 2.0**-21 * 2.0 * np.sqrt(ll[near]) / R + 6e-7     # sin / cos / rsqrt units through tca, thc, the basis
        worst = np.max((err + mufu) / margin)
        assert worst < 0.2, (name, worst)


def test_mosaic_scan_error_budget():
    """
    Stage S of the fused kernel (stage_mosaic in csrc/xrt_kernels.cuh) pre-tests every crystallite layer in float32:
    x, y by Box-Muller from the layer's Philox words (1 - u1 from its top 32 bits, a 23-bit angle), then
    sin(theta_i) = |x D.r_0 + y D.r_1 + D.n| / sqrt(x^2 + y^2 + 1) with the three dot products rounded from FP64.
    Restated here in numpy float32 against the FP64 arithmetic of mosaic_xy / mosaic_normal / bragg_dtheta: the error
    on sin(theta_i) must stay within a quarter of the margin mosaic_err = 2e-6 (the MUFU units add ~4e-7 at most:
    2^-22 relative on rsqrt, 2^-21 absolute on sin / cos scaled by sin(sigma) r < 0.02).
    """
    f32 = np.float32
    rng = np.random.default_rng(17)
    m = 400000
    n = rng.normal(size=(m, 3)); n /= np.linalg.norm(n, axis=1)[:, None]
    d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    for sigma_deg in (0.4 / 2.3548, 2.0):
        s = np.sin(np.radians(sigma_deg))
        w1 = rng.integers(0, 2**32, m, dtype=np.uint64)        # Philox word x (top 32 bits of the 40-bit radius uniform)
        w1[: m // 8] = 2**32 - 1 - rng.integers(256, 2**22, m // 8).astype(np.uint64)   # deep tail: 1 - u1 tiny
        lo8 = rng.integers(0, 256, m, dtype=np.uint64)
        ang24 = rng.integers(0, 2**24, m, dtype=np.uint64)
        # ---- FP64 (mosaic_xy with box_muller(u01_40, u01_24), mosaic_normal, |D.n_m|)
        u1 = (w1.astype(np.float64) * 256 + lo8) / 2.0**40
        u2 = ang24 / 2.0**24
        r = np.sqrt(-2.0 * np.log(1.0 - u1))
        x, y = s * r * np.cos(2 * np.pi * u2), s * r * np.sin(2 * np.pi * u2)
        r0 = np.stack([n[:, 1], n[:, 2] - n[:, 0], -n[:, 1]], axis=1); r0 /= np.linalg.norm(r0, axis=1)[:, None]
        r1 = np.cross(n, r0); r1 /= np.linalg.norm(r1, axis=1)[:, None]
        inv = 1.0 / np.sqrt(x * x + y * y + 1.0)
        nm = (x * inv)[:, None] * r0 + (y * inv)[:, None] * r1 + inv[:, None] * n
        sI64 = np.abs(np.einsum('ij,ij->i', d, nm))
        # ---- float32 (stage_mosaic)
        na = (2**32 - 1 - w1).astype(np.float64)
        omu = (na.astype(f32) * f32(2.0**-32) + f32(2.0**-33)).astype(f32)
        rr = np.sqrt((f32(-1.3862943611198906) * np.log2(omu).astype(f32)).astype(f32)).astype(f32)
        a23 = (np.floor(ang24 / 2) / 2.0**23).astype(f32)
        ang = (f32(6.283185307179586) * (a23 - f32(0.5))).astype(f32)
        x32 = (-f32(s) * rr * np.cos(ang).astype(f32)).astype(f32)
        y32 = (-f32(s) * rr * np.sin(ang).astype(f32)).astype(f32)
        dr0 = np.einsum('ij,ij->i', d, r0).astype(f32)
        dr1 = np.einsum('ij,ij->i', d, r1).astype(f32)
        dn = np.einsum('ij,ij->i', d, n).astype(f32)
        t = (x32 * dr0 + (y32 * dr1 + dn).astype(f32)).astype(f32)
        q = (x32 * x32 + (y32 * y32 + f32(1)).astype(f32)).astype(f32)
        sI32 = (np.abs(t) / np.sqrt(q).astype(f32)).astype(f32)
        ok = na >= 65536                                        # below that the lane decides exactly
        err = np.abs(sI32.astype(np.float64) - sI64)[ok]
        assert err.max() < 5e-7, (sigma_deg, err.max())


def test_bragg_pretest_inequalities_are_conservative():
    """
    The two levels of the Bragg pre-test (bragg_cull_test / bragg_cull_uniform in csrc/xrt_trace.cuh) restated in
    numpy: whenever the bound rejects a ray, the exact rocking-curve test -- dtheta = asin(sB) - asin(sI),
    p = exp(-dtheta^2 / 2 sigma^2) * reflectivity, reflected iff p >= u -- rejects it too; and the bound is not
    vacuous (it rejects most rays that fail by a wide margin).
    """
    rng = np.random.default_rng(5)
    n = 2000000
    for fwhm, refl, err in ((48.07e-6, 1.0, 2.7e-7), (9e-6, 0.37, 1e-9), (400e-6, 1.0, 5e-6)):
        sigma = fwhm / (2 * np.sqrt(2 * np.log(2)))
        two_sigma2 = 2 * sigma * sigma
        sB = rng.uniform(0.05, 0.999, n)
        # sI around sB on every scale from far inside to far outside the rocking curve, plus measurement error <= err
        sI = np.clip(sB + rng.choice([-1.0, 1.0], n) * 10.0**rng.uniform(-8, -1.5, n), 0.0, 1.0)
        sB_seen = sB + rng.uniform(-err, err, n)            # what the pre-test sees (approximate deviate)
        dtheta = np.arcsin(sB) - np.arcsin(sI)
        x = dtheta * dtheta / two_sigma2
        u = rng.random(n)
        passes = np.exp(-x) * refl >= u
        # first level: T = 1.05 * sqrt(40 * two_sigma2) + 2e-6
        T = 1.05 * np.sqrt(40.0 * two_sigma2) + 2e-6
        gap = np.abs(sB_seen - sI)
        diff = gap - err
        c2 = 2.0 * gap + (1.0 - sI * sI)
        cull1 = (diff > 0) & (diff * diff > T * T * c2)
        assert not np.any(cull1 & (x < 40.0)), 'first level rejected a ray with x < 40'
        assert not np.any(cull1 & passes)
        # second level: ln(refl / u) in float32, 1e-3 + 0.1 % too large
        lim = (np.float32(0.6931471805599453) * (np.log2(np.float32(refl)) - np.log2(u.astype(np.float32)))).astype(np.float32)
        bound = (np.abs(lim) * np.float32(1e-3) + (lim + np.float32(1e-3))).astype(np.float64) * two_sigma2
        cull2 = (diff > 0) & (diff * diff > bound * c2)
        assert not np.any(cull2 & passes), 'second level rejected a ray the exact test reflects'
        fails_wide = x > 60.0
        assert cull1[fails_wide].mean() > 0.95
        assert cull2[~passes & (x > 3.0 * np.maximum(np.log(refl / u), 0.5) + 1.0)].mean() > 0.9


def test_scene_cache_key(monkeypatch):
    """
    The content key under which raytrace_single keeps a prepared scene (_driver._scene_key): equal for equal element
    configs whatever the seed, different for any changed element value, for a changed XRT_* switch (the library reads
    them at scene creation and at launch), rank, world size or device; None -- never cached -- for plasma sources, Poisson
    ray counts, file-backed tables, and with the cache switched off.
    """
    import types
    fake = types.SimpleNamespace(cuda=types.SimpleNamespace(current_device=lambda: 0))
    monkeypatch.setattr(_driver, '_torch', lambda: fake)
    monkeypatch.delenv('XRT_SCENE_CACHE', raising=False)
    monkeypatch.delenv('XRT_NO_CULL', raising=False)

    def cfg_of(name, **general):
        c = scenes.get(name)
        c['general'].update(general)
        return xconfig.get_config(xconfig.to_numpy(c))

    k0 = _driver._scene_key(cfg_of('sphere', random_seed=1), 0, 1)
    assert isinstance(k0, bytes)
    assert _driver._scene_key(cfg_of('sphere', random_seed=99, output_run_suffix='0007'), 0, 1) == k0
    moved = cfg_of('sphere')
    moved['optics']['detector']['origin'] = moved['optics']['detector']['origin'] + 1e-9
    assert _driver._scene_key(moved, 0, 1) != k0
    more = cfg_of('sphere')
    more['sources']['source']['intensity'] = more['sources']['source']['intensity'] + 1
    assert _driver._scene_key(more, 0, 1) != k0
    assert _driver._scene_key(cfg_of('sphere'), 1, 2) != k0
    monkeypatch.setenv('XRT_NO_CULL', '1')
    assert _driver._scene_key(cfg_of('sphere'), 0, 1) != k0
    monkeypatch.delenv('XRT_NO_CULL')
    assert _driver._scene_key(cfg_of('sphere'), 0, 1) == k0
    fake.cuda.current_device = lambda: 3
    assert _driver._scene_key(cfg_of('sphere'), 0, 1) != k0
    fake.cuda.current_device = lambda: 0
    for name in ('plasma_cubic', 'plasma_toroidal', 'plasma_datafile', 'sphere_rocking_file'):
        assert _driver._scene_key(cfg_of(name), 0, 1) is None, name
    poisson = cfg_of('sphere')
    poisson['sources']['source']['use_poisson'] = True
    assert _driver._scene_key(poisson, 0, 1) is None
    monkeypatch.setenv('XRT_SCENE_CACHE', '0')
    assert _driver._scene_key(cfg_of('sphere'), 0, 1) is None
