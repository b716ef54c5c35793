# -*- coding: utf-8 -*-
"""
Plasma (extended) sources.

CPU: the vectorised bundle table of the product (xicsrt_b200/plasma.py) against the oracle's
restatement of the reference's per-bundle loop, fed from the same MT19937 stream.
GPU (``-m gpu``): rays generated from the uploaded bundle table stay inside their bundle's
voxel and emission cone, carry the bundle's Doppler shift, and end-to-end counts agree
statistically with the oracle.
"""
import copy
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import scenes, sources as osources
from oracle.stream import LegacyStream
from xicsrt_b200 import _lib as L, config as xconfig, plasma, scene as xscene, voigt


class StreamAdapter:
    """The product's host-random interface on top of the oracle's legacy stream."""

    def __init__(self, stream):
        self.s = stream

    def uniform(self, lo, hi, n):
        return self.s.uniform(lo, hi, n)

    def poisson(self, lam):
        return self.s.poisson(lam)

    def poisson_array(self, lam):
        return np.array([self.s.poisson(x) for x in lam], dtype=np.int64)


def prepared(name):
    cfg = xconfig.get_config(xconfig.to_numpy(scenes.get(name)))
    return xscene.prepare(cfg)


@pytest.mark.parametrize('name', ['plasma_cubic', 'plasma_cubic_poisson', 'plasma_toroidal', 'plasma_datafile',
                                  'plasma_voigt', 'plasma_flat_xy'])
def test_bundle_properties_match_oracle(name):
    _, sname, sparam, sfilters, optics = prepared(name)
    seed = scenes.get(name)['general']['random_seed']
    got = plasma.bundle_properties(sparam, sfilters, StreamAdapter(LegacyStream(seed)))
    ref = osources.plasma_bundles(copy.deepcopy(sparam), LegacyStream(seed), sfilters)
    for key in ('origin', 'spread', 'solid_angle', 'mask'):
        assert np.array_equal(got[key], ref[key]), key
    m = ref['mask']
    for key in ('temperature', 'emissivity', 'velocity'):
        np.testing.assert_allclose(got[key][m], ref[key][m], rtol=1e-14, atol=0, err_msg=key)
    np.testing.assert_allclose(plasma.bundle_intensity(sparam, got)[m], osources.bundle_intensity(sparam, ref)[m],
                               rtol=1e-14)


@pytest.mark.parametrize('name', ['plasma_cubic', 'plasma_toroidal'])
def test_bundle_table_matches_reference_ray_counts(golden_dir, name):
    """Without Poisson the rays per bundle are int(intensity): the table's total is the reference's ray count."""
    _, sname, sparam, sfilters, optics = prepared(name)
    seed = scenes.get(name)['general']['random_seed']
    b = plasma.build_bundles(sparam, sfilters, StreamAdapter(LegacyStream(seed)))
    gold = np.load(f'{golden_dir}/{name}.npz')
    assert b['n_rays'] == len(gold['iter/source/mask'])
    # the rays of the reference sit in their bundle's voxel, in table order
    org = gold['iter/source/origin']
    begin = 0
    half = sparam['voxel_size'] / 2
    for row, end in zip(b['table'], b['end']):
        local = org[begin:int(end)] - row['origin']
        assert np.all(np.abs(local) <= half * (1 + 1e-12))
        begin = int(end)
    assert C.sizeof(L.XrtBundle) == b['table'].dtype.itemsize


def test_linewidth_and_cone_models_of_the_bundle_sources():
    """Host side of the per-bundle Voigt tables and cone parameters (xicsrt_b200/plasma.py)."""
    _, sname, sparam, sfilters, optics = prepared('plasma_voigt')
    assert plasma.line_model(sparam) == 'table'
    seed = scenes.get('plasma_voigt')['general']['random_seed']
    b = plasma.build_bundles(sparam, sfilters, StreamAdapter(LegacyStream(seed)))
    n = len(b['end'])
    assert b['voigt_x'].shape == (n, plasma.N_TABLE) and b['voigt_cdf'].shape == (n, plasma.N_TABLE)
    gamma = voigt.natural_gamma(sparam['linewidth'], sparam['wavelength'])
    temp = b['props']['temperature'][b['counts'] > 0]
    assert len(np.unique(temp)) > 5                                   # a different table per bundle
    for row in (0, n // 2, n - 1):
        sigma = voigt.doppler_sigma(temp[row], sparam['mass_number'], sparam['wavelength'])
        assert np.isclose(b['table']['wave_sigma'][row], sigma, rtol=1e-14)
        x, cdf = voigt.cdf_table(gamma, float(sigma))
        assert np.array_equal(b['voigt_x'][row], x) and np.array_equal(b['voigt_cdf'][row], cdf)
    # T == 0 with a natural linewidth is given 1 eV (_XicsrtSourceGeneric.py:333-339); without one sigma = 0
    assert np.isclose(plasma.bundle_sigma(sparam, [0.0])[0], voigt.doppler_sigma(1.0, sparam['mass_number'], sparam['wavelength']))
    assert plasma.bundle_sigma(dict(sparam, linewidth=0.0), [0.0])[0] == 0.0
    assert plasma.line_model(dict(sparam, wavelength_dist='monochrome')) == 'const'
    # cone parameter per distribution (include/xrt.h: XrtBundle.cos_spread)
    s = np.array([0.05, 0.1])
    for name, fn in (('isotropic', np.cos), ('flat', np.tan), ('flat_xy', np.tan), ('isotropic_xy', np.sin)):
        assert np.array_equal(plasma.cone_parameter(name, s), fn(s))
    with pytest.raises(NotImplementedError):
        plasma.cone_kind(dict(sparam, angular_dist='gaussian'))
    for name in ('plasma_flat', 'plasma_flat_xy', 'plasma_isotropic_xy'):
        _, _, sp, sf, _ = prepared(name)
        bb = plasma.build_bundles(sp, sf, StreamAdapter(LegacyStream(1)))
        kind = plasma.cone_kind(sp)
        assert np.array_equal(bb['table']['cos_spread'], plasma.cone_parameter(kind, bb['props']['spread'][bb['counts'] > 0]))


def test_poisson_counts_and_limits():
    _, sname, sparam, sfilters, optics = prepared('plasma_cubic_poisson')
    from xicsrt_b200._driver import HostRandom
    totals = []
    for it in range(20):
        b = plasma.build_bundles(sparam, sfilters, HostRandom(5, 1 + it))
        totals.append(b['n_rays'])
        assert b['end'][-1] == b['counts'].sum() and np.all(np.diff(b['end'].astype(np.int64)) > 0)
    props = plasma.bundle_properties(sparam, sfilters, HostRandom(5, 1))
    expect = plasma.bundle_intensity(sparam, props).sum()
    assert abs(np.mean(totals) - expect) < 5 * np.sqrt(expect / 20)
    assert len(set(totals)) > 1                                   # a fresh draw per iteration
    again = plasma.build_bundles(sparam, sfilters, HostRandom(5, 3))
    assert again['n_rays'] == totals[2]                           # and reproducible per (seed, iteration)

    p = dict(sparam, max_rays=10)
    with pytest.raises(ValueError, match='too many rays'):
        plasma.build_bundles(p, sfilters, HostRandom(5, 1))
    p = dict(sparam, use_poisson=False, emissivity=1e6)
    with pytest.raises(ValueError, match='less than one'):
        plasma.build_bundles(p, sfilters, HostRandom(5, 1))


def test_integrated_test_00_ray_budget():
    """testing/integrated_test_00.ipynb: generated rays ~ emissivity * volume * time for spread = pi."""
    src = {'class_name': 'XicsrtPlasmaCubic', 'xsize': 0.01, 'ysize': 0.01, 'zsize': 0.01, 'target': [0, 0, 1.0],
           'emissivity': 1e12, 'time_resolution': 1.0, 'spread': np.pi, 'use_poisson': True, 'bundle_count': 500,
           'bundle_volume': 1e-9, 'max_rays': int(1e8)}
    cfg = xconfig.get_config(xconfig.to_numpy({'general': {}, 'sources': {'s': src},
                                               'optics': {'d': {'class_name': 'XicsrtOpticDetector', 'origin': [0, 0, 1.0]}}}))
    _, sname, sparam, sfilters, optics = xscene.prepare(cfg)
    from xicsrt_b200._driver import HostRandom
    b = plasma.build_bundles(sparam, sfilters, HostRandom(1, 1))
    expected = 1e12 * 0.01**3
    assert abs(b['n_rays'] - expected) < 5 * np.sqrt(expected)


# ---------------------------------------------------------------------------

@pytest.fixture(scope='module')
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


class FixedUniforms:
    """Host-random interface that hands out prepared uniforms (the ones injected into the device kernel)."""

    def __init__(self, u):
        self.u, self.k = u, 0

    def uniform(self, lo, hi, n):
        out = lo + (hi - lo) * self.u[self.k]
        self.k += 1
        return out


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['plasma_cubic', 'plasma_cubic_poisson', 'plasma_toroidal', 'plasma_datafile'])
def test_device_bundle_table_matches_host_restatement(torch, name):
    """xrt_bundles_generate with injected centre uniforms against the numpy restatement (pinned to the oracle above)."""
    cfg = scenes.get(name)
    cfg['sources']['source']['bundle_count'] = 5000
    _, sname, sparam, sfilters, optics = xscene.prepare(xconfig.get_config(xconfig.to_numpy(cfg)))
    dev = torch.device('cuda', 0)
    n = int(sparam['bundle_count'])
    u = np.random.default_rng(8).random((3, n))
    ref = plasma.bundle_properties(sparam, sfilters, FixedUniforms(u))
    inten = plasma.bundle_intensity(sparam, ref)

    db = plasma.DeviceBundles(torch, dev, dict(sparam, use_poisson=True, max_rays=None), sfilters, L.load())
    total = db.generate(5, 1 << 32, inject_u=torch.from_numpy(u).to(dev))
    table = db.table.cpu().numpy()
    got_int = db.intensity.cpu().numpy()
    counts = db.counts.cpu().numpy()
    m = ref['mask']
    assert np.array_equal(got_int >= 0, m)
    np.testing.assert_allclose(table[:, 0:3], ref['origin'], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(table[:, 3], np.cos(ref['spread']), rtol=1e-13)
    np.testing.assert_allclose(got_int[m], inten[m], rtol=1e-12)
    np.testing.assert_allclose(table[m][:, 5:8], ref['velocity'][m] / voigt.C_LIGHT, rtol=1e-13, atol=1e-30)
    thermal = ref['temperature'] > 0
    sigma = np.where(thermal, voigt.doppler_sigma(np.abs(ref['temperature']), sparam['mass_number'], sparam['wavelength']), 0.0)
    np.testing.assert_allclose(table[m][:, 4], sigma[m], rtol=1e-13)
    # Poisson counts: zero where filtered, mean and variance of (k - lam) / sqrt(lam) as expected
    assert not counts[~m].any() and total == counts.sum()
    lam = inten[m]
    ok = lam > 0
    z = (counts[m][ok] - lam[ok]) / np.sqrt(lam[ok])
    assert abs(z.mean()) < 5 / np.sqrt(len(z)) and abs(z.var() - 1) < 0.15
    # without Poisson the count is the truncated intensity
    if np.all(lam >= 1):
        db2 = plasma.DeviceBundles(torch, dev, dict(sparam, use_poisson=False, max_rays=None), sfilters, L.load())
        db2.generate(5, 1 << 32, inject_u=torch.from_numpy(u).to(dev))
        assert np.array_equal(db2.counts.cpu().numpy()[m], lam.astype(np.int64))


@pytest.mark.gpu
def test_device_poisson_sampler_small_and_large_means(torch):
    """Both branches of the sampler (inversion below 10, transformed rejection above) against scipy's pmf."""
    from scipy import stats
    cfg = scenes.get('plasma_cubic_poisson')
    dev = torch.device('cuda', 0)
    for emis in (5.3e13, 1.4e14, 5.3e14, 5.3e16):       # means of about 3, 8, 30 and 3000 rays per bundle
        cfg['sources']['source'].update({'bundle_count': 200000, 'emissivity': emis, 'spread_radius': None,
                                         'spread': float(np.radians(5.0))})
        _, sname, sparam, sfilters, optics = xscene.prepare(xconfig.get_config(xconfig.to_numpy(cfg)))
        db = plasma.DeviceBundles(torch, dev, dict(sparam, max_rays=None), sfilters, L.load())
        db.generate(11, 1 << 32)
        lam = float(db.intensity[0].cpu())
        k = db.counts.cpu().numpy()
        assert abs(k.mean() - lam) < 5 * np.sqrt(lam / len(k))
        lo, hi = int(max(0, lam - 6 * np.sqrt(lam))), int(lam + 6 * np.sqrt(lam)) + 2
        obs = np.bincount(np.clip(k, lo, hi) - lo, minlength=hi - lo + 1).astype(float)
        exp = stats.poisson.pmf(np.arange(lo, hi + 1), lam) * len(k)
        exp[0] += stats.poisson.cdf(lo - 1, lam) * len(k)
        exp[-1] += stats.poisson.sf(hi, lam) * len(k)
        keep = exp > 20
        chi2 = np.sum((obs[keep] - exp[keep])**2 / exp[keep])
        p = stats.chi2.sf(chi2, keep.sum() - 1)
        assert p > 1e-4, (lam, chi2, keep.sum(), p)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['plasma_cubic_poisson', 'plasma_toroidal', 'plasma_datafile'])
def test_device_rays_follow_their_bundles(torch, name):
    from xicsrt_b200 import _driver
    cfg = scenes.get(name)
    cfg['sources']['source']['bundle_count'] = 400
    cfg['sources']['source']['time_resolution'] *= 300
    tracer = _driver.Tracer(xconfig.get_config(xconfig.to_numpy(cfg)), seed=17)
    n = tracer.n_rays
    assert n > 2000
    ids = torch.arange(n, dtype=torch.int64, device=tracer.device)
    rays, mask = tracer.history(0, ids)
    r = rays.cpu().numpy()[0]
    org, dirs, lam = r[0:3].T, r[3:6].T, r[6]
    end = tracer.bundles.end.cpu().numpy()
    assert end[-1] == n
    which = np.searchsorted(end, np.arange(n), side='right')
    tab = tracer.bundles.table.cpu().numpy()[which]
    t_origin, t_cos, t_sigma, t_vel = tab[:, 0:3], tab[:, 3], tab[:, 4], tab[:, 5:8]
    half = tracer.source_param['voxel_size'] / 2
    assert np.all(np.abs(org - t_origin) <= half * (1 + 1e-12))
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-12)
    axis = tracer.source_param['target'] - org
    axis /= np.linalg.norm(axis, axis=1)[:, None]
    cosang = np.einsum('ij,ij->i', axis, dirs)
    assert np.all(cosang >= t_cos - 1e-12)
    # wavelength: Doppler-shifted normal about lambda0 with the bundle's sigma
    lam0 = tracer.source_param['wavelength']
    shift = 1 - np.einsum('ij,ij->i', t_vel, dirs)
    z = (lam / shift - lam0) / np.where(t_sigma > 0, t_sigma, 1.0)
    z = z[t_sigma > 0]
    assert abs(z.mean()) < 5 / np.sqrt(len(z)) and abs(z.std() - 1) < 5 / np.sqrt(2 * len(z))
    # a new table every iteration, reproducible per (seed, iteration)
    first = tracer.bundles.table.clone()
    tracer.begin_iteration(1)
    assert not torch.equal(first, tracer.bundles.table)
    tracer._new_bundles(0)
    assert torch.equal(first, tracer.bundles.table) and tracer.n_rays == n
    tracer.close()


@pytest.mark.gpu
def test_plasma_end_to_end_statistics(torch):
    """XicsrtPlasmaCubic -> spherical crystal -> detector: counts per element vs the oracle."""
    import xicsrt_b200
    cfg = scenes.get('plasma_cubic_poisson')
    cfg['sources']['source'].update({'bundle_count': 300, 'emissivity': 6e11})
    cfg['general']['keep_history'] = False
    cfg['general']['number_of_iter'] = 3
    ref = oracle.raytrace(copy.deepcopy(cfg))
    cfg['general']['random_seed'] = 99
    got = xicsrt_b200.raytrace(cfg)
    n_ref, n_got = ref['total']['meta']['source']['num_out'], got['total']['meta']['source']['num_out']
    assert abs(n_ref - n_got) < 6 * np.sqrt(n_ref + n_got)
    for elem in ('crystal', 'detector'):
        p_ref = ref['total']['meta'][elem]['num_out'] / n_ref
        p_got = got['total']['meta'][elem]['num_out'] / n_got
        p = 0.5 * (p_ref + p_got)
        z = (p_got - p_ref) / np.sqrt(p * (1 - p) * (1 / n_ref + 1 / n_got))
        assert abs(z) < 4.5, (elem, p_ref, p_got, z)
    assert got['total']['image']['detector'].sum() == got['total']['meta']['detector']['num_out']


# ---------------------------------------------------------------------------
# per-ray parity of the plasma sources: the oracle's bundle centres, ray counts and per-ray draws injected

class LoggedRandom:
    """Host-random interface that replays what the oracle's stream recorded for the bundle table."""

    def __init__(self, stream):
        self.centres = [v for s, _, v, _ in stream.log if s.startswith('plasma.center.')]
        self.counts = np.array([int(v[0]) for s, _, v, _ in stream.log if s == 'plasma.count'], dtype=np.int64)
        self.k = 0

    def uniform(self, lo, hi, n):
        u = self.centres[self.k]
        self.k += 1
        return lo + (hi - lo) * u

    def poisson_array(self, lam):
        assert len(lam) == len(self.counts)
        return self.counts


def _concat(stream, site):
    parts = [v for s, _, v, _ in stream.log if s == site]
    return np.concatenate(parts) if parts else None


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['plasma_cubic', 'plasma_cubic_poisson', 'plasma_toroidal', 'plasma_datafile',
                                  'plasma_voigt', 'plasma_flat', 'plasma_flat_xy'])
def test_injected_plasma_rays_match_oracle_and_reference(torch, golden_dir, name):
    """
    Every ray of a plasma source -- voxel origin, focused cone of its bundle (isotropic / flat /
    flat_xy), thermal or per-bundle Voigt wavelength, Doppler shift -- within 1e-9 of the oracle
    and of the unmodified reference's stored rays, from the same draws.
    """
    import harness
    from test_gpu_source import run_source
    cfg = scenes.get(name)
    single, stream, oscene = oracle.trace_recorded(cfg)
    ref = single['history']['source']
    n = len(ref['mask'])
    _, sname, sparam, sfilters, optics = prepared(name)
    bundles = plasma.build_bundles(sparam, sfilters, LoggedRandom(stream))
    assert bundles['n_rays'] == n
    desc, layout, keep = xscene.flatten(sname, sparam, sfilters, optics, bundles=bundles)
    scene = xscene.DeviceScene(desc, layout)
    origin = np.stack([_concat(stream, f'src.origin.{i}') for i in range(3)])
    cone = np.stack([_concat(stream, 'src.cone.0'), _concat(stream, 'src.cone.1')])
    wave = _concat(stream, 'src.wave')
    assert origin.shape == (3, n) and cone.shape == (2, n) and wave.shape == (n,)
    got = run_source(torch, scene, n, origin, cone, wave)
    scene.close()
    harness.assert_rays_close(got, ref, f'{name}/source', 1e-9)
    gold = np.load(f'{golden_dir}/{name}.npz')
    gref = {k: gold[f'iter/source/{k}'] for k in ('origin', 'direction', 'wavelength', 'mask')}
    harness.assert_rays_close(got, gref, f'{name}/source (golden)', 1e-9)


@pytest.mark.gpu
def test_device_voigt_tables_match_scipy(torch):
    """xrt_bundle_voigt_tables (device Faddeeva function) against the scipy-built tables of the host restatement."""
    dev = torch.device('cuda', 0)
    lam0, mass = 3.9492, 39.948
    temps = np.array([0.3, 1.0, 25.0, 400.0, 1550.0, 9000.0, 1e5, 5.0])
    counts = np.array([3, 1, 2, 9, 1, 1, 4, 0], dtype=np.int64)
    for linewidth in (1e11, 1e13, 1e14, 3e15):
        gamma = float(voigt.natural_gamma(linewidth, lam0))
        sigma = voigt.doppler_sigma(temps, mass, lam0)
        table = np.zeros((len(temps), 8))
        table[:, 4] = sigma
        t_table = torch.from_numpy(table).to(dev)
        t_counts = torch.from_numpy(counts).to(dev)
        x = torch.full((len(temps), plasma.N_TABLE), -1.0, dtype=torch.float64, device=dev)
        cdf = torch.full((len(temps), plasma.N_TABLE), -1.0, dtype=torch.float64, device=dev)
        L.check(L.load().xrt_bundle_voigt_tables(t_table.data_ptr(), t_counts.data_ptr(), len(temps), gamma,
                                                 plasma.N_TABLE, x.data_ptr(), cdf.data_ptr(), None))
        torch.cuda.synchronize()
        x, cdf = x.cpu().numpy(), cdf.cpu().numpy()
        for b in range(len(temps)):
            if counts[b] == 0:
                assert np.all(x[b] == -1.0) and np.all(cdf[b] == -1.0)       # rows of empty bundles are not built
                continue
            rx, rcdf = voigt.cdf_table(gamma, float(sigma[b]), gridsize=plasma.N_TABLE)
            np.testing.assert_allclose(x[b], rx, rtol=1e-12, atol=1e-15 * np.max(np.abs(rx)))   # the centre edge is ~0
            np.testing.assert_allclose(cdf[b], rcdf, rtol=1e-11)
            np.testing.assert_allclose(np.diff(cdf[b]), np.diff(rcdf), rtol=1e-10, atol=4e-16)    # bin weights (diff of O(1) sums)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['plasma_voigt', 'plasma_isotropic_xy', 'plasma_flat'])
def test_plasma_linewidth_and_cones_end_to_end(torch, name):
    """raytrace(config) on the device path (device-built bundle and Voigt tables, Philox draws) vs the oracle."""
    from scipy import stats
    import xicsrt_b200
    cfg = scenes.get(name)
    # many bundles with a few rays each: the fraction of a bundle's rays that meets the Bragg condition depends
    # strongly on where the bundle sits relative to the Rowland circle, and the two sides draw different bundles
    cfg['sources']['source'].update({'bundle_count': 4000, 'use_poisson': True, 'max_rays': int(1e8)})
    cfg['sources']['source']['time_resolution'] *= 10
    cfg['general'].update({'keep_history': True, 'number_of_iter': 2, 'max_lost': 0})
    ref = oracle.raytrace(copy.deepcopy(cfg))
    cfg['general']['random_seed'] = 1234
    got = xicsrt_b200.raytrace(cfg)
    n_ref, n_got = ref['total']['meta']['source']['num_out'], got['total']['meta']['source']['num_out']
    assert n_ref > 20000 and abs(n_ref - n_got) < 6 * np.sqrt(n_ref + n_got)
    for elem in ('crystal', 'detector'):
        p_ref = ref['total']['meta'][elem]['num_out'] / n_ref
        p_got = got['total']['meta'][elem]['num_out'] / n_got
        p = 0.5 * (p_ref + p_got)
        z = (p_got - p_ref) / np.sqrt(p * (1 - p) * (1 / n_ref + 1 / n_got))
        assert abs(z) < 4.5, (elem, p_ref, p_got, z)
    # found rays: wavelength and direction distributions at the source
    a, b = got['found']['history']['source'], ref['found']['history']['source']
    assert len(a['wavelength']) > 200
    assert stats.ks_2samp(a['wavelength'], b['wavelength']).pvalue > 1e-4
    for ax in range(3):
        assert stats.ks_2samp(a['direction'][:, ax], b['direction'][:, ax]).pvalue > 1e-4
