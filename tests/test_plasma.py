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


@pytest.mark.parametrize('name', ['plasma_cubic', 'plasma_cubic_poisson', 'plasma_toroidal', 'plasma_datafile'])
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
    from xicsrt_b200 import plasma as xp
    b = xp.build_bundles(tracer.source_param, tracer.source_filters, _driver.HostRandom(17, 1))
    assert b['n_rays'] == n
    end = b['end'].astype(np.int64)
    which = np.searchsorted(end, np.arange(n), side='right')
    tab = b['table'][which]
    half = tracer.source_param['voxel_size'] / 2
    assert np.all(np.abs(org - tab['origin']) <= half * (1 + 1e-12))
    assert np.allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-12)
    axis = tracer.source_param['target'] - org
    axis /= np.linalg.norm(axis, axis=1)[:, None]
    cosang = np.einsum('ij,ij->i', axis, dirs)
    assert np.all(cosang >= tab['cos_spread'] - 1e-12)
    # wavelength: Doppler-shifted normal about lambda0 with the bundle's sigma
    lam0 = tracer.source_param['wavelength']
    shift = 1 - np.einsum('ij,ij->i', tab['velocity_c'], dirs)
    z = (lam / shift - lam0) / np.where(tab['wave_sigma'] > 0, tab['wave_sigma'], 1.0)
    z = z[tab['wave_sigma'] > 0]
    assert abs(z.mean()) < 5 / np.sqrt(len(z)) and abs(z.std() - 1) < 5 / np.sqrt(2 * len(z))
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
