# -*- coding: utf-8 -*-
"""
The multi-GPU host logic on CPU: two ranks over gloo.  Each rank plays one GPU's
share of an iteration (its own oracle run stands in for the kernel), then the product's
own collective code runs: the packed counter/image all-reduce, the variable-length
history gather onto rank 0 and the seed agreement for ``random_seed=None``.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import oracle
        from oracle import scenes
        from xicsrt_b200 import _driver

        cfg = scenes.get('sphere')
        cfg['general']['random_seed'] = 100 + rank          # this rank's share of the rays
        cfg['sources']['source']['intensity'] = 3000 + 500 * rank
        res = oracle.raytrace(cfg)
        names = list(res['total']['meta'].keys())

        # packed [counts | images] exactly as Tracer lays it out
        counts = [res['total']['meta'][n]['num_out'] for n in names]
        images = [res['total']['image'][n].ravel() for n in names[1:]]
        packed = torch.from_numpy(np.concatenate([np.array(counts, dtype=np.int64)] +
                                                 [im.astype(np.int64) for im in images]))
        mine = packed.clone()
        _driver.allreduce_packed(packed)

        # this rank's found + lost histories in the rows layout the replay kernel writes (XRT_HIST_ROWS)
        n_found = len(res['found']['history'][names[-1]]['mask'])
        n_lost = len(res['lost']['history'][names[-1]]['mask'])
        n = n_found + n_lost
        rays = torch.zeros((len(names), 7, n), dtype=torch.float64)
        mask = torch.zeros((len(names), n), dtype=torch.uint8)
        for e, name in enumerate(names):
            for kind, lo, hi in (('found', 0, n_found), ('lost', n_found, n)):
                v = _driver.rows_views(rays, mask, e, lo, hi)
                h = res[kind]['history'][name]
                v['origin'].copy_(torch.from_numpy(h['origin']))
                v['direction'].copy_(torch.from_numpy(h['direction']))
                v['wavelength'].copy_(torch.from_numpy(h['wavelength']))
                v['mask'].copy_(torch.from_numpy(h['mask'].astype(np.uint8)))
        g_rays, g_mask, g_found = _driver.gather_rows(torch, rays, mask, n_found)
        h_rays, h_mask = _driver.to_host(torch, g_rays, g_mask)
        found, lost = _driver._row_dicts(names, h_rays, h_mask, g_found)
        out = {'found': {'history': found}, 'lost': {'history': lost}}

        seed = _driver._resolve_seed(None, world)
        assert _driver._dist_info() == (rank, world)
        begin, count = _driver.shard_range(10**9 + 7, rank, world)
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), mine=mine.numpy(), packed=packed.numpy(), seed=seed,
                 begin=begin, count=count,
                 found_origin=out['found']['history']['detector']['origin'],
                 own_found=res['found']['history']['detector']['origin'],
                 lost_mask=out['lost']['history']['crystal']['mask'])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reduce_and_gather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / 'rank0.npz')
    r1 = np.load(tmp_path / 'rank1.npz')
    # the reduced buffer is the same on both ranks and equals the sum of the two shares
    assert np.array_equal(r0['packed'], r1['packed'])
    assert np.array_equal(r0['packed'], r0['mine'] + r1['mine'])
    assert r0['packed'][0] == 3000 + 3500
    # rank 0 holds the concatenation of both ranks' found rays in rank order
    assert np.array_equal(r0['found_origin'], np.concatenate([r0['own_found'], r1['own_found']]))
    assert np.array_equal(r1['found_origin'], r1['own_found'])
    assert len(r0['lost_mask']) > len(r1['lost_mask'])
    # one seed for all ranks, contiguous ray-id ranges
    assert int(r0['seed']) == int(r1['seed'])
    assert int(r0['begin']) == 0 and int(r0['begin']) + int(r0['count']) == int(r1['begin'])
    assert int(r1['begin']) + int(r1['count']) == 10**9 + 7
