# -*- coding: utf-8 -*-
"""
Run under torchrun on N GPUs: the sharded run must give exactly the single-GPU result
(counters, images, found histories), because Philox is counted by global ray id.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        tests/scripts/dist_history_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    import xicsrt_b200
    from xicsrt_b200 import _driver, config as xconfig
    from oracle import scenes

    for name in ('sphere', 'plasma_cubic_poisson', 'mesh_torus'):
        cfg = scenes.get(name)
        if name.startswith('plasma'):
            cfg['sources']['source']['time_resolution'] *= 2000
        else:
            cfg['sources']['source']['intensity'] = 1000003
        cfg['general'].update({'number_of_iter': 2, 'history_max_lost': 500, 'random_seed': 77})
        res = xicsrt_b200.raytrace(cfg)                      # sharded over the ranks

        if rank == 0:
            # the same run on this GPU alone
            full = xconfig.get_config(xconfig.to_numpy(scenes.get(name)))
            full['sources']['source'].update(cfg['sources']['source'])
            full['general'].update(cfg['general'])
            tr = _driver.Tracer(full, 77, rank=0, world=1)
            parts = [_driver.run_iteration(tr, it, max_lost=250) for it in range(2)]
            tr.close()
            one = _driver.combine_raytrace(parts)
            for elem in one['total']['meta']:
                assert res['total']['meta'][elem]['num_out'] == one['total']['meta'][elem]['num_out'], (name, elem)
                if one['total']['image'].get(elem) is not None:
                    assert np.array_equal(res['total']['image'][elem], one['total']['image'][elem]), (name, elem)
                for key in ('origin', 'direction', 'wavelength', 'mask'):
                    a, b = res['found']['history'][elem][key], one['found']['history'][elem][key]
                    assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), (name, elem, key)
            n_lost = len(res['lost']['history'][elem]['mask'])
            assert 0 < n_lost <= 500
            print(f'{name}: {world} ranks == 1 rank  (source {res["total"]["meta"]["source"]["num_out"]}, '
                  f'found {len(res["found"]["history"][elem]["mask"])}, lost kept {n_lost})', flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('DIST_HISTORY_OK', flush=True)


if __name__ == '__main__':
    main()
