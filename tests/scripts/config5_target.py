# -*- coding: utf-8 -*-
"""
BASELINE.json configs[4] / SURVEY.md section 8d config 5 at full size, under torchrun on N GPUs:
XicsrtPlasmaCubic (1e5 bundles, Poisson counts) -> spherical Bragg crystal -> detector, 1e10 rays
in one iteration sharded by ray id over the ranks, history off, images + counters reduced over
NCCL; then the history of a 1e6-ray strided subsample of the ids (every element) replayed on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        tests/scripts/config5_target.py [--rays 1e10] [--steps 3]

Prints one JSON line (rank 0): device-timed rays/s for the iteration (bundle table + fused kernel
+ all-reduce, max over ranks), the history pass and consistency checks.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rays', type=float, default=1e10)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--history', type=float, default=1e6)
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    import bench
    from xicsrt_b200 import _driver, config as xconfig

    cfg = bench.workload_config('config5', int(args.rays), seed=0, history=False)
    full = xconfig.get_config(xconfig.to_numpy(cfg))
    tracer = _driver.Tracer(full, seed=0, rank=rank, world=world)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(it):
        tracer.begin_iteration(it)
        tracer.trace(it, keep_images=True)
        tracer.allreduce()

    step(0)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launched = 0
    e0.record()
    for it in range(1, 1 + args.steps):
        step(it)
        launched += tracer.n_rays
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    meta, image = tracer.counts_and_images(True)

    out = None
    if rank == 0:
        # a strided subsample of the whole id range (consecutive ids would all come from the first few bundles)
        n_hist = int(min(args.history, tracer.n_rays))
        ids = torch.arange(n_hist, dtype=torch.int64, device=dev) * (tracer.n_rays // n_hist)
        rays, mask = tracer.history(args.steps, ids)          # warm-up
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        rays, mask = tracer.history(args.steps, ids, out=(rays, mask))
        h1.record()
        torch.cuda.synchronize()
        host_rays, host_mask = rays.cpu().numpy(), mask.cpu().numpy().astype(bool)
        n_elem = host_rays.shape[0]
        found = host_mask[-1]
        # consistency: history masks are cumulative, found rays sit on the detector plane
        assert np.all(~host_mask[1:] | host_mask[:-1])
        det = full['optics']['detector']
        z = (host_rays[-1][0:3][:, found].T - np.asarray(det['origin'])) @ np.asarray(det['zaxis'])
        assert np.all(np.abs(z) < 1e-9)
        assert image['detector'].sum() == meta['detector']
        out = {'workload': 'config5: XicsrtPlasmaCubic(1e5 bundles, Poisson) -> XicsrtOpticSphericalCrystal -> XicsrtOpticDetector',
               'n_gpus': world, 'steps': args.steps, 'rays_per_step': launched // args.steps,
               'rays_per_sec': launched / float(t[0]), 'ms_per_step': float(t[0]) * 1e3 / args.steps,
               'detected_last_step': meta['detector'], 'crystal_last_step': meta['crystal'],
               'history': {'rays': n_hist, 'elements': n_elem, 'found': int(found.sum()),
                           'ms': h0.elapsed_time(h1), 'GBps': 57.0 * n_hist * n_elem / (h0.elapsed_time(h1) * 1e-3) / 1e9},
               'timing': 'CUDA events on the launch stream, max over ranks; bundle table, fused kernel and all-reduce inside'}
        print(json.dumps(out), flush=True)
    tracer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
