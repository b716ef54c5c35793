"""Phase timing of the multi-rank history path (development): torchrun --nproc-per-node 2 tests/scripts/dist_history_timing.py"""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rank = int(os.environ['RANK']); torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
import bench
from xicsrt_b200 import _driver, config as xconfig
n = 100_000_000 * dist.get_world_size()
for k in range(5):
    cfg = xconfig.get_config(xconfig.to_numpy(bench.workload_config('config2', n, seed=k, history=True)))
    dist.barrier(); torch.cuda.synchronize(); t = [time.perf_counter()]
    tr = _driver.Tracer(cfg, k, rank=rank, world=dist.get_world_size()); torch.cuda.synchronize(); t.append(time.perf_counter())
    found, lost = tr.select_ids(0, 5000); torch.cuda.synchronize(); t.append(time.perf_counter())
    tr.allreduce(); meta, image = tr.counts_and_images(True); t.append(time.perf_counter())
    ids = torch.cat([found, lost]); rays, mask = tr.history(0, ids, rows=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    rays, mask, nf = _driver.gather_rows(torch, rays, mask, int(found.numel())); torch.cuda.synchronize(); t.append(time.perf_counter())
    h = _driver.to_host(torch, rays, mask); t.append(time.perf_counter())
    tr.close()
    if rank == 0:
        names = ['tracer', 'select', 'reduce+meta', 'replay', 'gather', 'to_host']
        print(k, ' '.join(f'{a} {1e3 * (t[i + 1] - t[i]):.1f}' for i, a in enumerate(names)), 'total %.1f ms' % (1e3 * (t[-1] - t[0])), flush=True)
dist.barrier(); dist.destroy_process_group()
