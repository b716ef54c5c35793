"""Host-side profile of the public API call (development): python tests/scripts/e2e_profile.py config4 [rays]"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import xicsrt_b200
name = sys.argv[1] if len(sys.argv) > 1 else 'config2'
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else (1_000_000_000 if name in ('config2', 'config5') else 100_000_000)
cfg = bench.workload_config(name, n)
cfg['general']['keep_history'] = False
for _ in range(3):
    xicsrt_b200.raytrace(bench.workload_config(name, n) | {})
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    xicsrt_b200.raytrace(bench.workload_config(name, n))
torch.cuda.synchronize()
print(name, 'ms per call', (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    xicsrt_b200.raytrace(bench.workload_config(name, n))
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
