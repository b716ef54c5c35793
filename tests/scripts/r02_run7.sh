set -x
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tests/scripts/quick_rate.py config4 config3 > gpurun_out/r02/run7_default.jsonl 2>&1; cut -c1-130 gpurun_out/r02/run7_default.jsonl
python - <<'P'
import time, sys
sys.path.insert(0, '.')
import bench, xicsrt_b200, torch
for n in (1e6, 1e8):
    for k in range(4):
        cfg = bench.workload_config('config2', int(n), seed=k, history=True)
        cfg['general']['history_max_lost'] = 10000
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = xicsrt_b200.raytrace(cfg)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print('history e2e', n, 'found', len(res['found']['history']['detector']['mask']), 'lost', len(res['lost']['history']['detector']['mask']), 'ms %.1f' % (dt * 1e3), 'rays/s %.3e' % (n / dt), flush=True)
P
python -c "import __graft_entry__ as g; g.smoke()"
