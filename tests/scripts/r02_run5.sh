set -x
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q -k "mesh or save_images or images_are_saved or history or api_output" 2>&1 | tail -5
python tests/scripts/quick_rate.py config4 > gpurun_out/r02/run5_c4.jsonl 2>&1; cut -c1-150 gpurun_out/r02/run5_c4.jsonl
python bench.py --steps 5 --warmup 3 --no-cpu --quick > gpurun_out/r02/bench_run5_quick.json 2> gpurun_out/r02/bench_run5_quick.err; tail -3 gpurun_out/r02/bench_run5_quick.err
python - <<'P'
import time, sys
sys.path.insert(0, '.')
import bench, xicsrt_b200, torch
for n in (1e7, 1e8):
    for k in range(3):
        cfg = bench.workload_config('config2', int(n), seed=k, history=True)
        cfg['general']['history_max_lost'] = 10000
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = xicsrt_b200.raytrace(cfg)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print('history e2e', n, 'found', len(res['found']['history']['detector']['mask']), 'ms %.1f' % (dt * 1e3), 'rays/s %.3e' % (n / dt), flush=True)
P
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh gpurun_out/r02/run5_c4 k_trace k_traceILj9ELi0ELj0ELb0 1e8 $Q config4
