set -x
mkdir -p gpurun_out/r02
python tests/scripts/quick_rate.py config3 > gpurun_out/r02/run4_c3.jsonl 2>&1; cut -c1-120 gpurun_out/r02/run4_c3.jsonl
python -m pytest tests/test_gpu_scale.py tests/test_gpu_statistics.py -m gpu -x -q -k "mosaic" 2>&1 | tail -4
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02/bench_run4.json 2> gpurun_out/r02/bench_run4.err
tail -5 gpurun_out/r02/bench_run4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_run4_ref.json 2>&1
