set -x
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r02/base_config2.json 2> gpurun_out/r02/base_config2.err
python bench.py --workload config3 --rays 1e8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02/base_config3.json 2>&1
python bench.py --workload config4 --rays 1e8 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02/base_config4.json 2>&1
python bench.py --workload config5 --rays 1e9 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02/base_config5.json 2>&1
NCU="ncu --set full --import-source on --clock-control none"
$NCU -k regex:k_trace --launch-skip 1 -c 1 -f -o gpurun_out/r02/base_trace_c2 python bench.py --rays 1e8 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02/ncu_c2.log 2>&1
$NCU -k regex:k_record --launch-skip 1 -c 1 -f -o gpurun_out/r02/base_record python bench.py --rays 1e8 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02/ncu_rec.log 2>&1
$NCU -k regex:k_trace --launch-skip 1 -c 1 -f -o gpurun_out/r02/base_trace_c3 python bench.py --workload config3 --rays 1e7 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02/ncu_c3.log 2>&1
$NCU -k regex:k_trace --launch-skip 1 -c 1 -f -o gpurun_out/r02/base_trace_c4 python bench.py --workload config4 --rays 1e7 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02/ncu_c4.log 2>&1
$NCU -k regex:k_trace --launch-skip 1 -c 1 -f -o gpurun_out/r02/base_trace_c5 python bench.py --workload config5 --rays 1e8 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02/ncu_c5.log 2>&1
ls -la gpurun_out/r02
