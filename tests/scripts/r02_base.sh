set -x
mkdir -p gpurun_out/r02
python bench.py --steps 10 --warmup 3 > gpurun_out/r02/base_config2.json 2> gpurun_out/r02/base_config2.err
python bench.py --workload config3 --rays 1e8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02/base_config3.json 2>&1
python bench.py --workload config4 --rays 1e8 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02/base_config4.json 2>&1
python bench.py --workload config5 --rays 1e9 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02/base_config5.json 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu"
profiles/capture.sh gpurun_out/r02/base_trace_c2 k_trace k_traceILj0ELi0ELj63 1e8 $B --rays 1e8
profiles/capture.sh gpurun_out/r02/base_record k_record k_recordILj0ELi0ELj63 16777216 $B --rays 1e8
profiles/capture.sh gpurun_out/r02/base_trace_c3 k_trace k_traceILj32ELi0ELj0 1e7 $B --workload config3 --rays 1e7
profiles/capture.sh gpurun_out/r02/base_trace_c4 k_trace k_traceILj9ELi0ELj0 1e7 $B --workload config4 --rays 1e7
profiles/capture.sh gpurun_out/r02/base_trace_c5 k_trace k_traceILj128ELi0ELj0 1e8 $B --workload config5 --rays 1e8
python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
ls -la gpurun_out/r02
