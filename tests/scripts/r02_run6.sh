set -x
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tests/scripts/quick_rate.py config4 config3 config2 > gpurun_out/r02/run6_default.jsonl 2>&1
XRT_LIB_PATH=$PWD/build/var/libxrt_mesh3.so python tests/scripts/quick_rate.py config4 > gpurun_out/r02/run6_mesh3.jsonl 2>&1
for v in ms8 ms24; do XRT_LIB_PATH=$PWD/build/var/libxrt_$v.so python tests/scripts/quick_rate.py config3 > gpurun_out/r02/run6_$v.jsonl 2>&1; done
cat gpurun_out/r02/run6_*.jsonl | cut -c1-130
