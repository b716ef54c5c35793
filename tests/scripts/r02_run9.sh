set -x
mkdir -p gpurun_out/r02
python tests/scripts/quick_rate.py config2 config5 config3 step > gpurun_out/r02/run9_default.jsonl 2>&1
XRT_LIB_PATH=$PWD/build/var/libxrt_p7.so python tests/scripts/quick_rate.py config2 config5 config3 step > gpurun_out/r02/run9_p7.jsonl 2>&1
cat gpurun_out/r02/run9_*.jsonl | cut -c1-120
