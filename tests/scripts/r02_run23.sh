set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config3 config4 > $O/run23_default.jsonl 2>&1; cut -c1-120 $O/run23_default.jsonl
XRT_LIB_PATH=$PWD/build/var/libxrt_mb3.so python tests/scripts/quick_rate.py config3 config4 > $O/run23_mb3.jsonl 2>&1; cut -c1-120 $O/run23_mb3.jsonl
