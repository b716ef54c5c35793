set -x
O=gpurun_out/r02; mkdir -p $O
XRT_LIB_PATH=$PWD/build/var/libxrt_rb4.so python tests/scripts/quick_rate.py config4 > $O/run26_rb4.jsonl 2>&1; cut -c1-120 $O/run26_rb4.jsonl
python tests/scripts/quick_rate.py config4 > $O/run26_default.jsonl 2>&1; cut -c1-120 $O/run26_default.jsonl
