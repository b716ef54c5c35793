set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config5 > $O/run41_default.jsonl 2>&1; cut -c1-110 $O/run41_default.jsonl
XRT_LIB_PATH=build/var/libxrt_b3.so python tests/scripts/quick_rate.py config5 > $O/run41_b3.jsonl 2>&1; cut -c1-110 $O/run41_b3.jsonl
