set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 > $O/run24_default.jsonl 2>&1; cut -c1-120 $O/run24_default.jsonl
XRT_LIB_PATH=$PWD/build/var/libxrt_rb2.so python tests/scripts/quick_rate.py config4 > $O/run24_rb2.jsonl 2>&1; cut -c1-120 $O/run24_rb2.jsonl
timeout 900 python -m pytest tests -m gpu -x -q -k "mesh" 2>&1 | tail -3
