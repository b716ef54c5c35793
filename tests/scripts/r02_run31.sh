set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config3 config2 > $O/run31_default.jsonl 2>&1; cut -c1-120 $O/run31_default.jsonl
for v in q128 q96; do XRT_LIB_PATH=$PWD/build/var/libxrt_$v.so python tests/scripts/quick_rate.py config3 > $O/run31_$v.jsonl 2>&1; cut -c1-120 $O/run31_$v.jsonl; done
XRT_LIB_PATH=$PWD/build/var/libxrt_q128.so timeout 600 python -m pytest tests -m gpu -x -q -k "mosaic" 2>&1 | tail -3
