set -x
O=gpurun_out/r02; mkdir -p $O
for v in ms0 ms0po ms1po; do XRT_LIB_PATH=$PWD/build/var/libxrt_$v.so python tests/scripts/quick_rate.py config3 > $O/run18_$v.jsonl 2>&1; cut -c1-120 $O/run18_$v.jsonl; done
python tests/scripts/quick_rate.py config3 > $O/run18_default.jsonl 2>&1; cut -c1-120 $O/run18_default.jsonl
