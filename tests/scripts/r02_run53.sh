set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 > $O/run53_default.jsonl 2>&1; cut -c1-110 $O/run53_default.jsonl
for v in nb4 nb4t2 all; do
XRT_LIB_PATH=build/var/libxrt_$v.so python tests/scripts/quick_rate.py config4 > $O/run53_$v.jsonl 2>&1; cut -c1-110 $O/run53_$v.jsonl
done
