set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config2 config2 doppler > $O/run29_default.jsonl 2>&1; cut -c1-120 $O/run29_default.jsonl
