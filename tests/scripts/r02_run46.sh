set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config5 > $O/run46_a.jsonl 2>&1; cut -c1-110 $O/run46_a.jsonl
python tests/scripts/quick_rate.py config5 --flush > $O/run46_b.jsonl 2>&1; cut -c1-110 $O/run46_b.jsonl
python tests/scripts/quick_rate.py config5 --flush --steps 20 > $O/run46_c.jsonl 2>&1; cut -c1-110 $O/run46_c.jsonl
python tests/scripts/quick_rate.py config2 --flush > $O/run46_d.jsonl 2>&1; cut -c1-110 $O/run46_d.jsonl
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv
