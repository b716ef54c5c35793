set -x
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tests/scripts/quick_rate.py config2 box focused doppler step config5 > gpurun_out/r02/run8_default.jsonl 2>&1
XRT_NO_STAGE2=1 python tests/scripts/quick_rate.py config2 config5 > gpurun_out/r02/run8_nostage2.jsonl 2>&1
cat gpurun_out/r02/run8_*.jsonl | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02/launches_run8_c2.csv python tests/scripts/quick_rate.py config2 --steps 3 > /dev/null 2>&1
grep -h "k_" gpurun_out/r02/launches_run8_c2.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -4
