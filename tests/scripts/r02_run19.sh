set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config3 > $O/run19_default.jsonl 2>&1; cut -c1-120 $O/run19_default.jsonl
XRT_NO_MOSAIC32=1 python tests/scripts/quick_rate.py config3 > $O/run19_nom32.jsonl 2>&1; cut -c1-120 $O/run19_nom32.jsonl
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "mosaic" ) > $O/run19_pytest.log 2>&1; tail -5 $O/run19_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_run19_c3.csv python tests/scripts/quick_rate.py config3 --steps 2 > /dev/null 2>&1
grep -E "k_mosaic32|k_trace" $O/launches_run19_c3.csv | tail -4 | cut -c50-80,200-400
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run19_c3_m32 k_mosaic32 k_mosaic32ILi0ELb0 1e8 $Q config3
