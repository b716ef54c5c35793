set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 > $O/run11_default.jsonl 2>&1; cut -c1-220 $O/run11_default.jsonl
XRT_NO_MESH_SORT=1 python tests/scripts/quick_rate.py config4 > $O/run11_nosort.jsonl 2>&1; cut -c1-220 $O/run11_nosort.jsonl
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "sorted_mesh" 2>&1 | tail -15
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_run11_c4.csv python tests/scripts/quick_rate.py config4 --steps 2 > /dev/null 2>&1
grep -E "k_mesh|k_trace" $O/launches_run11_c4.csv | tail -8 | cut -c1-60,200-400
