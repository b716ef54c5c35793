set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 > $O/run15_default.jsonl 2>&1; cut -c1-120 $O/run15_default.jsonl
XRT_MESH_SUB=1 python tests/scripts/quick_rate.py config4 > $O/run15_sub1.jsonl 2>&1; cut -c1-120 $O/run15_sub1.jsonl
XRT_MESH_SUB=3 python tests/scripts/quick_rate.py config4 > $O/run15_sub3.jsonl 2>&1; cut -c1-120 $O/run15_sub3.jsonl
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "sorted_mesh or mesh_torus" > $O/run15_pytest.log 2>&1; tail -5 $O/run15_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_run15_c4.csv python tests/scripts/quick_rate.py config4 --steps 2 > /dev/null 2>&1
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run15_c4_refine k_mesh_refine k_mesh_refineILj9ELb0 1e8 $Q config4
