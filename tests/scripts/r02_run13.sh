set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 > $O/run13_default.jsonl 2>&1; cut -c1-120 $O/run13_default.jsonl
XRT_NO_MESH_DIRGRID=1 python tests/scripts/quick_rate.py config4 > $O/run13_nogrid.jsonl 2>&1; cut -c1-120 $O/run13_nogrid.jsonl
XRT_MESH_TILE=2 python tests/scripts/quick_rate.py config4 > $O/run13_tile2.jsonl 2>&1; cut -c1-120 $O/run13_tile2.jsonl
XRT_MESH_TILE=5 python tests/scripts/quick_rate.py config4 > $O/run13_tile5.jsonl 2>&1; cut -c1-120 $O/run13_tile5.jsonl
XRT_LIB_PATH=$PWD/build/var/libxrt_rb3.so python tests/scripts/quick_rate.py config4 > $O/run13_rb3.jsonl 2>&1; cut -c1-120 $O/run13_rb3.jsonl
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "sorted_mesh or mesh_torus" 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_run13_c4.csv python tests/scripts/quick_rate.py config4 --steps 2 > /dev/null 2>&1
grep -E "k_mesh|k_trace" $O/launches_run13_c4.csv | tail -5 | cut -c50-75,200-400
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run13_c4_refine k_mesh_refine k_mesh_refineILj9ELb0 1e8 $Q config4
profiles/capture.sh $O/run13_c4_coarse k_mesh_coarse k_mesh_coarseILj9ELb0 1e8 $Q config4
