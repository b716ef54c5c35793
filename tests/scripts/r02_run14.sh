set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 > $O/run14_default.jsonl 2>&1; cut -c1-120 $O/run14_default.jsonl
XRT_MESH_TILE=2 python tests/scripts/quick_rate.py config4 > $O/run14_tile2.jsonl 2>&1; cut -c1-120 $O/run14_tile2.jsonl
XRT_MESH_TILE=1 python tests/scripts/quick_rate.py config4 > $O/run14_tile1.jsonl 2>&1; cut -c1-120 $O/run14_tile1.jsonl
XRT_LIB_PATH=$PWD/build/var/libxrt_rb3.so python tests/scripts/quick_rate.py config4 > $O/run14_rb3.jsonl 2>&1; cut -c1-120 $O/run14_rb3.jsonl
XRT_MESH_TILE=2 XRT_LIB_PATH=$PWD/build/var/libxrt_rb3.so python tests/scripts/quick_rate.py config4 > $O/run14_rb3t2.jsonl 2>&1; cut -c1-120 $O/run14_rb3t2.jsonl
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "sorted_mesh or mesh_torus" > $O/run14_pytest.log 2>&1; tail -5 $O/run14_pytest.log
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run14_c4_refine k_mesh_refine k_mesh_refineILj9ELb0 1e8 $Q config4
