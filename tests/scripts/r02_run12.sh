set -x
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "sorted_mesh or mesh_torus" 2>&1 | tail -5
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run12_c4_coarse k_mesh_coarse k_mesh_coarseILj9ELb0 1e8 $Q config4
profiles/capture.sh $O/run12_c4_trace k_trace k_traceILj9ELi0ELj0ELb0 1e8 $Q config4
