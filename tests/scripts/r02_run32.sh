set -x
O=gpurun_out/r02; mkdir -p $O
( time timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "sorted_mesh" ) > $O/run32_pytest.log 2>&1; tail -5 $O/run32_pytest.log
timeout 900 python -m pytest tests -m gpu -x -q -k "mesh or plasma" 2>&1 | tail -3
