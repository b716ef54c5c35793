set -x
O=gpurun_out/r02; mkdir -p $O
( time timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "mosaic or mesh" ) > $O/run22_pytest.log 2>&1; tail -5 $O/run22_pytest.log
