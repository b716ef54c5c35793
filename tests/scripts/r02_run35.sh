set -x
O=gpurun_out/r02; mkdir -p $O
( time timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "mosaic_threshold or mosaic_sin or mosaic_planar" ) > $O/run35_pytest.log 2>&1; tail -5 $O/run35_pytest.log
