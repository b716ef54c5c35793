set -x
O=gpurun_out/r02; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_statistics.py -m gpu -x -q -k "two_kernel_path or fused_kernel" ) > $O/run27_pytest.log 2>&1; tail -4 $O/run27_pytest.log
