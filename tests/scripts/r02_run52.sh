set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 plasma_mesh > $O/run52_default.jsonl 2>&1; cut -c1-110 $O/run52_default.jsonl
XRT_LIB_PATH=build/var/libxrt_head.so python tests/scripts/quick_rate.py config4 plasma_mesh > $O/run52_head.jsonl 2>&1; cut -c1-110 $O/run52_head.jsonl
( time timeout 1500 python -m pytest tests/test_gpu_scale.py tests/test_gpu_statistics.py -m gpu -x -q -k "mesh" ) > $O/run52_pytest.log 2>&1; tail -4 $O/run52_pytest.log
