set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config5 config2 > $O/run39_default.jsonl 2>&1; cut -c1-110 $O/run39_default.jsonl
XRT_LIB_PATH=build/var/libxrt_head.so python tests/scripts/quick_rate.py config5 > $O/run39_head.jsonl 2>&1; cut -c1-110 $O/run39_head.jsonl
( time timeout 1500 python -m pytest tests/test_gpu_scale.py tests/test_gpu_statistics.py tests/test_plasma.py -m gpu -x -q -k "work_skipping or two_kernel_path or plasma" ) > $O/run39_pytest.log 2>&1; tail -4 $O/run39_pytest.log
