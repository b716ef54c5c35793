set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config3 scene:mosaic_sphere > $O/run51_default.jsonl 2>&1; cut -c1-110 $O/run51_default.jsonl
XRT_LIB_PATH=build/var/libxrt_head.so python tests/scripts/quick_rate.py config3 scene:mosaic_sphere > $O/run51_head.jsonl 2>&1; cut -c1-110 $O/run51_head.jsonl
( time timeout 1500 python -m pytest tests/test_gpu_scale.py tests/test_gpu_statistics.py tests/test_gpu_parity.py -m gpu -x -q -k "mosaic or two_kernel or config3 or work_skipping" ) > $O/run51_pytest.log 2>&1; tail -4 $O/run51_pytest.log
