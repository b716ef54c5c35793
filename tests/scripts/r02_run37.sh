set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config2 config5 box focused doppler > $O/run37_default.jsonl 2>&1; cut -c1-110 $O/run37_default.jsonl
( time timeout 1500 python -m pytest tests/test_gpu_scale.py tests/test_gpu_statistics.py -m gpu -x -q -k "work_skipping or two_kernel_path" ) > $O/run37_pytest.log 2>&1; tail -4 $O/run37_pytest.log
