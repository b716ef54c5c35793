set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 > $O/run16_default.jsonl 2>&1; cut -c1-120 $O/run16_default.jsonl
XRT_LIB_PATH=$PWD/build/var/libxrt_rb3.so python tests/scripts/quick_rate.py config4 > $O/run16_rb3.jsonl 2>&1; cut -c1-120 $O/run16_rb3.jsonl
( time python -m pytest tests -m gpu -x -q ) > $O/run16_pytest.log 2>&1; tail -5 $O/run16_pytest.log
