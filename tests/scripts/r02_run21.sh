set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config4 box focused doppler step > $O/run21_default.jsonl 2>&1; cut -c1-120 $O/run21_default.jsonl
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "mesh" ) > $O/run21_pytest.log 2>&1; tail -5 $O/run21_pytest.log
