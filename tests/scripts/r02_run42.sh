set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config5 config2 box focused > $O/run42_default.jsonl 2>&1; cut -c1-110 $O/run42_default.jsonl
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/run42_pytest.log 2>&1; tail -4 $O/run42_pytest.log
