set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config3 > $O/run17_default.jsonl 2>&1; cut -c1-120 $O/run17_default.jsonl
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "mosaic" ) > $O/run17_pytest.log 2>&1; tail -5 $O/run17_pytest.log
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run17_c3_trace k_trace k_traceILj32ELi0ELj0ELb0 1e8 $Q config3
