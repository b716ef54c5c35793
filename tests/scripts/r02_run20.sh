set -x
O=gpurun_out/r02; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/run20_pytest.log 2>&1; tail -5 $O/run20_pytest.log
( time python bench.py --steps 20 --warmup 3 ) > $O/bench_run20.json 2> $O/bench_run20.err
tail -3 $O/bench_run20.err
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run20_c3_trace k_trace k_traceILj32ELi0ELj0ELb0 1e8 $Q config3
