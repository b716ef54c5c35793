# quick check of the in-tree build on a GPU box: smoke, the injected-parity tests, the scene cache test, bench --quick
set -x
O=gpurun_out/r02s; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistics.py tests/test_plasma.py -m gpu -x -q ) > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu --quick > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench rc=$?"
