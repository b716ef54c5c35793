set -x
mkdir -p gpurun_out/r02
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02/run10_pytest.log 2>&1; tail -5 gpurun_out/r02/run10_pytest.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02/bench_run10.json 2> gpurun_out/r02/bench_run10.err
tail -5 gpurun_out/r02/bench_run10.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02/bench_run10_ref.json 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/run10_smoke.log 2>&1; tail -2 gpurun_out/r02/run10_smoke.log
