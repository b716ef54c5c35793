# final validation of the round at HEAD: GPU tests, bench (+ reference arm), smoke, launch list and ncu captures of the config-2 kernels
set -x
O=gpurun_out/r02f; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1; tail -4 $O/pytest.log
( time python bench.py --steps 20 --warmup 3 ) > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --steps 3 --warmup 3 --rays 1e8 --no-cpu --quick > $O/bench_1e8.json 2> $O/bench_1e8.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_1e8.csv python bench.py --steps 3 --warmup 3 --rays 1e8 --no-cpu --quick > $O/ncu_launches.log 2>&1
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/c2_cull k_cull32 k_cull32ILi0ELb0 1e9 $Q config2
profiles/capture.sh $O/c2_trace k_trace k_traceILj0ELi0ELj63ELb0 1e9 $Q config2
profiles/capture.sh $O/c5_cull k_cull32 k_cull32ILi3ELb0 1e9 $Q config5
