set -x
O=gpurun_out/r02f; mkdir -p $O
python bench.py --steps 3 --warmup 3 --rays 1e8 --no-cpu --quick > $O/bench_1e8.json 2> $O/bench_1e8.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_1e8.csv python bench.py --steps 3 --warmup 3 --rays 1e8 --no-cpu --quick > $O/ncu_launches.log 2>&1
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/c2_cull k_cull32 k_cull32ILi0ELb0 1e9 $Q config2
profiles/capture.sh $O/c2_trace k_trace k_traceILj0ELi0ELj63ELb0 1e9 $Q config2
ls $O
