# ncu captures of every benchmarked kernel at HEAD (digested on the box), plus the launch list of the bench command
set -x
O=gpurun_out/r02h; mkdir -p $O
python bench.py --steps 3 --warmup 3 --rays 1e8 --no-cpu > $O/bench_1e8.json 2> $O/bench_1e8.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_1e8.csv python bench.py --steps 3 --warmup 3 --rays 1e8 --no-cpu > $O/ncu_launches.log 2>&1
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/c2_cull k_cull32 k_cull32ILi0ELb0 1e9 $Q config2
profiles/capture.sh $O/c2_trace k_trace k_traceILj0ELi0ELj63ELb0 1e9 $Q config2
profiles/capture.sh $O/c3_trace k_trace k_traceILj32ELi0ELj0ELb0 1e8 $Q config3
profiles/capture.sh $O/c4_trace k_trace k_traceILj9ELi0ELj0ELb0 1e8 $Q config4
profiles/capture.sh $O/c5_cull k_cull32 k_cull32ILi3ELb0 1e9 $Q config5
profiles/capture.sh $O/c5_trace k_trace k_traceILj128ELi0ELj0ELb0 1e9 $Q config5
profiles/capture.sh $O/record k_record k_recordILj0ELi0ELj63 16777216 python bench.py --steps 1 --warmup 1 --no-cpu --rays 1e8
ls -la $O
