set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config2 config3 config4 config5 > $O/run49_default.jsonl 2>&1; cut -c1-110 $O/run49_default.jsonl
XRT_LIB_PATH=build/var/libxrt_head.so python tests/scripts/quick_rate.py config2 config3 config4 config5 > $O/run49_head.jsonl 2>&1; cut -c1-110 $O/run49_head.jsonl
python bench.py --steps 5 --warmup 3 --no-cpu --quick > $O/run49_bench.json 2> $O/run49_bench.err; tail -2 $O/run49_bench.err
XRT_LIB_PATH=build/var/libxrt_head.so python bench.py --steps 5 --warmup 3 --no-cpu --quick > $O/run49_bench_head.json 2> $O/run49_bench_head.err
