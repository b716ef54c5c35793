set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py config2 config3 config5 scene:sphere_voigt > $O/run47_default.jsonl 2>&1; cut -c1-110 $O/run47_default.jsonl
XRT_LIB_PATH=build/var/libxrt_head.so python tests/scripts/quick_rate.py config2 config3 config5 scene:sphere_voigt > $O/run47_head.jsonl 2>&1; cut -c1-110 $O/run47_head.jsonl
python bench.py --steps 5 --warmup 3 --no-cpu --quick > $O/run47_bench.json 2> $O/run47_bench.err; tail -2 $O/run47_bench.err
XRT_LIB_PATH=build/var/libxrt_head.so python bench.py --steps 5 --warmup 3 --no-cpu --quick > $O/run47_bench_head.json 2> $O/run47_bench_head.err
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/run47_pytest.log 2>&1; tail -4 $O/run47_pytest.log
