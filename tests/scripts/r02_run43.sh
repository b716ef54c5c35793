set -x
O=gpurun_out/r02; mkdir -p $O
( time timeout 1500 python -m pytest tests/test_gpu_statistics.py -m gpu -x -q ) > $O/run43_pytest.log 2>&1; tail -4 $O/run43_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu --quick > $O/run43_bench.json 2> $O/run43_bench.err; tail -2 $O/run43_bench.err
XRT_SCENE_CACHE=0 python bench.py --steps 20 --warmup 3 --no-cpu --quick > $O/run43_bench_nocache.json 2> $O/run43_bench_nocache.err
python tests/scripts/e2e_profile.py config2 > $O/run43_e2e_c2.log 2>&1
