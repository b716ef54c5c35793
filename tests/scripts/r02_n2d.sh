set -x
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/n2d_pytest.log 2>&1; tail -2 $O/n2d_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_n2d_ref.json 2> $O/bench_n2d_ref.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2d.json 2> $O/bench_n2d.err
tail -3 $O/bench_n2d.err
