set -x
O=gpurun_out/r02; mkdir -p $O
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/n2b_pytest.log 2>&1; tail -3 $O/n2b_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2b.json 2> $O/bench_n2b.err
tail -3 $O/bench_n2b.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --quick > $O/bench_n2b_strong.json 2> $O/bench_n2b_strong.err
tail -3 $O/bench_n2b_strong.err
