set -x
O=gpurun_out/r02; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8b.json 2> $O/bench_n8b.err
tail -3 $O/bench_n8b.err
