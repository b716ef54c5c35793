set -x
O=gpurun_out/r02; mkdir -p $O
python bench.py --steps 5 --warmup 3 --no-cpu --quick > $O/run50_bench.json 2> $O/run50_bench.err; tail -2 $O/run50_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu --quick > $O/run50_bench20.json 2> $O/run50_bench20.err
nvidia-smi --query-gpu=timestamp,index,clocks.sm --format=csv,noheader,nounits -i 0; date
