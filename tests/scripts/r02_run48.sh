set -x
O=gpurun_out/r02; mkdir -p $O
profiles/capture.sh $O/run48_record k_record k_recordILj0ELi0ELj63 16777216 python bench.py --steps 1 --warmup 1 --no-cpu --quick --rays 1e8
