# ncu captures of the config 3 / config 4 kernels and of the history replay at the head of the round (after r02_final.sh)
set -x
O=gpurun_out/r02f; mkdir -p $O
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/c4_coarse k_mesh_coarse k_mesh_coarseILj9ELb0 1e8 $Q config4
profiles/capture.sh $O/c4_refine k_mesh_refine k_mesh_refineILj9ELb0 1e8 $Q config4
profiles/capture.sh $O/c3_mosaic32 k_mosaic32 k_mosaic32ILi0ELb0 1e8 $Q config3
profiles/capture.sh $O/c3_trace k_trace k_traceILj32ELi0ELj0ELb0 1e8 $Q config3
profiles/capture.sh $O/record k_record k_recordILj0ELi0ELj63 16777216 python bench.py --steps 1 --warmup 1 --no-cpu --quick --rays 1e8
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_config4.csv $Q config4 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_config3.csv $Q config3 > /dev/null 2>&1
