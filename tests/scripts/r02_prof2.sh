set -x
mkdir -p gpurun_out/r02
python tests/scripts/quick_rate.py config2 config5 config5 > gpurun_out/r02/k1k2b_default.jsonl 2>&1
cat gpurun_out/r02/k1k2b_default.jsonl | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02/launches_k1k2_c2.csv python tests/scripts/quick_rate.py config2 --steps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02/launches_k1k2_c5.csv python tests/scripts/quick_rate.py config5 --steps 3 > /dev/null 2>&1
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh gpurun_out/r02/k1_c2 k_cull32 k_cull32ILi0ELb0 1e9 $Q config2
profiles/capture.sh gpurun_out/r02/k2_c2 k_trace k_traceILj0ELi0ELj63ELb0 1e9 $Q config2
profiles/capture.sh gpurun_out/r02/k1_c5 k_cull32 k_cull32ILi3ELb0 1e9 $Q config5
profiles/capture.sh gpurun_out/r02/k2_c5 k_trace k_traceILj128ELi0ELj0ELb0 1e9 $Q config5
ls gpurun_out/r02
