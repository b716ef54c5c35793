set -x
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tests/scripts/quick_rate.py config2 box focused doppler step config5 config3 config4 > gpurun_out/r02/run3_default.jsonl 2>&1
XRT_LIB_PATH=$PWD/build/var/libxrt_u1.so python tests/scripts/quick_rate.py config2 > gpurun_out/r02/run3_u1.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02/launches_run3_c2.csv python tests/scripts/quick_rate.py config2 --steps 3 > /dev/null 2>&1
grep -h "k_" gpurun_out/r02/launches_run3_c2.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -4
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh gpurun_out/r02/run3_k1_c2 k_cull32 k_cull32ILi0ELb0 1e9 $Q config2
profiles/capture.sh gpurun_out/r02/run3_k2_c2 k_trace k_traceILj0ELi0ELj63ELb0 1e9 $Q config2
profiles/capture.sh gpurun_out/r02/run3_c3 k_trace k_traceILj32ELi0ELj0ELb0 1e8 $Q config3
