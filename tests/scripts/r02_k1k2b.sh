set -x
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tests/scripts/quick_rate.py config2 box focused doppler step config5 > gpurun_out/r02/k1k2c_default.jsonl 2>&1
for v in mb2 u1 cb8 cu3bb5; do
  XRT_LIB_PATH=$PWD/build/var/libxrt_$v.so python tests/scripts/quick_rate.py config2 box config5 > gpurun_out/r02/k1k2c_$v.jsonl 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02/launches_k1k2c_c2.csv python tests/scripts/quick_rate.py config2 --steps 3 > /dev/null 2>&1
grep -h "k_" gpurun_out/r02/launches_k1k2c_c2.csv | cut -d, -f5,15 | tail -6
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh gpurun_out/r02/k1c_c2 k_cull32 k_cull32ILi0ELb0 1e9 $Q config2
profiles/capture.sh gpurun_out/r02/k2c_c2 k_trace k_traceILj0ELi0ELj63ELb0 1e9 $Q config2
