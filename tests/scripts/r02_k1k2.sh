set -x
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tests/scripts/quick_rate.py config2 box focused doppler step config5 nobroad nocull config3 config4 > gpurun_out/r02/k1k2_default.jsonl 2>&1
for v in cb8 cb5 cu1 cu3 u1b3; do
  XRT_LIB_PATH=$PWD/build/var/libxrt_$v.so python tests/scripts/quick_rate.py config2 box config5 > gpurun_out/r02/k1k2_$v.jsonl 2>&1
done
cat gpurun_out/r02/k1k2_*.jsonl | cut -c1-330
