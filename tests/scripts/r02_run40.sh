set -x
O=gpurun_out/r02; mkdir -p $O
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run40_c5_cull k_cull32 k_cull32ILi3ELb0 1e9 $Q config5
