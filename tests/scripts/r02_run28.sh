set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/e2e_profile.py config4 > $O/run28_e2e_c4.log 2>&1
python tests/scripts/e2e_profile.py config2 > $O/run28_e2e_c2.log 2>&1
