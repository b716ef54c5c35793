set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py plasma_mesh config4 > $O/run36_default.jsonl 2>&1; cut -c1-110 $O/run36_default.jsonl
XRT_NO_MESH_SORT=1 python tests/scripts/quick_rate.py plasma_mesh config4 > $O/run36_nosort.jsonl 2>&1; cut -c1-110 $O/run36_nosort.jsonl
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "mesh" ) > $O/run36_pytest.log 2>&1; tail -4 $O/run36_pytest.log
