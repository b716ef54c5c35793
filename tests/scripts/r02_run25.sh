set -x
O=gpurun_out/r02; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/run25_pytest.log 2>&1; tail -4 $O/run25_pytest.log
( time python bench.py --steps 20 --warmup 3 ) > $O/bench_run25.json 2> $O/bench_run25.err; tail -3 $O/bench_run25.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_run25_ref.json 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/run25_smoke.log 2>&1; tail -1 $O/run25_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_run25_c4.csv python tests/scripts/quick_rate.py config4 --steps 2 > /dev/null 2>&1
Q="python tests/scripts/quick_rate.py --steps 1"
profiles/capture.sh $O/run25_c4_refine k_mesh_refine k_mesh_refineILj9ELb0 1e8 $Q config4
