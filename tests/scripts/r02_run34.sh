set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py plasma_mesh scene:torus_bragg scene:cylinder scene:plane_crystal_xy scene:apertures scene:mosaic_plane scene:sphere_voigt scene:mesh_sphere scene:plane_mirror > $O/run34_default.jsonl 2>&1; cut -c1-110 $O/run34_default.jsonl
XRT_NO_MESH_SORT=1 python tests/scripts/quick_rate.py plasma_mesh > $O/run34_nosort.jsonl 2>&1; cut -c1-110 $O/run34_nosort.jsonl
