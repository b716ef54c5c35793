# device rates of the other scene families and config-2 variants at the head of the round
set -x
O=gpurun_out/r02; mkdir -p $O
python tests/scripts/quick_rate.py plasma_mesh scene:torus_bragg scene:cylinder scene:plane_crystal_xy scene:apertures scene:mosaic_plane scene:sphere_voigt scene:mesh_sphere scene:plane_mirror > $O/families_head.jsonl 2>&1; cut -c1-110 $O/families_head.jsonl
python tests/scripts/quick_rate.py box focused doppler step > $O/variants_head.jsonl 2>&1; cut -c1-110 $O/variants_head.jsonl
