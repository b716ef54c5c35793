// feature set FT_MESHLEAN (see xrt_trace.cuh)
#include "xrt_variants.h"
namespace xrt {
TraceKernel trace_kernel_mesh(int split, uint32_t, bool hist, size_t *smem) {
    *smem = block_smem_bytes<FT_MESHLEAN>();
    return trace_kernel_ft<FT_MESHLEAN, true>(split, hist);
}
MeshCoarseKernel mesh_coarse_kernel_mesh(bool hist, size_t *smem) {
    *smem = mesh_coarse_smem_bytes<FT_MESHLEAN>();
    return hist ? k_mesh_coarse<FT_MESHLEAN, true> : k_mesh_coarse<FT_MESHLEAN, false>;
}
MeshRefineKernel mesh_refine_kernel_mesh(bool hist) {
    return hist ? k_mesh_refine<FT_MESHLEAN, true> : k_mesh_refine<FT_MESHLEAN, false>;
}
void record_launch_mesh(int mode, uint32_t, const RecordLaunch &a) { record_launch_ft<FT_MESHLEAN>(mode, a); }
}  // namespace xrt
