// feature set FT_FULL (see xrt_trace.cuh)
#include "xrt_variants.h"
namespace xrt {
TraceKernel trace_kernel_full(int split, uint32_t, bool hist, size_t *smem) {
    *smem = block_smem_bytes<FT_FULL>();
    return trace_kernel_ft<FT_FULL, true>(split, hist);
}
MeshCoarseKernel mesh_coarse_kernel_full(bool hist, size_t *smem) {
    *smem = mesh_coarse_smem_bytes<FT_FULL>();
    return hist ? k_mesh_coarse<FT_FULL, true> : k_mesh_coarse<FT_FULL, false>;
}
MeshRefineKernel mesh_refine_kernel_full(bool hist) {
    return hist ? k_mesh_refine<FT_FULL, true> : k_mesh_refine<FT_FULL, false>;
}
void record_launch_full(int mode, uint32_t, const RecordLaunch &a) { record_launch_ft<FT_FULL>(mode, a); }
}  // namespace xrt
