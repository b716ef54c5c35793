// feature set FT_MOSAICLEAN (see xrt_trace.cuh)
#include "xrt_variants.h"
namespace xrt {
TraceKernel trace_kernel_mosaic(int split, uint32_t, bool hist, size_t *smem) {
    *smem = block_smem_bytes<FT_MOSAICLEAN>();
    return trace_kernel_ft<FT_MOSAICLEAN, false>(split, hist);
}
void record_launch_mosaic(int mode, uint32_t, const RecordLaunch &a) { record_launch_ft<FT_MOSAICLEAN>(mode, a); }
}  // namespace xrt
