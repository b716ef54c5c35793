// feature set FT_SRCLEAN (see xrt_trace.cuh)
#include "xrt_variants.h"
namespace xrt {
TraceKernel trace_kernel_src(int split, uint32_t, bool hist, size_t *smem) {
    *smem = block_smem_bytes<FT_SRCLEAN>();
    return trace_kernel_ft<FT_SRCLEAN, false>(split, hist);
}
void record_launch_src(int mode, uint32_t, const RecordLaunch &a) { record_launch_ft<FT_SRCLEAN>(mode, a); }
}  // namespace xrt
