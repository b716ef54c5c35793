// feature set FT_MID (see xrt_trace.cuh)
#include "xrt_variants.h"
namespace xrt {
TraceKernel trace_kernel_mid(int split, uint32_t, bool hist, size_t *smem) {
    *smem = block_smem_bytes<FT_MID>();
    return trace_kernel_ft<FT_MID, true>(split, hist);
}
void record_launch_mid(int mode, uint32_t, const RecordLaunch &a) { record_launch_ft<FT_MID>(mode, a); }
}  // namespace xrt
