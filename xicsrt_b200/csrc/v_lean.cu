// lean feature set: plane / sphere optics, box source with an isotropic cone, constant / uniform / normal line;
// includes the pre-instantiated spherical-crystal spectrometer (KN_SPECTROMETER).
#include "xrt_variants.h"
namespace xrt {
TraceKernel trace_kernel_lean(int split, uint32_t known, bool hist, size_t *smem) {
    if (split == 0 && (known & KN_SPECTROMETER) == KN_SPECTROMETER) {
        *smem = block_smem_bytes<0, KN_SPECTROMETER>();
        return hist ? k_trace<0, 0, KN_SPECTROMETER, true> : k_trace<0, 0, KN_SPECTROMETER, false>;
    }
    *smem = block_smem_bytes<0>();
    return trace_kernel_ft<0, true>(split, hist);
}
void record_launch_lean(int mode, uint32_t known, const RecordLaunch &a) {
    if (mode == REC_PHILOX && a.split == 0 && (known & KN_SPECTROMETER) == KN_SPECTROMETER)
        record_launch_ft<0, KN_SPECTROMETER>(mode, a);
    else
        record_launch_ft<0>(mode, a);
}
}  // namespace xrt
