// FP32 broad phase kernels (k_cull32), one per source kind, with and without the lost-sample emission
#include "xrt_variants.h"
namespace xrt {
CullKernel cull_kernel(int src_mode, bool hist) {
    switch (src_mode) {
    case CULL_POINT: return hist ? k_cull32<CULL_POINT, true> : k_cull32<CULL_POINT, false>;
    case CULL_BOX: return hist ? k_cull32<CULL_BOX, true> : k_cull32<CULL_BOX, false>;
    case CULL_FOCUSED: return hist ? k_cull32<CULL_FOCUSED, true> : k_cull32<CULL_FOCUSED, false>;
    default: return hist ? k_cull32<CULL_BUNDLES, true> : k_cull32<CULL_BUNDLES, false>;
    }
}
Mosaic32Kernel mosaic32_kernel(int src_mode, bool hist) {
    switch (src_mode) {
    case CULL_POINT: return hist ? k_mosaic32<CULL_POINT, true> : k_mosaic32<CULL_POINT, false>;
    case CULL_BOX: return hist ? k_mosaic32<CULL_BOX, true> : k_mosaic32<CULL_BOX, false>;
    default: return hist ? k_mosaic32<CULL_FOCUSED, true> : k_mosaic32<CULL_FOCUSED, false>;
    }
}
}  // namespace xrt
