// xrt_mesh.cuh -- triangle-mesh optics (xicsrt/optics/_ShapeMesh.py).
#pragma once
#include "../../include/xrt.h"
#include "xrt_math.cuh"

namespace xrt {

// Filled in with the mesh row of the scope table; until then a scene holding a
// mesh optic is refused by xrt_scene_create (XRT_EUNSUPPORTED), so this is
// never reached.
__device__ __forceinline__ bool mesh_intersect(const XrtOpticDesc &, V3, V3, V3 &X, V3 &n) {
    X = nan3();
    n = nan3();
    return false;
}

}  // namespace xrt
