// bvh_build.cuh -- host entry point of the GPU BVH build.
#pragma once
#include "lj_cuda.h"
#include <string.h>

#include "lj_bvh_build.h"

namespace lj {

struct BvhResult {
    DevNode8 *nodes = nullptr;  // device, num_nodes entries, root = 0
    DevPrim *prims = nullptr;   // device, leaf order
    int num_nodes = 0;
    int num_prims_placed = 0;   // must equal the primitive count
    int depth = 0;              // levels of wide nodes
    int ploc_rounds = 0;
    Box3 bounds;                // union of the fp32 primitive boxes (what rtcGetSceneBounds returns, scene.cpp:29-31)
    double sah_cost = 0;
    int launches = 0;
};

// sc must already carry the device pointers of shapes / positions / indices.
cudaError_t build_bvh8(const DevScene &sc, const int *d_prim_shape, const int *d_prim_local, int n,
                       cudaStream_t stream, BvhResult *out);

}  // namespace lj
