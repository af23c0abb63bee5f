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

// sc must already carry the device pointers of shapes / positions / indices.  The n entries of prim_shape / prim_local
// are primitive REFERENCES: a large triangle may appear several times, each time with its own box in d_ref_box (6
// floats: lo.xyz, hi.xyz; the part of the triangle inside one cell of a spatial split, scene.cu split_references);
// d_ref_box == nullptr: one reference per primitive, boxes computed from the primitives.
cudaError_t build_bvh8(const DevScene &sc, const int *d_prim_shape, const int *d_prim_local, const float *d_ref_box, int n,
                       cudaStream_t stream, BvhResult *out);

}  // namespace lj
