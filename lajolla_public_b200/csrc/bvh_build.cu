// bvh_build.cu -- GPU BVH construction kernels (sm_100a).  See lj_bvh_build.h for the stages.
// The one library primitive used is cub::DeviceRadixSort for the Morton keys (one-off build step,
// not on the per-sample path).
#include "bvh_build.cuh"

#include <stdlib.h>

#if defined(LJ_HOSTSIM)
#include <numeric>
#include <vector>
#else
#include <cub/device/device_radix_sort.cuh>
#endif

namespace lj {

namespace {

__device__ __forceinline__ int float_order_key(float f) {
    int i = lj_float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float float_from_order_key(int k) {
    int i = k >= 0 ? k : k ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f;
    memcpy(&f, &i, 4);
    return f;
#endif
}

__global__ void k_prim_boxes(const LJ_GRID_CONSTANT DevScene sc, const int *prim_shape, const int *prim_local, int n,
                             DevPrim *prims_unsorted, Box3 *boxes, int *scene_bounds /*6 ordered ints*/) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Box3 b = box_empty();
    if (i < n) {
        prims_unsorted[i] = make_prim(sc, prim_shape[i], prim_local[i], b);
        boxes[i] = b;
    }
    // warp reduce then one atomic per warp per component
    float v[6] = {b.lo.x, b.lo.y, b.lo.z, b.hi.x, b.hi.y, b.hi.z};
#pragma unroll
    for (int c = 0; c < 6; c++) {
        float x = c < 3 ? lj_warp_min(v[c]) : lj_warp_max(v[c]);
        if (LJ_LANE() == 0) {
            if (c < 3) atomicMin(&scene_bounds[c], float_order_key(x));
            else atomicMax(&scene_bounds[c], float_order_key(x));
        }
    }
}

__global__ void k_morton(const Box3 *boxes, int n, const int *scene_bounds, uint64_t *keys, uint32_t *vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Box3 sb;
    sb.lo = mk3(float_from_order_key(scene_bounds[0]), float_from_order_key(scene_bounds[1]), float_from_order_key(scene_bounds[2]));
    sb.hi = mk3(float_from_order_key(scene_bounds[3]), float_from_order_key(scene_bounds[4]), float_from_order_key(scene_bounds[5]));
    Box3 b = boxes[i];
    keys[i] = morton63((b.lo + b.hi) * 0.5f, sb);
    vals[i] = (uint32_t)i;
}

__global__ void k_gather(const uint32_t *order, int n, const DevPrim *prims_unsorted, const Box3 *boxes, DevPrim *prims, Box3 *leaf_box) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t src = order[i];
    prims[i] = prims_unsorted[src];
    leaf_box[i] = boxes[src];
}

__global__ void k_hierarchy(const uint64_t *keys, int n, int *left, int *right, int *parent_internal, int *parent_leaf,
                            int *range_first, int *range_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    karras_node(keys, n, i, left, right, parent_internal, parent_leaf, range_first, range_count);
}

__global__ void k_refit(int n, const Box3 *leaf_box, Box3 *node_box, const int *left, const int *right,
                        const int *parent_internal, const int *parent_leaf, int *visit) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    refit_from_leaf(k, leaf_box, node_box, left, right, parent_internal, parent_leaf, visit);
}

__global__ void k_emit2(int n_internal, const Box3 *leaf_box, const Box3 *node_box, const int *left, const int *right,
                        const int *range_first, const int *range_count, int max_leaf, DevNode2 *nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    nodes[i] = emit_node2(i, leaf_box, node_box, left, right, range_first, range_count, max_leaf);
}

// SAH cost of the binary tree: sum over internal nodes of A(node)/A(root) * 1.2 + leaves * 1.0
__global__ void k_sah(int n_internal, const Box3 *node_box, const Box3 *leaf_box, int n, double *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0;
    float root = box_half_area(node_box[0]);
    if (i < n_internal) c += 1.2 * box_half_area(node_box[i]) / root;
    if (i < n) c += 1.0 * box_half_area(leaf_box[i]) / root;
    c = lj_warp_sum(c);
    if (LJ_LANE() == 0 && c != 0) atomicAdd(out, c);
}

template <typename T>
cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)); }

}  // namespace

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

cudaError_t build_bvh2(const DevScene &sc, const int *d_prim_shape, const int *d_prim_local, int n,
                       cudaStream_t stream, BvhResult *out) {
    // leaf size: LJ_BVH_MAX_LEAF in [1, 8] (default 4)
    int max_leaf = 4;
    if (const char *e = getenv("LJ_BVH_MAX_LEAF")) max_leaf = atoi(e);
    max_leaf = max_leaf < 1 ? 1 : (max_leaf > 8 ? 8 : max_leaf);
    cudaError_t err = cudaSuccess;
    DevPrim *prims_unsorted = nullptr, *prims = nullptr;
    Box3 *boxes = nullptr, *leaf_box = nullptr, *node_box = nullptr;
    int *range_first = nullptr, *range_count = nullptr;
    int *scene_bounds = nullptr, *left = nullptr, *right = nullptr, *parent_internal = nullptr, *parent_leaf = nullptr, *visit = nullptr;
    uint64_t *keys = nullptr, *keys_sorted = nullptr;
    uint32_t *vals = nullptr, *vals_sorted = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    DevNode2 *nodes = nullptr;
    double *d_sah = nullptr;
    const int T = 256;
    const int nb = (n + T - 1) / T;
    const int n_internal = n > 1 ? n - 1 : 1;
    int h_bounds[6];
    int init_bounds[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};

    CK(dalloc(&prims_unsorted, n)); CK(dalloc(&prims, n));
    CK(dalloc(&boxes, n)); CK(dalloc(&leaf_box, n)); CK(dalloc(&node_box, n_internal));
    CK(dalloc(&scene_bounds, 6)); CK(dalloc(&left, n_internal)); CK(dalloc(&right, n_internal));
    CK(dalloc(&parent_internal, n_internal)); CK(dalloc(&parent_leaf, n)); CK(dalloc(&visit, n_internal));
    CK(dalloc(&keys, n)); CK(dalloc(&keys_sorted, n)); CK(dalloc(&vals, n)); CK(dalloc(&vals_sorted, n));
    CK(dalloc(&nodes, n_internal)); CK(dalloc(&d_sah, 1));
    CK(dalloc(&range_first, n_internal)); CK(dalloc(&range_count, n_internal));
    CK(cudaMemcpyAsync(scene_bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(parent_internal, 0xff, sizeof(int) * n_internal, stream));
    CK(cudaMemsetAsync(parent_leaf, 0xff, sizeof(int) * n, stream));
    CK(cudaMemsetAsync(visit, 0, sizeof(int) * n_internal, stream));
    CK(cudaMemsetAsync(d_sah, 0, sizeof(double), stream));

    LJ_LAUNCH(k_prim_boxes, nb, T, stream, sc, d_prim_shape, d_prim_local, n, prims_unsorted, boxes, scene_bounds);
    LJ_LAUNCH(k_morton, nb, T, stream, boxes, n, scene_bounds, keys, vals);
#if defined(LJ_HOSTSIM)
    {
        std::vector<uint32_t> order(n);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });
        for (int i = 0; i < n; i++) { keys_sorted[i] = keys[order[i]]; vals_sorted[i] = vals[order[i]]; }
        (void)cub_bytes;
    }
#else
    CK(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys, keys_sorted, vals, vals_sorted, n, 0, 63, stream));
    CK(cudaMalloc(&cub_tmp, cub_bytes ? cub_bytes : 1));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys, keys_sorted, vals, vals_sorted, n, 0, 63, stream));
#endif
    LJ_LAUNCH(k_gather, nb, T, stream, vals_sorted, n, prims_unsorted, boxes, prims, leaf_box);
    if (n > 1) {
        LJ_LAUNCH(k_hierarchy, nb, T, stream, keys_sorted, n, left, right, parent_internal, parent_leaf, range_first, range_count);
        LJ_LAUNCH(k_refit, nb, T, stream, n, leaf_box, node_box, left, right, parent_internal, parent_leaf, visit);
        LJ_LAUNCH(k_emit2, nb, T, stream, n - 1, leaf_box, node_box, left, right, range_first, range_count, max_leaf, nodes);
        LJ_LAUNCH(k_sah, nb, T, stream, n - 1, node_box, leaf_box, n, d_sah);
        out->launches = 8;
    } else {
        // single primitive: node 0 = { leaf 0, empty box }
        Box3 lb;
        CK(cudaMemcpyAsync(&lb, leaf_box, sizeof(Box3), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        DevNode2 nd;
        float inf = INFINITY;
        nd.n0 = mk4(lb.lo.x, lb.hi.x, lb.lo.y, lb.hi.y);
        nd.n1 = mk4(inf, -inf, inf, -inf);
        nd.n2 = mk4(lb.lo.z, lb.hi.z, inf, -inf);
        nd.n3 = mk4(u2f((uint32_t)~0), u2f((uint32_t)~0), 0.f, 0.f);
        CK(cudaMemcpyAsync(nodes, &nd, sizeof(nd), cudaMemcpyHostToDevice, stream));
        out->launches = 4;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_bounds, scene_bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&out->sah_cost, d_sah, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int c = 0; c < 3; c++) {
        (&out->bounds.lo.x)[c] = float_from_order_key(h_bounds[c]);
        (&out->bounds.hi.x)[c] = float_from_order_key(h_bounds[3 + c]);
    }
    out->nodes = nodes;
    out->prims = prims;
    out->num_nodes = n_internal;
    nodes = nullptr;
    prims = nullptr;
done:
    cudaFree(prims_unsorted); cudaFree(prims); cudaFree(boxes); cudaFree(leaf_box); cudaFree(node_box);
    cudaFree(scene_bounds); cudaFree(left); cudaFree(right); cudaFree(parent_internal); cudaFree(parent_leaf);
    cudaFree(visit); cudaFree(keys); cudaFree(keys_sorted); cudaFree(vals); cudaFree(vals_sorted);
    cudaFree(cub_tmp); cudaFree(nodes); cudaFree(d_sah); cudaFree(range_first); cudaFree(range_count);
    return err;
}

}  // namespace lj
