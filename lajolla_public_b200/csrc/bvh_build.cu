// bvh_build.cu -- GPU BVH construction kernels (sm_100a).  See lj_bvh_build.h for the stages.
// The one library primitive used is cub::DeviceRadixSort for the Morton keys (one-off build step,
// not on the per-sample path).
#include "bvh_build.cuh"

#include <stdlib.h>

#include <utility>

#if defined(LJ_HOSTSIM)
#include <numeric>
#include <vector>
#else
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
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

__global__ void k_prim_boxes(const LJ_GRID_CONSTANT DevScene sc, const int *prim_shape, const int *prim_local, const float *ref_box, int n,
                             DevPrim *prims_unsorted, Box3 *boxes, int *scene_bounds /*6 ordered ints*/) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Box3 b = box_empty();
    if (i < n) {
        prims_unsorted[i] = make_prim(sc, prim_shape[i], prim_local[i], b);
        if (ref_box) {  // a split reference: the clipped box (inside the primitive's own box by construction)
            b.lo = mk3(ref_box[6 * i], ref_box[6 * i + 1], ref_box[6 * i + 2]);
            b.hi = mk3(ref_box[6 * i + 3], ref_box[6 * i + 4], ref_box[6 * i + 5]);
        }
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

// Morton order: primitive records and their boxes become tree leaves 0..n-1; every leaf is its own cluster.
__global__ void k_gather(const uint32_t *order, int n, const DevPrim *prims_unsorted, const Box3 *boxes, DevPrim *prims_sorted,
                         Tree2 tree, int *cluster) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t src = order[i];
    prims_sorted[i] = prims_unsorted[src];
    tree.box[i] = boxes[src];
    tree.count[i] = 1;
    cluster[i] = i;
}

__global__ void k_ploc_nearest(Tree2 tree, const int *cluster, int m, int radius, int *nearest) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    nearest[i] = ploc_nearest(tree, cluster, m, i, radius);
}

__global__ void k_ploc_merge(Tree2 tree, const int *cluster, const int *nearest, int m, int *next_node, int *cluster_tmp, int *keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    keep[i] = ploc_merge(tree, cluster, nearest, i, next_node, cluster_tmp);
}

__global__ void k_ploc_compact(const int *cluster_tmp, const int *keep, const int *offset, int m, int *cluster_out, int *m_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (keep[i]) cluster_out[offset[i]] = cluster_tmp[i];
    if (i == m - 1) *m_out = offset[i] + keep[i];
}

__global__ void k_collapse(CollapseCtx ctx, const CollapseItem *queue_in, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    collapse_node(ctx, queue_in[i]);
}

// SAH cost of the binary tree: sum over internal nodes of A(node)/A(root) * 1.2 + leaves * 1.0
__global__ void k_sah(Tree2 tree, int root, double *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0;
    float ra = box_half_area(tree.box[root]);
    if (i < 2 * tree.n - 1 && ra > 0) c = (i < tree.n ? 1.0 : 1.2) * box_half_area(tree.box[i]) / ra;
    c = lj_warp_sum(c);
    if (LJ_LANE() == 0 && c != 0) atomicAdd(out, c);
}

template <typename T>
cudaError_t dalloc(T **p, size_t n) { return lj_dev_alloc((void **)p, (n ? n : 1) * sizeof(T)); }

}  // namespace

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = e_; goto done; } } while (0)

static cudaError_t exclusive_scan(void *tmp, size_t &tmp_bytes, const int *in, int *out, int m, cudaStream_t stream) {
#if defined(LJ_HOSTSIM)
    (void)tmp; (void)stream;
    tmp_bytes = 1;
    if (in) { int acc = 0; for (int i = 0; i < m; i++) { out[i] = acc; acc += in[i]; } }
    return cudaSuccess;
#else
    return cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, m, stream);
#endif
}

cudaError_t build_bvh8(const DevScene &sc, const int *d_prim_shape, const int *d_prim_local, const float *d_ref_box, int n,
                       cudaStream_t stream, BvhResult *out) {
    // search radius of the PLOC neighbour search: LJ_PLOC_RADIUS in [1, 64] (default 32)
    int radius = 32;
    if (const char *e = getenv("LJ_PLOC_RADIUS")) radius = atoi(e);
    radius = radius < 1 ? 1 : (radius > 64 ? 64 : radius);
    cudaError_t err = cudaSuccess;
    DevPrim *prims_unsorted = nullptr, *prims_sorted = nullptr, *prims = nullptr;
    Box3 *boxes = nullptr;
    Tree2 tree = {nullptr, nullptr, nullptr, nullptr, n};
    int *scene_bounds = nullptr, *cluster_a = nullptr, *cluster_b = nullptr, *cluster_tmp = nullptr, *nearest = nullptr;
    int *keep = nullptr, *offset = nullptr, *d_ints = nullptr;  // d_ints: [0] next binary node, [1] cluster count, [2..5] collapse counters
    uint64_t *keys = nullptr, *keys_sorted = nullptr;
    uint32_t *vals = nullptr, *vals_sorted = nullptr;
    CollapseItem *queue_a = nullptr, *queue_b = nullptr;
    void *cub_tmp = nullptr, *scan_tmp = nullptr;
    size_t cub_bytes = 0, scan_bytes = 0;
    DevNode8 *nodes = nullptr;
    double *d_sah = nullptr;
    const int T = 256;
    const int nb = (n + T - 1) / T;
    const int n_tree = 2 * n - 1;
    const int max_nodes8 = n > 1 ? n - 1 : 1;
    int h_bounds[6], h_ints[6];
    int init_bounds[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    int m = n, root2 = 0, launches = 0, rounds = 0, levels = 0;

    CK(dalloc(&prims_unsorted, n)); CK(dalloc(&prims_sorted, n)); CK(dalloc(&prims, n));
    CK(dalloc(&boxes, n)); CK(dalloc(&tree.box, n_tree)); CK(dalloc(&tree.left, n_tree)); CK(dalloc(&tree.right, n_tree));
    CK(dalloc(&tree.count, n_tree));
    CK(dalloc(&scene_bounds, 6)); CK(dalloc(&cluster_a, n)); CK(dalloc(&cluster_b, n)); CK(dalloc(&cluster_tmp, n));
    CK(dalloc(&nearest, n)); CK(dalloc(&keep, n)); CK(dalloc(&offset, n)); CK(dalloc(&d_ints, 6));
    CK(dalloc(&keys, n)); CK(dalloc(&keys_sorted, n)); CK(dalloc(&vals, n)); CK(dalloc(&vals_sorted, n));
    CK(dalloc(&queue_a, n)); CK(dalloc(&queue_b, n));
    CK(dalloc(&nodes, max_nodes8)); CK(dalloc(&d_sah, 1));
    CK(cudaMemcpyAsync(scene_bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(d_sah, 0, sizeof(double), stream));
    CK(cudaMemsetAsync(nodes, 0, sizeof(DevNode8) * max_nodes8, stream));

    LJ_LAUNCH(k_prim_boxes, nb, T, stream, sc, d_prim_shape, d_prim_local, d_ref_box, n, prims_unsorted, boxes, scene_bounds);
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
    CK(lj_dev_alloc(&cub_tmp, cub_bytes ? cub_bytes : 1));
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys, keys_sorted, vals, vals_sorted, n, 0, 63, stream));
#endif
    LJ_LAUNCH(k_gather, nb, T, stream, vals_sorted, n, prims_unsorted, boxes, prims_sorted, tree, cluster_a);
    launches = 5;

    // ---- PLOC rounds: nearest neighbour -> merge mutual pairs -> order-preserving compaction
    h_ints[0] = n; h_ints[1] = n;
    CK(cudaMemcpyAsync(d_ints, h_ints, 2 * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(exclusive_scan(nullptr, scan_bytes, nullptr, offset, n, stream));
    CK(lj_dev_alloc(&scan_tmp, scan_bytes ? scan_bytes : 1));
    while (m > 1) {
        int g = (m + T - 1) / T;
        LJ_LAUNCH(k_ploc_nearest, g, T, stream, tree, cluster_a, m, radius, nearest);
        LJ_LAUNCH(k_ploc_merge, g, T, stream, tree, cluster_a, nearest, m, d_ints + 0, cluster_tmp, keep);
        CK(exclusive_scan(scan_tmp, scan_bytes, keep, offset, m, stream));
        LJ_LAUNCH(k_ploc_compact, g, T, stream, cluster_tmp, keep, offset, m, cluster_b, d_ints + 1);
        CK(cudaMemcpyAsync(h_ints, d_ints, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (h_ints[1] >= m || h_ints[1] < 1) { err = cudaErrorUnknown; goto done; }  // every round merges at least one pair
        m = h_ints[1];
        std::swap(cluster_a, cluster_b);
        launches += 4;
        rounds++;
    }
    root2 = n > 1 ? 2 * n - 2 : 0;
    LJ_LAUNCH(k_sah, (n_tree + T - 1) / T, T, stream, tree, root2, d_sah);

    // ---- collapse, one level of wide nodes per launch
    {
        CollapseCtx ctx;
        ctx.tree = tree;
        ctx.prims_sorted = prims_sorted;
        ctx.prims_out = prims;
        ctx.nodes8 = nodes;
        ctx.counters = d_ints + 2;
        ctx.max_leaf = kMaxLeafPrims;  // LJ_BVH_MAX_LEAF in [1, 3] for tuning runs
        if (const char *e = getenv("LJ_BVH_MAX_LEAF")) ctx.max_leaf = atoi(e) < 1 ? 1 : (atoi(e) > kMaxLeafPrims ? kMaxLeafPrims : atoi(e));
        CollapseItem first = {root2, 0, 1, 0};
        int init[4] = {1, 0, 0, 0};
        CK(cudaMemcpyAsync(queue_a, &first, sizeof(first), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_ints + 2, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        int count = 1;
        while (count > 0) {
            ctx.queue_out = queue_b;
            CK(cudaMemsetAsync(d_ints + 5, 0, sizeof(int), stream));
            LJ_LAUNCH(k_collapse, (count + 63) / 64, 64, stream, ctx, queue_a, count);
            CK(cudaMemcpyAsync(h_ints + 2, d_ints + 2, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            count = h_ints[5];
            std::swap(queue_a, queue_b);
            launches++;
            levels++;
            if (levels > 4096) { err = cudaErrorUnknown; goto done; }
        }
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
    out->num_nodes = h_ints[2];
    out->num_prims_placed = h_ints[3];
    out->depth = h_ints[4];
    out->ploc_rounds = rounds;
    out->launches = launches + 1;
    nodes = nullptr;
    prims = nullptr;
done:
    lj_dev_free(prims_unsorted); lj_dev_free(prims_sorted); lj_dev_free(prims); lj_dev_free(boxes);
    lj_dev_free(tree.box); lj_dev_free(tree.left); lj_dev_free(tree.right); lj_dev_free(tree.count);
    lj_dev_free(scene_bounds); lj_dev_free(cluster_a); lj_dev_free(cluster_b); lj_dev_free(cluster_tmp); lj_dev_free(nearest);
    lj_dev_free(keep); lj_dev_free(offset); lj_dev_free(d_ints);
    lj_dev_free(keys); lj_dev_free(keys_sorted); lj_dev_free(vals); lj_dev_free(vals_sorted);
    lj_dev_free(queue_a); lj_dev_free(queue_b);
    lj_dev_free(cub_tmp); lj_dev_free(scan_tmp); lj_dev_free(nodes); lj_dev_free(d_sah);
    return err;
}

}  // namespace lj
