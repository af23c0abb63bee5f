// lj_bvh_build.h -- per-thread bodies of the GPU BVH build (replaces rtcCommitScene,
// reference scene.cpp:20-27, and register_embree, shapes/triangle_mesh.inl:1-22, sphere.inl:151-162).
// Stage list: primitive boxes + scene bounds -> 63-bit Morton keys -> radix sort -> Karras 2012
// hierarchy -> bottom-up refit -> emit DevNode2/DevPrim in leaf order.  Bodies are LJ_HD so the
// tests can run the same code serially on the host; the kernels in bvh_build.cu are thin wrappers.
#pragma once
#include "lj_scene_dev.h"

#if defined(__CUDA_ARCH__)
#define LJ_ATOMIC_ADD_INT(p, v) atomicAdd((p), (v))
#define LJ_THREADFENCE() __threadfence()
#else
#define LJ_ATOMIC_ADD_INT(p, v) lj::host_fetch_add((p), (v))
#define LJ_THREADFENCE() ((void)0)
#endif

namespace lj {

inline int host_fetch_add(int *p, int v) { int o = *p; *p = o + v; return o; }

struct Box3 { V3 lo, hi; };
LJ_HD Box3 box_empty() { Box3 b; b.lo = mk3(LJ_INF); b.hi = mk3(-LJ_INF); return b; }
LJ_HD Box3 box_union(const Box3 &a, const Box3 &b) { Box3 r; r.lo = min3v(a.lo, b.lo); r.hi = max3v(a.hi, b.hi); return r; }
LJ_HD float box_half_area(const Box3 &b) {
    V3 d = b.hi - b.lo;
    if (d.x < 0 || d.y < 0 || d.z < 0) return 0;
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

struct BuildInputs {  // unsorted primitive list, one entry per triangle / sphere
    const int *prim_shape;  // shape id
    const int *prim_local;  // triangle index inside the mesh (0 for spheres)
    int n;
};

// Unsorted primitive record + box.  Vertex positions are exactly the fp32 values the reference
// hands to Embree (triangle_mesh.inl:11-14 casts to float).
LJ_HD DevPrim make_prim(const DevScene &sc, int shape_id, int local, Box3 &box) {
    const DevShape &sh = sc.shapes[shape_id];
    DevPrim p;
    if (sh.type == 0) {
        p.a = mk4(sh.cx, sh.cy, sh.cz, sh.radius);
        p.b = mk4(0, 0, 0, 0);
        p.c = mk4(0, u2f((uint32_t)shape_id), u2f(0u), u2f(1u));
        box.lo = mk3(sh.cx - sh.radius, sh.cy - sh.radius, sh.cz - sh.radius);  // sphere.inl:1-10
        box.hi = mk3(sh.cx + sh.radius, sh.cy + sh.radius, sh.cz + sh.radius);
        return p;
    }
    const int *idx = sc.indices + 3 * (sh.tri_offset + local);
    const float *P = sc.positions;
    V3 A = mk3(P[3 * idx[0]], P[3 * idx[0] + 1], P[3 * idx[0] + 2]);
    V3 B = mk3(P[3 * idx[1]], P[3 * idx[1] + 1], P[3 * idx[1] + 2]);
    V3 C = mk3(P[3 * idx[2]], P[3 * idx[2] + 1], P[3 * idx[2] + 2]);
    p.a = mk4(A.x, A.y, A.z, B.x);
    p.b = mk4(B.y, B.z, C.x, C.y);
    p.c = mk4(C.z, u2f((uint32_t)shape_id), u2f((uint32_t)local), u2f(0u));
    box.lo = min3v(A, min3v(B, C));
    box.hi = max3v(A, max3v(B, C));
    return p;
}

LJ_HD uint64_t expand_bits_21(uint32_t v) {  // spread the low 21 bits to every third bit
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
LJ_HD uint64_t morton63(V3 c, const Box3 &scene) {
    V3 ext = scene.hi - scene.lo;
    float fx = ext.x > 0 ? (c.x - scene.lo.x) / ext.x : 0.f;
    float fy = ext.y > 0 ? (c.y - scene.lo.y) / ext.y : 0.f;
    float fz = ext.z > 0 ? (c.z - scene.lo.z) / ext.z : 0.f;
    const float s = 2097152.f;  // 2^21
    uint32_t ix = (uint32_t)fminf(fmaxf(fx * s, 0.f), s - 1);
    uint32_t iy = (uint32_t)fminf(fmaxf(fy * s, 0.f), s - 1);
    uint32_t iz = (uint32_t)fminf(fmaxf(fz * s, 0.f), s - 1);
    return (expand_bits_21(ix) << 2) | (expand_bits_21(iy) << 1) | expand_bits_21(iz);
}

LJ_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x == 0 ? 64 : __builtin_clzll(x);
#endif
}
LJ_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x == 0 ? 32 : __builtin_clz(x);
#endif
}

// Karras 2012: common-prefix length with the index as tie breaker for duplicate keys.
LJ_HD int karras_delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz64(a ^ b);
}

// Internal node i of n-1: children and parent links.  child encoding here: >=0 internal, <0 leaf ~k
// (k = position in sorted order).
LJ_HD void karras_node(const uint64_t *keys, int n, int i, int *left, int *right, int *parent_internal, int *parent_leaf,
                       int *range_first, int *range_count) {
    int d = karras_delta(keys, n, i, i + 1) - karras_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    int dmin = karras_delta(keys, n, i, i - d);
    int lmax = 2;
    while (karras_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (karras_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = karras_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    int lc = (lo == gamma) ? ~gamma : gamma;
    int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    range_first[i] = lo;          // the subtree of node i covers sorted primitives [lo, hi]
    range_count[i] = hi - lo + 1;
    if (lc >= 0) parent_internal[lc] = i; else parent_leaf[~lc] = i;
    if (rc >= 0) parent_internal[rc] = i; else parent_leaf[~rc] = i;
}

// Bottom-up refit from sorted leaf k: the second thread to reach a node computes its box.
LJ_HD void refit_from_leaf(int k, const Box3 *leaf_box, Box3 *node_box, const int *left, const int *right,
                           const int *parent_internal, const int *parent_leaf, int *visit) {
    int node = parent_leaf[k];
    while (node >= 0) {
        LJ_THREADFENCE();
        if (LJ_ATOMIC_ADD_INT(&visit[node], 1) == 0) return;
        LJ_THREADFENCE();
        int lc = left[node], rc = right[node];
        Box3 lb = lc >= 0 ? node_box[lc] : leaf_box[~lc];
        Box3 rb = rc >= 0 ? node_box[rc] : leaf_box[~rc];
        node_box[node] = box_union(lb, rb);
        node = parent_internal[node];
    }
}

// Emit the traversal node of internal node i (children's boxes stored in the parent).  A child
// subtree holding at most max_leaf primitives is collapsed into one leaf: Karras subtrees cover a
// contiguous range of the sorted primitive array, so the leaf is just (first, count).
LJ_HD int leaf_ref(int first, int count) { return ~((first << 3) | (count - 1)); }
LJ_HD DevNode2 emit_node2(int i, const Box3 *leaf_box, const Box3 *node_box, const int *left, const int *right,
                          const int *range_first, const int *range_count, int max_leaf) {
    int lc = left[i], rc = right[i];
    Box3 lb = lc >= 0 ? node_box[lc] : leaf_box[~lc];
    Box3 rb = rc >= 0 ? node_box[rc] : leaf_box[~rc];
    int c0 = lc >= 0 ? (range_count[lc] <= max_leaf ? leaf_ref(range_first[lc], range_count[lc]) : lc) : leaf_ref(~lc, 1);
    int c1 = rc >= 0 ? (range_count[rc] <= max_leaf ? leaf_ref(range_first[rc], range_count[rc]) : rc) : leaf_ref(~rc, 1);
    DevNode2 nd;
    nd.n0 = mk4(lb.lo.x, lb.hi.x, lb.lo.y, lb.hi.y);
    nd.n1 = mk4(rb.lo.x, rb.hi.x, rb.lo.y, rb.hi.y);
    nd.n2 = mk4(lb.lo.z, lb.hi.z, rb.lo.z, rb.hi.z);
    nd.n3 = mk4(u2f((uint32_t)c0), u2f((uint32_t)c1), 0.f, 0.f);
    return nd;
}

}  // namespace lj
