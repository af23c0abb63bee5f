// lj_bvh_build.h -- per-thread bodies of the GPU BVH build (replaces rtcCommitScene,
// reference scene.cpp:20-27, and register_embree, shapes/triangle_mesh.inl:1-22, sphere.inl:151-162).
// Stage list: primitive boxes + scene bounds -> 63-bit Morton keys -> radix sort -> PLOC
// (parallel locally-ordered clustering, Meister & Bittner 2018: bottom-up agglomeration of the
// Morton-ordered clusters by smallest merged surface area, i.e. SAH-driven) -> top-down collapse of
// the binary tree into 8-wide nodes with quantised child boxes (Ylitie, Karras & Laine 2017) and
// octant-ordered child slots -> primitives rewritten in leaf order.  Bodies are LJ_HD so the tests
// can run the same code serially on the host; the kernels in bvh_build.cu are thin wrappers.
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

// ---- PLOC ------------------------------------------------------------------------------------
// Binary tree under construction: node ids [0, n) are the Morton-sorted primitives, ids [n, 2n-1) are
// internal nodes in creation order (the root is created last).
struct Tree2 {
    Box3 *box;    // 2n-1
    int *left;    // 2n-1 (internal nodes only)
    int *right;
    int *count;   // primitives below the node
    int n;        // number of primitives
};

// Nearest neighbour of cluster slot i among slots [i-radius, i+radius]: smallest surface area of the
// merged box; the lowest slot wins ties, which makes mutual-nearest pairs well defined.
LJ_HD int ploc_nearest(const Tree2 &t, const int *cluster, int m, int i, int radius) {
    Box3 bi = t.box[cluster[i]];
    int lo = i - radius < 0 ? 0 : i - radius;
    int hi = i + radius > m - 1 ? m - 1 : i + radius;
    float best = LJ_INF;
    int bj = -1;
    for (int j = lo; j <= hi; j++) {
        if (j == i) continue;
        float a = box_half_area(box_union(bi, t.box[cluster[j]]));
        if (a < best || bj < 0) { best = a; bj = j; }
    }
    return bj;
}

// Merge step for slot i: mutual nearest neighbours become one new internal node, kept in the lower slot.
// Returns 1 if slot i survives into the next round.
LJ_HD int ploc_merge(const Tree2 &t, const int *cluster, const int *nearest, int i, int *next_node, int *cluster_out) {
    int j = nearest[i];
    if (j >= 0 && nearest[j] == i) {
        if (i > j) return 0;
        int a = cluster[i], b = cluster[j];
        int id = LJ_ATOMIC_ADD_INT(next_node, 1);
        t.left[id] = a;
        t.right[id] = b;
        t.box[id] = box_union(t.box[a], t.box[b]);
        t.count[id] = t.count[a] + t.count[b];
        cluster_out[i] = id;
        return 1;
    }
    cluster_out[i] = cluster[i];
    return 1;
}

// ---- collapse to 8-wide compressed nodes ------------------------------------------------------
constexpr int kMaxLeafPrims = 3;      // primitives per leaf slot (unary count in 3 meta bits)
constexpr int kBvh8StackLimit = 30;   // traversal stack entries available above the sentinel (lj_bvh.h)

struct CollapseItem { int node2, node8, depth, _pad; };

struct CollapseCtx {
    Tree2 tree;
    const DevPrim *prims_sorted;  // Morton order (tree leaf ids)
    DevPrim *prims_out;           // final leaf order
    DevNode8 *nodes8;
    int *counters;                // [0] next node8, [1] next prim, [2] max depth, [3] items queued for the next level
    CollapseItem *queue_out;
    int max_leaf;                 // primitives per leaf slot, 1..kMaxLeafPrims
};

LJ_HD bool tree_is_leaf_slot(const Tree2 &t, int id, int max_leaf) { return id < t.n || t.count[id] <= max_leaf; }

// primitives below a node holding at most kMaxLeafPrims of them
LJ_HD int tree_gather_prims(const Tree2 &t, int id, int *out) {
    int stack[4], sp = 0, k = 0;
    stack[sp++] = id;
    while (sp > 0) {
        int x = stack[--sp];
        if (x < t.n) { if (k < kMaxLeafPrims) out[k++] = x; }
        else { stack[sp++] = t.right[x]; if (sp < 4) stack[sp++] = t.left[x]; }
    }
    return k;
}

LJ_HD float exp2_from_biased(int e) { return u2f((uint32_t)e << 23); }
LJ_HD uint32_t pack4(const uint32_t *v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); }

// Smallest power-of-two step 2^e (biased exponent returned) with 255 steps covering `extent`.
LJ_HD int quant_exponent(float extent) {
    if (!(extent > 0)) return 1;  // any step: every quantised coordinate is 0
    uint32_t bits = f2u(extent / 255.f);
    int e = (int)((bits >> 23) & 0xff);
    if (bits & 0x7fffffu) e += 1;  // round the step up to a power of two
    if (e < 1) e = 1;
    if (e > 254) e = 254;
    while (e < 254 && exp2_from_biased(e) * 255.f < extent) e++;
    return e;
}

// One wide node: gather up to 8 children from the binary tree (greedy: always open the child with the
// largest surface area), give them octant-ordered slots, quantise their boxes, allocate their nodes /
// primitives, and queue the internal children for the next level.
LJ_HD void collapse_node(const CollapseCtx &c, const CollapseItem &it) {
    const Tree2 &t = c.tree;
    int ch[8], nch = 0;
    if (it.node2 < t.n) {
        ch[nch++] = it.node2;  // single-primitive scene
    } else {
        ch[nch++] = t.left[it.node2];
        ch[nch++] = t.right[it.node2];
    }
    // phase 1 opens nodes that must stay internal (more than kMaxLeafPrims primitives below); phase 2
    // spends the remaining slots on opening small subtrees so that leaf boxes are as tight as possible.
    for (int phase = 0; phase < 2; phase++) {
        while (nch < 8) {
            int best = -1;
            float best_area = -1;
            for (int k = 0; k < nch; k++) {
                int id = ch[k];
                bool openable = phase == 0 ? !tree_is_leaf_slot(t, id, c.max_leaf) : id >= t.n;
                if (!openable) continue;
                float a = box_half_area(t.box[id]);
                if (a > best_area) { best_area = a; best = k; }
            }
            if (best < 0) break;
            int id = ch[best];
            ch[best] = t.left[id];
            ch[nch++] = t.right[id];
        }
    }
    Box3 nb = box_empty();
    for (int k = 0; k < nch; k++) nb = box_union(nb, t.box[ch[k]]);
    // ---- slots: slot s "points" along (s&1 ? +x : -x, s&2 ? +y : -y, s&4 ? +z : -z); a ray with
    // sign octant o visits slot (o ^ 7 ^ priority) in decreasing priority (lj_bvh.h), so the child lying
    // furthest against the slot direction of o is entered first.  Greedy assignment by projected offset.
    V3 cen = (nb.lo + nb.hi) * 0.5f;
    float cost[8][8];
    for (int k = 0; k < nch; k++) {
        V3 d = (t.box[ch[k]].lo + t.box[ch[k]].hi) * 0.5f - cen;
        for (int s = 0; s < 8; s++)
            cost[k][s] = ((s & 1) ? d.x : -d.x) + ((s & 2) ? d.y : -d.y) + ((s & 4) ? d.z : -d.z);
    }
    int slot_child[8];
    bool child_done[8];
    for (int s = 0; s < 8; s++) { slot_child[s] = -1; child_done[s] = false; }
    for (int iter = 0; iter < nch; iter++) {
        int bk = -1, bs = -1;
        float bc = -LJ_INF;
        for (int k = 0; k < nch; k++) {
            if (child_done[k]) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_child[s] >= 0) continue;
                if (bk < 0 || cost[k][s] > bc) { bc = cost[k][s]; bk = k; bs = s; }
            }
        }
        slot_child[bs] = bk;
        child_done[bk] = true;
    }
    // ---- allocation
    uint32_t imask = 0;
    int n_inner = 0, n_prims = 0;
    int leaf_prims[8][kMaxLeafPrims], leaf_count[8];
    for (int s = 0; s < 8; s++) {
        leaf_count[s] = 0;
        if (slot_child[s] < 0) continue;
        int id = ch[slot_child[s]];
        if (tree_is_leaf_slot(t, id, c.max_leaf)) {
            leaf_count[s] = tree_gather_prims(t, id, leaf_prims[s]);
            n_prims += leaf_count[s];
        } else {
            imask |= 1u << s;
            n_inner++;
        }
    }
    int child_base = n_inner ? LJ_ATOMIC_ADD_INT(&c.counters[0], n_inner) : 0;
    int prim_base = n_prims ? LJ_ATOMIC_ADD_INT(&c.counters[1], n_prims) : 0;
    int queue_base = n_inner ? LJ_ATOMIC_ADD_INT(&c.counters[3], n_inner) : 0;
    // ---- quantisation frame
    int e[3];
    float step[3];
    const float lo3[3] = {nb.lo.x, nb.lo.y, nb.lo.z}, hi3[3] = {nb.hi.x, nb.hi.y, nb.hi.z};
    for (int a = 0; a < 3; a++) { e[a] = quant_exponent(hi3[a] - lo3[a]); step[a] = exp2_from_biased(e[a]); }
    uint32_t meta[8], qlo[3][8], qhi[3][8];
    int inner_rank = 0, prim_rel = 0;
    for (int s = 0; s < 8; s++) {
        meta[s] = 0;
        for (int a = 0; a < 3; a++) { qlo[a][s] = 255; qhi[a][s] = 0; }  // empty slot: inverted box
        if (slot_child[s] < 0) continue;
        int id = ch[slot_child[s]];
        const Box3 &b = t.box[id];
        const float blo[3] = {b.lo.x, b.lo.y, b.lo.z}, bhi[3] = {b.hi.x, b.hi.y, b.hi.z};
        for (int a = 0; a < 3; a++) {
            int l = (int)floorf((blo[a] - lo3[a]) / step[a]);
            int h = (int)ceilf((bhi[a] - lo3[a]) / step[a]);
            l = clampi(l, 0, 255);
            h = clampi(h, 0, 255);
            // conservative under the decode lo3 + q * step evaluated in fp32
            while (l > 0 && lo3[a] + (float)l * step[a] > blo[a]) l--;
            while (h < 255 && lo3[a] + (float)h * step[a] < bhi[a]) h++;
            qlo[a][s] = (uint32_t)l;
            qhi[a][s] = (uint32_t)h;
        }
        if (imask & (1u << s)) {
            meta[s] = (1u << 5) | (24u + (uint32_t)s);
            CollapseItem q;
            q.node2 = id; q.node8 = child_base + inner_rank; q.depth = it.depth + 1; q._pad = 0;
            c.queue_out[queue_base + inner_rank] = q;
            inner_rank++;
        } else {
            int cnt = leaf_count[s];
            meta[s] = (((1u << cnt) - 1u) << 5) | (uint32_t)prim_rel;
            for (int k = 0; k < cnt; k++) c.prims_out[prim_base + prim_rel + k] = c.prims_sorted[leaf_prims[s][k]];
            prim_rel += cnt;
        }
    }
    DevNode8 nd;
    nd.q0 = mk4(nb.lo.x, nb.lo.y, nb.lo.z, u2f((uint32_t)e[0] | ((uint32_t)e[1] << 8) | ((uint32_t)e[2] << 16) | (imask << 24)));
    nd.q1 = mk4(u2f((uint32_t)child_base), u2f((uint32_t)prim_base), u2f(pack4(meta)), u2f(pack4(meta + 4)));
    nd.q2 = mk4(u2f(pack4(qlo[0])), u2f(pack4(qlo[0] + 4)), u2f(pack4(qlo[1])), u2f(pack4(qlo[1] + 4)));
    nd.q3 = mk4(u2f(pack4(qlo[2])), u2f(pack4(qlo[2] + 4)), u2f(pack4(qhi[0])), u2f(pack4(qhi[0] + 4)));
    nd.q4 = mk4(u2f(pack4(qhi[1])), u2f(pack4(qhi[1] + 4)), u2f(pack4(qhi[2])), u2f(pack4(qhi[2] + 4)));
    c.nodes8[it.node8] = nd;
#if defined(__CUDA_ARCH__)
    atomicMax(&c.counters[2], it.depth);
#else
    if (it.depth > c.counters[2]) c.counters[2] = it.depth;
#endif
}

}  // namespace lj
