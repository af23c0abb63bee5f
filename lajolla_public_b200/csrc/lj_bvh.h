// lj_bvh.h -- ray/primitive tests and stack traversal of the GPU-built BVH.
// Replaces rtcIntersect1 / rtcOccluded1 (reference intersection.cpp:32,83) and the sphere user
// geometry callbacks (shapes/sphere.inl:40-149).  Hit conventions follow Embree's as the reference
// consumes them: p = (1-u-v) v0 + u v1 + v v2, Ng = (v1-v0)x(v2-v0)
// unnormalised, triangle hit iff tnear <= t <= tfar, sphere hit iff tnear <= t < tfar.
#pragma once
#include "lj_scene_dev.h"

namespace lj {

constexpr int kNoHit = -1;

struct Hit {
    float t, u, v;
    int prim;  // index into DevScene::prims (BVH leaf order), kNoHit on miss
};

// Pluecker edge-function test (the formulation of Embree's ROBUST triangle intersector) in fp32.
// The edge functions of a shared edge are exact negations of each other in the two triangles, so
// the inclusive sign test is watertight along shared edges.  Traversal only needs hit / t (any t it
// reports is re-evaluated in fp64 for the winning primitive, refine_hit_t); the barycentrics of the
// winner are computed once at the end by triangle_uv from the same edge functions.
LJ_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
LJ_HD bool hit_triangle(V3 A, V3 B, V3 C, V3 o, V3 d, float tnear, float tfar, float &t) {
    V3 v0 = A - o, v1 = B - o, v2 = C - o;
    V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
    float U = dot(cross(e0, v2 + v0), d);
    float V = dot(cross(e1, v0 + v1), d);
    float W = dot(cross(e2, v1 + v2), d);
    float mn = fminf(U, fminf(V, W)), mx = fmaxf(U, fmaxf(V, W));
    if (!(mn >= 0 || mx <= 0)) return false;
    if (U + V + W == 0) return false;
    V3 Ng = cross(e0, e1);
    float den = dot(Ng, d);
    if (den == 0) return false;
    float tt = dot(v0, Ng) * fast_rcp(den);
    if (!(tt >= tnear && tt <= tfar)) return false;
    t = tt;
    return true;
}
LJ_HD void triangle_uv(V3 A, V3 B, V3 C, V3 o, V3 d, float &u, float &v) {
    V3 v0 = A - o, v1 = B - o, v2 = C - o;
    V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
    float U = dot(cross(e0, v2 + v0), d);
    float V = dot(cross(e1, v0 + v1), d);
    float W = dot(cross(e2, v1 + v2), d);
    float r = 1 / (U + V + W);
    u = fminf(U * r, 1.0f);
    v = fminf(V * r, 1.0f);
}

// Ray/sphere, nearest root in [tnear, tfar) like sphere.inl:40-106.  fp32 needs a better
// conditioned discriminant than b^2-4ac: it is taken from the perpendicular offset of the centre
// (Haines et al., "Precision improvements for ray/sphere intersection").
LJ_HD bool hit_sphere(V3 c, float r, V3 o, V3 d, float tnear, float tfar, float &t) {
    V3 f = o - c;
    float a = dot(d, d);
    if (a == 0) return false;
    float bh = -dot(f, d);
    V3 l = f + (bh / a) * d;
    float disc = r * r - dot(l, l);
    if (disc < 0) return false;
    float cc = dot(f, f) - r * r;
    float q = bh + copysignf(sqrtf(a * disc), bh);
    float t0 = cc / q, t1 = q / a;
    if (q == 0) { t0 = 0; t1 = 0; }
    if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
    float tt = -1;
    if (t0 >= tnear && t0 < tfar) tt = t0;
    if (t1 >= tnear && t1 < tfar && tt < 0) tt = t1;
    if (!(tt >= tnear && tt < tfar)) return false;
    t = tt;
    return true;
}

LJ_HD bool prim_is_sphere(const V4 &c) { return f2u(c.w) != 0; }
LJ_HD int prim_shape_id(const V4 &c) { return (int)f2u(c.y); }
LJ_HD int prim_primitive_id(const V4 &c) { return (int)f2u(c.z); }

LJ_HD bool hit_prim(const DevPrim *prims, int i, V3 o, V3 d, float tnear, float tfar, float &t) {
    V4 a = ld4(&prims[i].a);
    V4 c = ld4(&prims[i].c);
    if (prim_is_sphere(c)) return hit_sphere(xyz(a), a.w, o, d, tnear, tfar, t);
    V4 b = ld4(&prims[i].b);
    return hit_triangle(xyz(a), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, d, tnear, tfar, t);
}
// barycentrics of the final hit (spheres carry none: their (u, v) come from the hit point, lj_shapes.h)
LJ_HD void prim_uv(const V4 &a, const V4 &b, const V4 &c, V3 o, V3 d, float &u, float &v) {
    u = 0; v = 0;
    if (!prim_is_sphere(c)) triangle_uv(xyz(a), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, d, u, v);
}
LJ_HD void prim_uv(const DevPrim *prims, int i, V3 o, V3 d, float &u, float &v) {
    V4 a = ld4(&prims[i].a), b = ld4(&prims[i].b), c = ld4(&prims[i].c);
    prim_uv(a, b, c, o, d, u, v);
}

// Once the closest primitive is known its t is re-evaluated in fp64 on the same fp32 inputs: a grazing
// ray makes the fp32 quotient dot(v0,Ng)/dot(Ng,d) lose half its digits, and the reference's double
// shading code consumes that t (intersection.cpp:39-40).  One evaluation per ray, not per candidate.
LJ_HD float refine_hit_t(const V4 &a, const V4 &b, const V4 &c, V3 o, V3 d, float t32) {
    double ox = o.x, oy = o.y, oz = o.z, dx = d.x, dy = d.y, dz = d.z;
    if (prim_is_sphere(c)) {
        double fx = ox - a.x, fy = oy - a.y, fz = oz - a.z, r = a.w;
        double A = dx * dx + dy * dy + dz * dz;
        double bh = -(fx * dx + fy * dy + fz * dz);
        double C = fx * fx + fy * fy + fz * fz - r * r;
        double disc = bh * bh - A * C;
        if (!(disc >= 0) || A == 0) return t32;
        double q = bh + (bh >= 0 ? sqrt(disc) : -sqrt(disc));
        double t0 = q != 0 ? C / q : 0, t1 = q / A;
        // keep the root the fp32 test chose
        return (float)(fabs(t0 - (double)t32) <= fabs(t1 - (double)t32) ? t0 : t1);
    }
    double v0x = a.x - ox, v0y = a.y - oy, v0z = a.z - oz;
    double e0x = (double)b.z - a.x, e0y = (double)b.w - a.y, e0z = (double)c.x - a.z;   // v2 - v0
    double e1x = (double)a.x - a.w, e1y = (double)a.y - b.x, e1z = (double)a.z - b.y;   // v0 - v1
    double nx = e0y * e1z - e0z * e1y, ny = e0z * e1x - e0x * e1z, nz = e0x * e1y - e0y * e1x;
    double den = nx * dx + ny * dy + nz * dz;
    if (den == 0) return t32;
    return (float)((v0x * nx + v0y * ny + v0z * nz) / den);
}

LJ_HD float refine_hit_t(const DevPrim *prims, int prim, V3 o, V3 d, float t32) {
    V4 a = ld4(&prims[prim].a), b = ld4(&prims[prim].b), c = ld4(&prims[prim].c);
    return refine_hit_t(a, b, c, o, d, t32);
}

// Stack traversal of the 8-wide compressed BVH (DevNode8, lj_scene_dev.h), after Ylitie, Karras &
// Laine 2017.  State is a pair of "groups": G = (child node base, hit bits << 24 | imask) names the
// still-unvisited hit children of one node, Gt = (primitive base, 24 hit bits) the primitives of its
// hit leaf slots.  Children are visited highest bit first; a node's children sit in octant-ordered
// slots (lj_bvh_build.h) and the bit of slot s is 24 + (s ^ octinv), so front-to-back order comes
// from the ray's sign octant alone and no distances are sorted.  Written as resumable steps so the
// persistent kernels (wavefront.cu) can interleave traversal with fetching new rays into idle lanes.
// Edge-tie policy (same as oracle/embree_shim.cpp): among candidates with exactly equal t the lowest
// (shape id, primitive id) wins -- independent of visiting order, which in the persistent kernels depends on
// which other rays share the warp (a ray through a shared edge would otherwise shade differently run to run).
constexpr int kStack8 = 48;         // entries; node groups need <= tree depth (checked at build, <= 30)
constexpr int kTriPostponeMax = 16; // primitive groups are postponed only below this fill level

struct U2 { uint32_t x, y; };

LJ_HD int bfind32(uint32_t v) {  // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
LJ_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// per byte: 0xff where the byte's top bit is set
LJ_HD uint32_t sign_extend_s8x4(uint32_t v) {
#if defined(__CUDA_ARCH__)
    uint32_t r;  // PRMT with the sign-replicate bit set in every selector nibble (__byte_perm masks that bit off)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(0u), "r"(0x0000ba98u));
    return r;
#else
    return ((v >> 7) & 0x01010101u) * 0xffu;
#endif
}
// 1 + q / 32768 for byte i of v: the byte is dropped into the mantissa of 1.0f with one PRMT (full-rate ALU)
// instead of a shift, a mask and an I2F (quarter-rate conversion pipe, 48 of them per node otherwise).
// `one` is 0x3f800000.  The persistent kernels pass it in from a kernel PARAMETER (WaveArgs::one_bits) so that the
// compiler has to hold it in a register: PRMT takes a single immediate, and with a compile-time constant there the
// selector was materialised into a register for every byte (48 extra moves per node step, profiles/r02a_*); with the
// constant in a register the selector is the immediate and the 48 PRMTs share one register.
template <int I>
LJ_HD float byte_m(uint32_t v, uint32_t one) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(v, one, 0x7604u | (I << 4)));
#else
    return u2f(one | (((v >> (8 * I)) & 0xffu) << 8));
#endif
}
constexpr float kByteScale = 32768.0f;

struct Trav {
    V3 o, d;
    V3 idir;       // 1/d with |d| clamped away from 0
    float tnear;
    Hit hit;       // hit.t doubles as the current tfar
    uint32_t octinv4;
    U2 G, Gt;      // current node group / primitive group (y == 0: empty)
    int sp;
};
// Where the group stack of a traversal lives: in the thread's local memory (TravL, the per-thread loops and the
// walk kernels), or in a per-warp scratch area addressed by ray (k_trace_q, wavefront.cu).
struct LocalStack {
    U2 e[kStack8];
    LJ_HD void put(int k, U2 v) { e[k] = v; }
    LJ_HD U2 get(int k) const { return e[k]; }
};
struct StridedStack {  // entry k of this ray at base[k * stride]
    U2 *base;
    int stride;
    LJ_HD void put(int k, U2 v) { base[(size_t)k * stride] = v; }
    LJ_HD U2 get(int k) const { return base[(size_t)k * stride]; }
};
struct TravL : Trav { LocalStack stack; };

// Invariant between steps (kept by trav_next_group): G.y is either 0 or carries at least one hit bit, and the
// stack is only non-empty while G or Gt is, so "finished" is simply both groups being empty.
LJ_HD bool trav_done(const Trav &tr) { return (tr.G.y | tr.Gt.y) == 0; }

LJ_HD V3 trav_idir(V3 d) {
    const float tiny = 8.271806e-25f;  // 2^-80: keeps 2^e * idir finite for any node scale
    return mk3(1.0f / (fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x)),
               1.0f / (fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y)),
               1.0f / (fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z)));
}
// the same with the hardware's approximate reciprocal (1 ulp; one MUFU instead of the ~12-instruction IEEE division):
// for kernels that recompute 1/d at every node step instead of keeping it (k_trace_q)
LJ_HD V3 trav_idir_fast(V3 d) {
#if defined(__CUDA_ARCH__)
    const float tiny = 8.271806e-25f;
    float x = fabsf(d.x) > tiny ? d.x : copysignf(tiny, d.x), y = fabsf(d.y) > tiny ? d.y : copysignf(tiny, d.y);
    float z = fabsf(d.z) > tiny ? d.z : copysignf(tiny, d.z), rx, ry, rz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rz) : "f"(z));
    return mk3(rx, ry, rz);
#else
    return trav_idir(d);
#endif
}
LJ_HD void trav_init(Trav &tr, V3 o, V3 d, float tnear, float tfar) {
    tr.o = o; tr.d = d;
    tr.idir = trav_idir(d);
    tr.tnear = tnear;
    tr.hit.prim = kNoHit; tr.hit.t = tfar; tr.hit.u = tr.hit.v = 0;
    uint32_t oct = (d.x < 0 ? 1u : 0u) | (d.y < 0 ? 2u : 0u) | (d.z < 0 ? 4u : 0u);
    tr.octinv4 = (7u ^ oct) * 0x01010101u;
    tr.sp = 0;
    tr.Gt.x = 0; tr.Gt.y = 0;
    // root group: one hit bit and an empty imask, so trav_node resolves it to node 0 whatever the octant
    tr.G.x = 0;
    tr.G.y = (tnear <= tfar) ? 0x80000000u : 0u;
}

// Pop the nearest unvisited child of node group tr.G, test its 8 children: the hit internal children
// become the new tr.G (the rest of the old group is pushed), the hit leaf slots become tr.Gt.
template <class Stack>
LJ_HD void trav_node(const DevNode8 *nodes, Trav &tr, Stack &stack, const uint32_t one = 0x3f800000u) {
    U2 G = tr.G;
    int bit = bfind32(G.y);
    G.y &= ~(1u << bit);
    if (G.y & 0xff000000u) stack.put(tr.sp++, G);
    uint32_t slot = ((uint32_t)(bit - 24) ^ (tr.octinv4 & 0xffu)) & 7u;
    uint32_t rel = (uint32_t)popc32(G.y & ~(0xffffffffu << slot) & 0xffu);
    const DevNode8 *nd = nodes + (G.x + rel);
    V4 q0 = ld4(&nd->q0), q1 = ld4(&nd->q1), q2 = ld4(&nd->q2), q3 = ld4(&nd->q3), q4 = ld4(&nd->q4);
    uint32_t pk = f2u(q0.w);
    // Plane distance of quantised coordinate q: q * (2^e / d) + (p - o) / d, evaluated as m * s + (a - s) with
    // m = 1 + q / 32768 (byte_m) and s = 32768 * 2^e / d: one PRMT and one FFMA per plane.
    float sx = u2f((pk & 0xffu) << 23) * kByteScale * tr.idir.x;
    float sy = u2f(((pk >> 8) & 0xffu) << 23) * kByteScale * tr.idir.y;
    float sz = u2f(((pk >> 16) & 0xffu) << 23) * kByteScale * tr.idir.z;
    float ox = (q0.x - tr.o.x) * tr.idir.x, oy = (q0.y - tr.o.y) * tr.idir.y, oz = (q0.z - tr.o.z) * tr.idir.z;
    // The addend carries the rounding of (p - o), of 1/d and of the subtraction of s: an absolute error of
    // 2^-23 (|a| + |s|) that does not shrink with the distance itself (|s| 2^-23 is 0.4 % of one quantisation
    // step).  Near planes are pulled in and far planes pushed out by that much so the test stays conservative.
    // (3 ulp: one more than the exact-reciprocal analysis, for trav_idir_fast's 1-ulp MUFU.RCP)
    const float kSlabErr = 3.6e-7f;
    float exn = (fabsf(ox) + fabsf(sx)) * kSlabErr, eyn = (fabsf(oy) + fabsf(sy)) * kSlabErr, ezn = (fabsf(oz) + fabsf(sz)) * kSlabErr;
    float oxn = (ox - sx) - exn, oxf = (ox - sx) + exn, oyn = (oy - sy) - eyn, oyf = (oy - sy) + eyn;
    float ozn = (oz - sz) - ezn, ozf = (oz - sz) + ezn;
    const float tmax_ray = tr.hit.t;
    const bool nx = tr.d.x < 0, ny = tr.d.y < 0, nz = tr.d.z < 0;
    uint32_t hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int half = 0; half < 2; half++) {
        uint32_t meta4 = f2u(half ? q1.w : q1.z);
        uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        uint32_t bit_index4 = (meta4 ^ (tr.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        uint32_t lox = f2u(half ? q2.y : q2.x), loy = f2u(half ? q2.w : q2.z), loz = f2u(half ? q3.y : q3.x);
        uint32_t hix = f2u(half ? q3.w : q3.z), hiy = f2u(half ? q4.y : q4.x), hiz = f2u(half ? q4.w : q4.z);
        uint32_t nearx = nx ? hix : lox, farx = nx ? lox : hix;
        uint32_t neary = ny ? hiy : loy, fary = ny ? loy : hiy;
        uint32_t nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
#define LJ_CHILD(J)                                                                                         \
        {                                                                                                       \
            float t0x = fmaf(byte_m<J>(nearx, one), sx, oxn), t1x = fmaf(byte_m<J>(farx, one), sx, oxf);        \
            float t0y = fmaf(byte_m<J>(neary, one), sy, oyn), t1y = fmaf(byte_m<J>(fary, one), sy, oyf);        \
            float t0z = fmaf(byte_m<J>(nearz, one), sz, ozn), t1z = fmaf(byte_m<J>(farz, one), sz, ozf);        \
            float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tr.tnear));                                            \
            float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax_ray));                                            \
            /* conservative: widen the far side by a few ulp (Ize, "Robust BVH ray traversal") */              \
            if (tn <= tf * 1.0000008f + 1e-37f) {                                                               \
                uint32_t cb = (child_bits4 >> (8 * J)) & 0xffu, bi = (bit_index4 >> (8 * J)) & 0xffu;           \
                hitmask |= cb << bi;                                                                            \
            }                                                                                                   \
        }
        LJ_CHILD(0) LJ_CHILD(1) LJ_CHILD(2) LJ_CHILD(3)
#undef LJ_CHILD
    }
    tr.G.x = f2u(q1.x);
    tr.G.y = (hitmask & 0xff000000u) ? ((hitmask & 0xff000000u) | (pk >> 24)) : 0u;  // keep the invariant of trav_done
    tr.Gt.x = f2u(q1.y);
    tr.Gt.y = hitmask & 0x00ffffffu;
}

LJ_HD bool prim_tie_less(const DevPrim *prims, int a, int b) {
    V4 ca = ld4(&prims[a].c), cb = ld4(&prims[b].c);
    int sa = prim_shape_id(ca), sb = prim_shape_id(cb);
    return sa != sb ? sa < sb : prim_primitive_id(ca) < prim_primitive_id(cb);
}

// Test ONE primitive of group tr.Gt (highest bit first).  ANY: returns true at the first hit.
template <bool ANY>
LJ_HD bool trav_prim(const DevPrim *prims, Trav &tr) {
    int k = bfind32(tr.Gt.y);
    tr.Gt.y &= ~(1u << k);
    int idx = (int)tr.Gt.x + k;
    float t;
    if (hit_prim(prims, idx, tr.o, tr.d, tr.tnear, tr.hit.t, t)) {
        if (ANY) { tr.hit.prim = idx; tr.hit.t = t; return true; }
        if (t < tr.hit.t || tr.hit.prim == kNoHit) { tr.hit.t = t; tr.hit.prim = idx; }
        else if (t == tr.hit.t && prim_tie_less(prims, idx, tr.hit.prim)) tr.hit.prim = idx;
    }
    return false;
}

// After a step: if both groups are used up, pop the next one.  A popped entry is either a node group or a
// postponed primitive group (no hit bits in the top byte); the latter lands in tr.Gt.
template <class Stack>
LJ_HD void trav_next_group(Trav &tr, const Stack &stack) {
    if ((tr.G.y | tr.Gt.y) == 0 && tr.sp > 0) {
        U2 e = stack.get(--tr.sp);
        if (e.y & 0xff000000u) tr.G = e; else tr.Gt = e;
    }
}

LJ_HD void trav_terminate(Trav &tr) { tr.G.y = 0; tr.Gt.y = 0; tr.sp = 0; }

// Closest hit: the winning primitive's t is refined in fp64 (see refine_hit_t).
LJ_HD void trav_finish_closest(const DevPrim *prims, Trav &tr) {
    if (tr.hit.prim == kNoHit) return;
    // (the winner's three records are fetched once for both: the two helpers loaded them separately and the compiler kept
    //  both sets of loads -- 3 % of the closest-hit kernel's global-load requests, profiles/r02k per-line counts)
    const V4 a = ld4(&prims[tr.hit.prim].a), b = ld4(&prims[tr.hit.prim].b), c = ld4(&prims[tr.hit.prim].c);
    prim_uv(a, b, c, tr.o, tr.d, tr.hit.u, tr.hit.v);
    tr.hit.t = refine_hit_t(a, b, c, tr.o, tr.d, tr.hit.t);
}

// Plain single-ray loop (query seam S2 and shading-side helpers).
template <bool ANY>
LJ_HD bool trace8(const DevNode8 *nodes, const DevPrim *prims, V3 o, V3 d, float tnear, float tfar, Hit &hit) {
    TravL tr;
    trav_init(tr, o, d, tnear, tfar);
    while (!trav_done(tr)) {
        if (tr.Gt.y == 0) trav_node(nodes, tr, tr.stack);
        while (tr.Gt.y != 0) {
            if (trav_prim<ANY>(prims, tr)) { trav_terminate(tr); break; }
        }
        trav_next_group(tr, tr.stack);
    }
    if (!ANY) trav_finish_closest(prims, tr);
    hit = tr.hit;
    return hit.prim != kNoHit;
}

}  // namespace lj
