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
// the inclusive sign test is watertight along shared edges.
LJ_HD bool hit_triangle(V3 A, V3 B, V3 C, V3 o, V3 d, float tnear, float tfar, float &t, float &u, float &v) {
    V3 v0 = A - o, v1 = B - o, v2 = C - o;
    V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
    float U = dot(cross(e0, v2 + v0), d);
    float V = dot(cross(e1, v0 + v1), d);
    float W = dot(cross(e2, v1 + v2), d);
    float mn = fminf(U, fminf(V, W)), mx = fmaxf(U, fmaxf(V, W));
    if (!(mn >= 0 || mx <= 0)) return false;
    float UVW = U + V + W;
    if (UVW == 0) return false;
    V3 Ng = cross(e0, e1);
    float den = dot(Ng, d);
    if (den == 0) return false;
    float tt = dot(v0, Ng) / den;
    if (!(tt >= tnear && tt <= tfar)) return false;
    t = tt;
    float r = 1 / UVW;
    u = fminf(U * r, 1.0f);
    v = fminf(V * r, 1.0f);
    return true;
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

LJ_HD bool hit_prim(const DevPrim *prims, int i, V3 o, V3 d, float tnear, float tfar, float &t, float &u, float &v) {
    V4 a = ld4(&prims[i].a);
    V4 c = ld4(&prims[i].c);
    if (prim_is_sphere(c)) {
        u = 0; v = 0;
        return hit_sphere(xyz(a), a.w, o, d, tnear, tfar, t);
    }
    V4 b = ld4(&prims[i].b);
    return hit_triangle(xyz(a), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, d, tnear, tfar, t, u, v);
}

// Once the closest primitive is known its t is re-evaluated in fp64 on the same fp32 inputs: a grazing
// ray makes the fp32 quotient dot(v0,Ng)/dot(Ng,d) lose half its digits, and the reference's double
// shading code consumes that t (intersection.cpp:39-40).  One evaluation per ray, not per candidate.
LJ_HD float refine_hit_t(const DevPrim *prims, int prim, V3 o, V3 d, float t32) {
    V4 a = ld4(&prims[prim].a), b = ld4(&prims[prim].b), c = ld4(&prims[prim].c);
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

// Stack traversal of the binary BVH (DevNode2), written as resumable steps so the persistent
// kernels (wavefront.cu) can interleave traversal with fetching new rays into idle lanes.
// Edge-tie policy: a later candidate replaces the current hit only if strictly
// nearer, so among exactly equal t the first one visited wins.
constexpr int kStackSize = 64;
constexpr int kSentinel = 0x7fffffff;

struct Trav {
    V3 o, d, inv;
    float tnear;
    Hit hit;   // hit.t doubles as the current tfar
    int node;  // >= 0 inner node, < 0 leaf, kSentinel = finished
    int leaf;  // postponed leaf (speculative while-while traversal), 0 = none
    int sp;
    int stack[kStackSize];
};

LJ_HD void trav_init(Trav &tr, V3 o, V3 d, float tnear, float tfar) {
    tr.o = o; tr.d = d;
    tr.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    tr.tnear = tnear;
    tr.hit.prim = kNoHit; tr.hit.t = tfar; tr.hit.u = tr.hit.v = 0;
    tr.sp = 0;
    tr.leaf = 0;
    tr.stack[tr.sp++] = kSentinel;
    tr.node = (tnear <= tfar) ? 0 : kSentinel;
}

// One inner-node step: slab tests of both children, descend into the nearer hit child.
LJ_HD void trav_inner(const DevNode2 *nodes, Trav &tr) {
    const int node = tr.node;
    const V3 o = tr.o, inv = tr.inv;
    V4 n0 = ld4(&nodes[node].n0), n1 = ld4(&nodes[node].n1);
    V4 n2 = ld4(&nodes[node].n2), n3 = ld4(&nodes[node].n3);
    float c0lox = (n0.x - o.x) * inv.x, c0hix = (n0.y - o.x) * inv.x;
    float c0loy = (n0.z - o.y) * inv.y, c0hiy = (n0.w - o.y) * inv.y;
    float c0loz = (n2.x - o.z) * inv.z, c0hiz = (n2.y - o.z) * inv.z;
    float c1lox = (n1.x - o.x) * inv.x, c1hix = (n1.y - o.x) * inv.x;
    float c1loy = (n1.z - o.y) * inv.y, c1hiy = (n1.w - o.y) * inv.y;
    float c1loz = (n2.z - o.z) * inv.z, c1hiz = (n2.w - o.z) * inv.z;
    float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tr.tnear));
    float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tr.hit.t));
    float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tr.tnear));
    float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tr.hit.t));
    // conservative: widen the far side by 2 ulp (Ize, "Robust BVH ray traversal")
    bool h0 = t0n <= t0f * 1.0000004f;
    bool h1 = t1n <= t1f * 1.0000004f;
    int c0 = (int)f2u(n3.x), c1 = (int)f2u(n3.y);
    if (h0 && h1) {
        bool swap = t1n < t0n;
        tr.node = swap ? c1 : c0;
        if (tr.sp < kStackSize) tr.stack[tr.sp++] = swap ? c0 : c1;
    } else if (h0) {
        tr.node = c0;
    } else if (h1) {
        tr.node = c1;
    } else {
        tr.node = tr.stack[--tr.sp];
    }
}

// Test the primitives of leaf reference `leaf`.  ANY: returns true at the first hit.
template <bool ANY>
LJ_HD bool trav_test_leaf(const DevPrim *prims, Trav &tr, int leaf) {
    int v = ~leaf;
    int first = v >> 3, count = (v & 7) + 1;
    for (int i = 0; i < count; i++) {
        float t, uu, vv;
        if (hit_prim(prims, first + i, tr.o, tr.d, tr.tnear, tr.hit.t, t, uu, vv)) {
            if (ANY) { tr.hit.prim = first + i; tr.hit.t = t; return true; }
            if (t < tr.hit.t || tr.hit.prim == kNoHit) {
                tr.hit.t = t; tr.hit.u = uu; tr.hit.v = vv; tr.hit.prim = first + i;
            }
        }
    }
    return false;
}
// Leaf step of the plain loop: test tr.node's primitives, then pop.
template <bool ANY>
LJ_HD bool trav_leaf(const DevPrim *prims, Trav &tr) {
    if (trav_test_leaf<ANY>(prims, tr, tr.node)) { tr.node = kSentinel; return true; }
    tr.node = tr.stack[--tr.sp];
    return false;
}

// Closest hit: the winning primitive's t is refined in fp64 (see refine_hit_t).
LJ_HD void trav_finish_closest(const DevPrim *prims, Trav &tr) {
    if (tr.hit.prim != kNoHit) tr.hit.t = refine_hit_t(prims, tr.hit.prim, tr.o, tr.d, tr.hit.t);
}

template <bool ANY>
LJ_HD bool trace2(const DevNode2 *nodes, const DevPrim *prims, V3 o, V3 d, float tnear, float tfar, Hit &hit) {
    Trav tr;
    trav_init(tr, o, d, tnear, tfar);
    while (tr.node != kSentinel) {
        if (tr.node >= 0) trav_inner(nodes, tr);
        else if (trav_leaf<ANY>(prims, tr)) break;
    }
    if (!ANY) trav_finish_closest(prims, tr);
    hit = tr.hit;
    return hit.prim != kNoHit;
}

}  // namespace lj
