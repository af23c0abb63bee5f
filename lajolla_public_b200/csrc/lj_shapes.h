// lj_shapes.h -- hit -> PathVertex assembly, shape sampling.  fp32 restatement of the reference's
// intersection.cpp:37-62, shapes/triangle_mesh.inl:24-169, shapes/sphere.inl:164-268,
// table_dist.cpp:27-38.  Quirks reproduced on purpose are marked TRAP.
#pragma once
#include "lj_bvh.h"

namespace lj {

struct Vertex {  // intersection.h:15-35
    V3 position, geometric_normal;
    Frame shading_frame;
    V2 st, uv;
    float uv_screen_size, mean_curvature, ray_radius;
    int shape_id, primitive_id, material_id, interior_medium_id, exterior_medium_id;
};

struct PointAndNormal { V3 position, normal; };

LJ_HD V3 ld3(const float *p, int i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
LJ_HD V2 ld2(const float *p, int i) { return mk2(p[2 * i], p[2 * i + 1]); }

// table_dist.cpp:27-33: upper_bound(cdf, cdf+n+1, u) - 1, clamped to [0, n-1].
LJ_HD int sample_table_1d(const float *cdf, int n, float u) {
    int lo = 0, hi = n + 1;  // first index with cdf[idx] > u
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return clampi(lo - 1, 0, n - 1);
}

// Sphere (u,v) as the intersect callback stores them (sphere.inl:91-97).
LJ_HD V2 sphere_st(V3 ng_unnormalised, float radius) {
    V3 c = ng_unnormalised / radius;
    float elevation = acosf(clampf(c.y, -1.f, 1.f));
    float azimuth = atan2f(c.z, c.x);
    return mk2(azimuth / kTwoPi, elevation / kPi);
}

struct ShadingInfo { V2 uv; Frame frame; float mean_curvature, inv_uv_size; };

// triangle_mesh.inl:77-169
LJ_HD ShadingInfo shading_info_mesh(const DevScene &sc, const DevShape &sh, int prim_id, V2 st, V3 ng) {
    const int *idx = sc.indices + 3 * (sh.tri_offset + prim_id);
    int i0 = idx[0], i1 = idx[1], i2 = idx[2];
    V2 uv0, uv1, uv2;
    if (sh.has_uvs) {
        uv0 = ld2(sc.uvs, i0); uv1 = ld2(sc.uvs, i1); uv2 = ld2(sc.uvs, i2);
    } else {
        uv0 = mk2(0, 0); uv1 = mk2(1, 0); uv2 = mk2(1, 1);
    }
    float b0 = 1 - st.x - st.y;
    V2 uv = b0 * uv0 + st.x * uv1 + st.y * uv2;
    V3 p0 = ld3(sc.positions, i0), p1 = ld3(sc.positions, i1), p2 = ld3(sc.positions, i2);
    V2 duvds = uv2 - uv0, duvdt = uv2 - uv1;
    float det = duvds.x * duvdt.y - duvdt.x * duvds.y;
    float dsdu = duvdt.y / det, dtdu = -duvds.y / det;
    float dsdv = duvdt.x / det, dtdv = -duvds.x / det;
    V3 dpdu, dpdv;
    if (fabsf(det) > 1e-8f) {
        V3 dpds = p2 - p0, dpdt = p2 - p1;  // TRAP: the reference's pairing, kept verbatim (:124-125)
        dpdu = dpds * dsdu + dpdt * dtdu;
        dpdv = dpds * dsdv + dpdt * dtdv;
    } else {
        coordinate_system(ng, dpdu, dpdv);
    }
    V3 sn = ng;
    float mean_curvature = 0;
    V3 tangent, bitangent;
    if (sh.has_normals) {
        V3 n0 = ld3(sc.normals, i0), n1 = ld3(sc.normals, i1), n2 = ld3(sc.normals, i2);
        sn = normalize(b0 * n0 + st.x * n1 + st.y * n2);
        tangent = normalize(dpdu - sn * dot(sn, dpdu));
        V3 dnds = n2 - n0, dndt = n2 - n1;
        V3 dndu = dnds * dsdu + dndt * dtdu;
        V3 dndv = dnds * dsdv + dndt * dtdv;
        bitangent = normalize(cross(sn, tangent));
        mean_curvature = (dot(dndu, tangent) + dot(dndv, bitangent)) / 2;
    } else {
        tangent = normalize(dpdu - sn * dot(sn, dpdu));
        bitangent = normalize(cross(sn, tangent));
    }
    ShadingInfo si;
    si.uv = uv;
    si.frame = make_frame(tangent, bitangent, sn);
    si.mean_curvature = mean_curvature;
    si.inv_uv_size = fmaxf(length(dpdu), length(dpdv));
    return si;
}

// sphere.inl:243-268.  TRAP: st (normalised to [-.5,.5]x[0,1]) is used as radians, as upstream.
LJ_HD ShadingInfo shading_info_sphere(const DevShape &sh, V2 st, V3 ng) {
    float r = sh.radius;
    float su = sinf(st.x), cu = cosf(st.x), sv = sinf(st.y), cv = cosf(st.y);
    V3 dpdu = mk3(-r * su * sv, r * cu * sv, 0);
    V3 dpdv = mk3(r * cu * cv, r * su * cv, -r * sv);
    V3 tangent = normalize(dpdu - ng * dot(ng, dpdu));
    ShadingInfo si;
    si.uv = st;
    si.frame = make_frame(tangent, normalize(cross(ng, tangent)), ng);
    si.mean_curvature = 1 / r;
    si.inv_uv_size = (length(dpdu) + length(dpdv)) / 2;
    return si;
}

// intersection.cpp:37-62: assemble the PathVertex of a hit.  (radius, spread) = ray differential.
LJ_HD Vertex make_vertex(const DevScene &sc, V3 org, V3 dir, const Hit &hit, float rd_radius, float rd_spread) {
    Vertex vx;
    V4 pa = ld4(&sc.prims[hit.prim].a);
    V4 pc = ld4(&sc.prims[hit.prim].c);
    vx.position = org + dir * hit.t;
    vx.shape_id = prim_shape_id(pc);
    vx.primitive_id = prim_primitive_id(pc);
    const DevShape &sh = sc.shapes[vx.shape_id];
    vx.material_id = sh.material_id;
    vx.interior_medium_id = sh.interior_medium_id;
    vx.exterior_medium_id = sh.exterior_medium_id;
    ShadingInfo si;
    if (prim_is_sphere(pc)) {
        V3 ng_raw = vx.position - xyz(pa);  // sphere.inl:84-85
        vx.st = sphere_st(ng_raw, pa.w);
        vx.geometric_normal = normalize(ng_raw);
        si = shading_info_sphere(sh, vx.st, vx.geometric_normal);
    } else {
        V4 pb = ld4(&sc.prims[hit.prim].b);
        V3 A = xyz(pa), B = mk3(pa.w, pb.x, pb.y), C = mk3(pb.z, pb.w, pc.x);
        vx.geometric_normal = normalize(cross(B - A, C - A));
        vx.st = mk2(hit.u, hit.v);
        si = shading_info_mesh(sc, sh, vx.primitive_id, vx.st, vx.geometric_normal);
    }
    vx.shading_frame = si.frame;
    vx.uv = si.uv;
    vx.mean_curvature = si.mean_curvature;
    vx.ray_radius = rd_radius + rd_spread * distance(org, vx.position);  // ray.h:40-42
    vx.uv_screen_size = vx.ray_radius / si.inv_uv_size;
    if (dot(vx.geometric_normal, vx.shading_frame.n) < 0) vx.geometric_normal = -vx.geometric_normal;
    return vx;
}

// ---- sampling points on shapes -------------------------------------------------------------
// triangle_mesh.inl:24-50
LJ_HD PointAndNormal sample_point_on_mesh(const DevScene &sc, const DevShape &sh, V2 uv, float w) {
    int tri = sample_table_1d(sc.tri_cdf + sh.cdf_offset, sh.num_tris, w);
    const int *idx = sc.indices + 3 * (sh.tri_offset + tri);
    V3 v0 = ld3(sc.positions, idx[0]), v1 = ld3(sc.positions, idx[1]), v2 = ld3(sc.positions, idx[2]);
    V3 e1 = v1 - v0, e2 = v2 - v0;
    float a = sqrtf(clampf(uv.x, 0.f, 1.f));
    float b1 = 1 - a, b2 = a * uv.y;
    V3 ng = normalize(cross(e1, e2));
    if (sh.has_normals) {
        V3 n0 = ld3(sc.normals, idx[0]), n1 = ld3(sc.normals, idx[1]), n2 = ld3(sc.normals, idx[2]);
        V3 sn = normalize((1 - b1 - b2) * n0 + b1 * n1 + b2 * n2);
        if (dot(ng, sn) < 0) ng = -ng;
    }
    PointAndNormal pn;
    pn.position = v0 + e1 * b1 + e2 * b2;
    pn.normal = ng;
    return pn;
}

// sphere.inl:164-213 (pbrt-v3 cone sampling)
LJ_HD PointAndNormal sample_point_on_sphere(const DevShape &sh, V3 ref, V2 uv) {
    V3 center = mk3(sh.cx, sh.cy, sh.cz);
    float r = sh.radius;
    PointAndNormal pn;
    if (distance_squared(ref, center) < r * r) {
        float z = 1 - 2 * uv.x;
        float r_ = sqrtf(fmaxf(0.f, 1 - z * z));
        float phi = 2 * kPi * uv.y;
        V3 off = mk3(r_ * cosf(phi), r_ * sinf(phi), z);
        pn.position = center + r * off;
        pn.normal = off;
        return pn;
    }
    V3 dir_to_center = normalize(center - ref);
    Frame frame = make_frame(dir_to_center);
    float sin_el_max_sq = r * r / distance_squared(ref, center);
    float cos_el_max = sqrtf(fmaxf(0.f, 1 - sin_el_max_sq));
    // cos_el = (1-u) + u cos_el_max, written through 1-cos so small lights keep their precision
    // in fp32 (1 - cos_max = sin^2_max / (1 + cos_max); sin^2 = (1-cos)(1+cos)).
    float omc = uv.x * (sin_el_max_sq / (1 + cos_el_max));
    float cos_el = 1 - omc;
    float sin_el_sq = omc * (2 - omc);
    float azimuth = uv.y * 2 * kPi;
    float dc = distance(ref, center);
    float ds = dc * cos_el - sqrtf(fmaxf(0.f, r * r - dc * dc * sin_el_sq));
    float cos_alpha = (dc * dc + r * r - ds * ds) / (2 * dc * r);
    float sin_alpha = sqrtf(fmaxf(0.f, 1 - cos_alpha * cos_alpha));
    V3 n = -to_world(frame, mk3(sin_alpha * cosf(azimuth), sin_alpha * sinf(azimuth), cos_alpha));
    pn.position = r * n + center;
    pn.normal = n;
    return pn;
}

LJ_HD PointAndNormal sample_point_on_shape(const DevScene &sc, int shape_id, V3 ref, V2 uv, float w) {
    const DevShape &sh = sc.shapes[shape_id];
    if (sh.type == 0) return sample_point_on_sphere(sh, ref, uv);
    return sample_point_on_mesh(sc, sh, uv, w);
}

// triangle_mesh.inl:52-58, sphere.inl:215-238
LJ_HD float pdf_point_on_shape(const DevScene &sc, int shape_id, const PointAndNormal &pn, V3 ref) {
    const DevShape &sh = sc.shapes[shape_id];
    if (sh.type != 0) return 1 / sh.total_area;
    V3 center = mk3(sh.cx, sh.cy, sh.cz);
    float r = sh.radius;
    if (distance_squared(ref, center) < r * r) return 1 / (4 * kPi * r * r);
    float sin_el_max_sq = r * r / distance_squared(ref, center);
    float cos_el_max = sqrtf(fmaxf(0.f, 1 - sin_el_max_sq));
    // 1 - cos of a small cone cancels badly in fp32: use sin^2 / (1 + cos) instead.
    float one_minus_cos = sin_el_max_sq / (1 + cos_el_max);
    float pdf_solid_angle = 1 / (2 * kPi * one_minus_cos);
    V3 dir = normalize(pn.position - ref);
    return pdf_solid_angle * fabsf(dot(pn.normal, dir)) / distance_squared(ref, pn.position);
}

LJ_HD float surface_area(const DevShape &sh) {
    return sh.type == 0 ? 4 * kPi * sh.radius * sh.radius : sh.total_area;
}

}  // namespace lj
