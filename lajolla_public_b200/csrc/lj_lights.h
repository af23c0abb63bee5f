// lj_lights.h -- light selection, sampling, pdf and emission in device memory.
// fp32 restatement of the reference's scene.cpp:61-67, light.cpp, lights/diffuse_area_light.inl,
// lights/envmap.inl:7-72, table_dist.cpp:117-151.
#pragma once
#include "lj_shapes.h"
#include "lj_texture.h"

namespace lj {

LJ_HD int sample_light(const DevScene &sc, float u) { return sample_table_1d(sc.light_cdf, sc.num_lights, u); }
LJ_HD float light_pmf(const DevScene &sc, int light_id) { return sc.light_pmf[light_id]; }
LJ_HD bool light_is_envmap(const DevLight &l) { return l.type == 1; }

// table_dist.cpp:117-140
LJ_HD V2 sample_table_2d(const DevTable2D &t, V2 rnd) {
    int w = t.width, h = t.height;
    int yo = sample_table_1d(t.cdf_marginals, h, rnd.y);
    float dy = rnd.y - t.cdf_marginals[yo];
    float ry = t.cdf_marginals[yo + 1] - t.cdf_marginals[yo];
    if (ry > 0) dy /= ry;
    const float *cdf = t.cdf_rows + (size_t)yo * (w + 1);
    int xo = sample_table_1d(cdf, w, rnd.x);
    float dx = rnd.x - cdf[xo];
    float rx = cdf[xo + 1] - cdf[xo];
    if (rx > 0) dx /= rx;
    return mk2((xo + dx) / w, (yo + dy) / h);
}
// table_dist.cpp:142-151
LJ_HD float pdf_table_2d(const DevTable2D &t, V2 xy) {
    int w = t.width, h = t.height;
    int x = (int)clampf(xy.x * w, 0.f, (float)(w - 1));
    int y = (int)clampf(xy.y * h, 0.f, (float)(h - 1));
    return t.pdf_marginals[y] * t.pdf_rows[(size_t)y * w + x] * w * h;
}

// lights/*.inl sample_point_on_light
LJ_HD PointAndNormal sample_point_on_light(const DevScene &sc, const DevLight &l, V3 ref, V2 uv, float w) {
    if (l.type == 0) return sample_point_on_shape(sc, l.shape_id, ref, uv, w);
    V2 p = sample_table_2d(sc.envmap_dist, uv);
    float azimuth = p.x * (2 * kPi), elevation = p.y * kPi;
    V3 local = mk3(sinf(azimuth) * sinf(elevation), cosf(elevation), -cosf(azimuth) * sinf(elevation));
    V3 world = xform_vector(l.to_world, local);
    PointAndNormal pn;
    pn.position = mk3(0);
    pn.normal = -world;  // envmap.inl:19: the normal slot carries -world_dir
    return pn;
}

LJ_HD V2 envmap_uv(V3 local_dir) {  // envmap.inl:25-31
    V2 uv = mk2(atan2f(local_dir.x, -local_dir.z) * kInvTwoPi, acosf(clampf(local_dir.y, -1.f, 1.f)) * kInvPi);
    if (uv.x < 0) uv.x += 1;
    return uv;
}

LJ_HD float pdf_point_on_light(const DevScene &sc, const DevLight &l, const PointAndNormal &pn, V3 ref) {
    if (l.type == 0) return pdf_point_on_shape(sc, l.shape_id, pn, ref);
    V3 local = xform_vector(l.to_local, -pn.normal);
    V2 uv = envmap_uv(local);
    float cos_el = local.y;
    float sin_el = sqrtf(clampf(1 - cos_el * cos_el, 0.f, 1.f));
    if (sin_el <= 0) return 0;
    return pdf_table_2d(sc.envmap_dist, uv) / (2 * kPi * kPi * sin_el);
}

// emission(light, view_dir, footprint, point) -- diffuse_area_light.inl:15-20, envmap.inl:46-72
LJ_HD V3 light_emission(const DevScene &sc, const DevLight &l, V3 view_dir, float footprint, const PointAndNormal &pn) {
    if (l.type == 0) {
        if (dot(pn.normal, view_dir) <= 0) return mk3(0);
        return mk3(l.intensity[0], l.intensity[1], l.intensity[2]);
    }
    V3 w = xform_vector(l.to_local, -view_dir);
    V2 uv = envmap_uv(w);
    float dudwx = -w.z / (w.x * w.x + w.z * w.z);
    float dudwz = w.x / (w.x * w.x + w.z * w.z);
    float dvdwy = -1 / sqrtf(fmaxf(1 - w.y * w.y, 0.f));
    // TRAP: upstream takes min() with the (negative) dv/dwy and never uses view_footprint; kept.
    float fp = fminf(sqrtf(dudwx * dudwx + dudwz * dudwz), dvdwy);
    (void)footprint;
    return eval_tex3(sc, l.values, uv, fp) * l.scale;
}

// emission(PathVertex, view_dir, scene) -- intersection.cpp:87-98
LJ_HD V3 vertex_emission(const DevScene &sc, const Vertex &vx, V3 view_dir) {
    int light_id = sc.shapes[vx.shape_id].area_light_id;
    PointAndNormal pn;
    pn.position = vx.position;
    pn.normal = vx.geometric_normal;
    return light_emission(sc, sc.lights[light_id], view_dir, vx.uv_screen_size, pn);
}

}  // namespace lj
