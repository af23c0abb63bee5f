// lj_camera.h -- primary-ray generation.  fp32 restatement of the reference's camera.cpp:23-47,
// filters/{box,tent,gaussian}.inl and ray.h:35-66.
#pragma once
#include "lj_scene_dev.h"

namespace lj {

LJ_HD V2 sample_filter(int type, float param, V2 rnd) {
    if (type == 0) {  // box.inl:1-4
        return mk2((2 * rnd.x - 1) * (param / 2), (2 * rnd.y - 1) * (param / 2));
    } else if (type == 1) {  // tent.inl:28-35
        float h = param / 2;
        float x = rnd.x < 0.5f ? h * (sqrtf(2 * rnd.x) - 1) : h * (1 - sqrtf(1 - 2 * (rnd.x - 0.5f)));
        float y = rnd.y < 0.5f ? h * (sqrtf(2 * rnd.y) - 1) : h * (1 - sqrtf(1 - 2 * (rnd.y - 0.5f)));
        return mk2(x, y);
    }
    // gaussian.inl:1-7 (Box-Muller)
    float r = param * sqrtf(-2 * logf(fmaxf(rnd.x, 1e-8f)));
    return mk2(r * cosf(2 * kPi * rnd.y), r * sinf(2 * kPi * rnd.y));
}

// camera.cpp:23-47.  screen_pos in [0,1]^2.
LJ_HD void sample_primary(const DevCamera &cam, V2 screen_pos, V3 &org, V3 &dir) {
    float px = screen_pos.x * cam.width, py = screen_pos.y * cam.height;
    float fx = floorf(px), fy = floorf(py);
    V2 off = sample_filter(cam.filter_type, cam.filter_param, mk2(px - fx, py - fy));
    float rx = (fx + 0.5f + off.x) / cam.width;
    float ry = (fy + 0.5f + off.y) / cam.height;
    V3 pt = xform_point(cam.sample_to_cam, mk3(rx, ry, 0));
    V3 d = normalize(pt);
    org = xform_point(cam.cam_to_world, mk3(0, 0, 0));
    dir = normalize(xform_vector(cam.cam_to_world, d));
}

// The wavefront path generator knows the pixel and the two sub-pixel uniforms separately; this
// avoids the (x + u)/w * w round trip of path_tracing.h:11-12 losing bits in fp32.
LJ_HD void sample_primary_pixel(const DevCamera &cam, int x, int y, V2 u, V3 &org, V3 &dir) {
    V2 off = sample_filter(cam.filter_type, cam.filter_param, u);
    float rx = (x + 0.5f + off.x) / cam.width;
    float ry = (y + 0.5f + off.y) / cam.height;
    V3 pt = xform_point(cam.sample_to_cam, mk3(rx, ry, 0));
    V3 d = normalize(pt);
    org = xform_point(cam.cam_to_world, mk3(0, 0, 0));
    dir = normalize(xform_vector(cam.cam_to_world, d));
}

// ray.h:35-66
LJ_HD float init_ray_spread(int w, int h) { return 0.25f / (float)(w > h ? w : h); }
LJ_HD float spread_reflect(float radius, float spread, float mean_curvature, float roughness) {
    float spec = spread + 2 * mean_curvature * radius;
    return fmaxf(spec * (1 - roughness) + 0.2f * roughness, 0.f);
}
LJ_HD float spread_refract(float radius, float spread, float mean_curvature, float eta, float roughness) {
    float spec = (spread + 2 * mean_curvature * radius) / eta;
    return fmaxf(spec * (1 - roughness) + 0.2f * roughness, 0.f);
}

}  // namespace lj
