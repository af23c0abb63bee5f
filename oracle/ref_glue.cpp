// oracle/ref_glue.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points over the UNMODIFIED reference objects (compiled from /root/reference/src by
// oracle/Makefile, linked with oracle/embree_shim.cpp): known-answer batch forms of the reference's
// own per-ray functions, and a dumper that flattens a parsed reference Scene into the .ljs container
// the parity tests feed to the CUDA library.  Only tests/, __graft_entry__.smoke() and bench.py's
// CPU-baseline leg load the resulting oracle/_ref/libljoracle.so.  Query/result layouts are the POD
// structs of include/lajolla_b200.h (fp32 in/out; all arithmetic in between is the reference's
// double code).  Functions called, by reference file:
//   parse_scene                      parsers/parse_scene.cpp:1602
//   intersect / occluded / emission  intersection.cpp:7-98
//   eval / pdf_sample_bsdf / sample_bsdf   material.cpp:90-123
//   sample_light / light_pmf         scene.cpp:61-67
//   sample_point_on_light / pdf_point_on_light / emission   light.cpp
//   sample_primary                   camera.cpp:23-47
//   eval(Texture)                    texture.h:161-163
//   init_pcg32 / next_pcg32          pcg.h:22-46
//   render                           render.cpp:155
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../include/lajolla_b200.h"
#include "camera.h"
#include "intersection.h"
#include "light.h"
#include "material.h"
#include "medium.h"
#include "parallel.h"
#include "parsers/parse_scene.h"
// pcg.h defines its two explicit specialisations non-inline; render.cpp.o already carries them.
#define next_pcg32_real ljo_next_pcg32_real
#include "pcg.h"
#undef next_pcg32_real
#include "render.h"
#include "scene.h"
#include "texture.h"
#include "image.h"

namespace {

RTCDevice g_device = nullptr;
bool g_parallel = false;

struct Writer {
    FILE *f;
    void i32(int32_t v) { fwrite(&v, 4, 1, f); }
    void f32(double v) { float x = (float)v; fwrite(&x, 4, 1, f); }
    void v3(const Vector3 &v) { f32(v.x); f32(v.y); f32(v.z); }
    void m44(const Matrix4x4 &m) { for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) f32(m(r, c)); }
};

template <typename T> struct TexVal;
template <> struct TexVal<Spectrum> { static void put(Writer &w, const Spectrum &s) { w.v3(s); } };
template <> struct TexVal<Real> { static void put(Writer &w, const Real &s) { w.f32(s); w.f32(s); w.f32(s); } };

// n3 = number of 3-channel images: 1-channel image ids are offset by it in the flat numbering.
template <typename T>
void write_texture(Writer &w, const Texture<T> &t, int n3) {
    constexpr bool is1 = std::is_same<T, Real>::value;
    if (auto *c = std::get_if<ConstantTexture<T>>(&t)) {
        w.i32(LJ_TEX_CONSTANT); w.i32(-1);
        TexVal<T>::put(w, c->value); TexVal<T>::put(w, c->value);
        w.f32(1); w.f32(1); w.f32(0); w.f32(0);
    } else if (auto *im = std::get_if<ImageTexture<T>>(&t)) {
        w.i32(LJ_TEX_IMAGE); w.i32(im->texture_id + (is1 ? n3 : 0));
        w.f32(0); w.f32(0); w.f32(0); w.f32(0); w.f32(0); w.f32(0);
        w.f32(im->uscale); w.f32(im->vscale); w.f32(im->uoffset); w.f32(im->voffset);
    } else if (auto *cb = std::get_if<CheckerboardTexture<T>>(&t)) {
        w.i32(LJ_TEX_CHECKERBOARD); w.i32(-1);
        TexVal<T>::put(w, cb->color0); TexVal<T>::put(w, cb->color1);
        w.f32(cb->uscale); w.f32(cb->vscale); w.f32(cb->uoffset); w.f32(cb->voffset);
    }
}
void write_empty_texture(Writer &w) {
    w.i32(LJ_TEX_CONSTANT); w.i32(-1);
    for (int i = 0; i < 10; i++) w.f32(i < 6 ? 0 : (i < 8 ? 1 : 0));
}

struct MatWriter {
    Writer &w;
    int n3;
    // slot -> writer; unfilled slots are written as constant 0
    void operator()(const Lambertian &m) const { head(LJ_MAT_LAMBERTIAN, 1); slots({[&] { write_texture(w, m.reflectance, n3); }}); }
    void operator()(const RoughPlastic &m) const {
        head(LJ_MAT_ROUGHPLASTIC, m.eta);
        slots({[&] { write_texture(w, m.diffuse_reflectance, n3); }, [&] { write_texture(w, m.specular_reflectance, n3); }, [&] { write_texture(w, m.roughness, n3); }});
    }
    void operator()(const RoughDielectric &m) const {
        head(LJ_MAT_ROUGHDIELECTRIC, m.eta);
        slots({[&] { write_texture(w, m.specular_transmittance, n3); }, [&] { write_texture(w, m.specular_reflectance, n3); }, [&] { write_texture(w, m.roughness, n3); }});
    }
    void operator()(const DisneyDiffuse &m) const {
        head(LJ_MAT_DISNEY_DIFFUSE, 1);
        slots({[&] { write_texture(w, m.base_color, n3); }, [&] { write_texture(w, m.subsurface, n3); }, [&] { write_texture(w, m.roughness, n3); }});
    }
    void operator()(const DisneyMetal &m) const {
        head(LJ_MAT_DISNEY_METAL, 1);
        slots({[&] { write_texture(w, m.base_color, n3); }, nullptr, [&] { write_texture(w, m.roughness, n3); }, [&] { write_texture(w, m.anisotropic, n3); }});
    }
    void operator()(const DisneyGlass &m) const {
        head(LJ_MAT_DISNEY_GLASS, m.eta);
        slots({[&] { write_texture(w, m.base_color, n3); }, nullptr, [&] { write_texture(w, m.roughness, n3); }, [&] { write_texture(w, m.anisotropic, n3); }});
    }
    void operator()(const DisneyClearcoat &m) const {
        head(LJ_MAT_DISNEY_CLEARCOAT, 1);
        slots({nullptr, nullptr, nullptr, nullptr, [&] { write_texture(w, m.clearcoat_gloss, n3); }});
    }
    void operator()(const DisneySheen &m) const {
        head(LJ_MAT_DISNEY_SHEEN, 1);
        slots({[&] { write_texture(w, m.base_color, n3); }, nullptr, nullptr, nullptr, nullptr, [&] { write_texture(w, m.sheen_tint, n3); }});
    }
    void operator()(const DisneyBSDF &m) const {
        head(LJ_MAT_DISNEY_BSDF, m.eta);
        slots({[&] { write_texture(w, m.base_color, n3); }, [&] { write_texture(w, m.subsurface, n3); }, [&] { write_texture(w, m.roughness, n3); },
               [&] { write_texture(w, m.anisotropic, n3); }, [&] { write_texture(w, m.clearcoat_gloss, n3); }, [&] { write_texture(w, m.sheen_tint, n3); },
               [&] { write_texture(w, m.specular_transmission, n3); }, [&] { write_texture(w, m.metallic, n3); }, [&] { write_texture(w, m.specular, n3); },
               [&] { write_texture(w, m.specular_tint, n3); }, [&] { write_texture(w, m.sheen, n3); }, [&] { write_texture(w, m.clearcoat, n3); }});
    }
    void head(int type, double eta) const { w.i32(type); w.f32(eta); }
    void slots(std::initializer_list<std::function<void()>> fs) const {
        int k = 0;
        for (auto &f : fs) { if (f) f(); else write_empty_texture(w); k++; }
        for (; k < LJ_NUM_TEX_SLOTS; k++) write_empty_texture(w);
    }
};

void write_volume(Writer &w, const VolumeSpectrum &v) {
    if (auto *c = std::get_if<ConstantVolume<Spectrum>>(&v)) {
        w.i32(0); w.i32(0); w.i32(0); w.i32(0);
        w.v3(c->value); w.v3(Vector3{0, 0, 0}); w.v3(Vector3{0, 0, 0}); w.f32(1);
    } else if (auto *g = std::get_if<GridVolume<Spectrum>>(&v)) {
        w.i32(1); w.i32(g->resolution.x); w.i32(g->resolution.y); w.i32(g->resolution.z);
        w.v3(g->max_data); w.v3(g->p_min); w.v3(g->p_max); w.f32(g->scale);
        for (const Spectrum &s : g->data) w.v3(s);
    }
}

lj_vertex to_abi(const PathVertex &v) {
    lj_vertex o;
    memset(&o, 0, sizeof(o));
    for (int i = 0; i < 3; i++) {
        o.position[i] = (float)v.position[i]; o.geometric_normal[i] = (float)v.geometric_normal[i];
        o.frame_x[i] = (float)v.shading_frame.x[i]; o.frame_y[i] = (float)v.shading_frame.y[i]; o.frame_n[i] = (float)v.shading_frame.n[i];
    }
    o.st[0] = (float)v.st.x; o.st[1] = (float)v.st.y; o.uv[0] = (float)v.uv.x; o.uv[1] = (float)v.uv.y;
    o.uv_screen_size = (float)v.uv_screen_size; o.mean_curvature = (float)v.mean_curvature; o.ray_radius = (float)v.ray_radius;
    o.shape_id = v.shape_id; o.primitive_id = v.primitive_id; o.material_id = v.material_id;
    o.interior_medium_id = v.interior_medium_id; o.exterior_medium_id = v.exterior_medium_id;
    return o;
}
PathVertex from_abi(const lj_vertex &o) {
    PathVertex v;
    v.position = Vector3{o.position[0], o.position[1], o.position[2]};
    v.geometric_normal = Vector3{o.geometric_normal[0], o.geometric_normal[1], o.geometric_normal[2]};
    v.shading_frame = Frame(Vector3{o.frame_x[0], o.frame_x[1], o.frame_x[2]}, Vector3{o.frame_y[0], o.frame_y[1], o.frame_y[2]},
                            Vector3{o.frame_n[0], o.frame_n[1], o.frame_n[2]});
    v.st = Vector2{o.st[0], o.st[1]}; v.uv = Vector2{o.uv[0], o.uv[1]};
    v.uv_screen_size = o.uv_screen_size; v.mean_curvature = o.mean_curvature; v.ray_radius = o.ray_radius;
    v.shape_id = o.shape_id; v.primitive_id = o.primitive_id; v.material_id = o.material_id;
    v.interior_medium_id = o.interior_medium_id; v.exterior_medium_id = o.exterior_medium_id;
    return v;
}
Ray to_ray(const lj_ray &r) {
    return Ray{Vector3{r.org[0], r.org[1], r.org[2]}, Vector3{r.dir[0], r.dir[1], r.dir[2]}, r.tnear, r.tfar};
}

}  // namespace

extern "C" {

// Counters kept by the shim (rtcIntersect1 / rtcOccluded1 calls).
void ljshim_get_counters(unsigned long long *closest, unsigned long long *any);
void ljshim_reset_counters(void);

void *ljo_scene_load(const char *xml_path, int num_threads) {
    try {
        if (!g_device) g_device = rtcNewDevice(nullptr);
        if (!g_parallel) { parallel_init(num_threads > 0 ? num_threads : 1); g_parallel = true; }
        std::unique_ptr<Scene> s = parse_scene(xml_path, g_device);
        return s.release();
    } catch (std::exception &e) {
        fprintf(stderr, "ljo_scene_load: %s\n", e.what());
        return nullptr;
    }
}
void ljo_scene_free(void *h) { delete (Scene *)h; }
// Joins the reference's worker threads (parallel.cpp:258-274); without it the process hangs at exit.
void ljo_shutdown(void) { if (g_parallel) { parallel_cleanup(); g_parallel = false; } }

void ljo_set_integrator(void *h, int integrator) { ((Scene *)h)->options.integrator = (Integrator)integrator; }  // scene.h:14-22
void ljo_set_spp(void *h, int spp) { ((Scene *)h)->options.samples_per_pixel = spp; }

// Flatten to the .ljs container (layout documented in lajolla_public_b200/ljs.py).
int ljo_scene_dump(void *h, const char *out_path) {
    const Scene &s = *(const Scene *)h;
    FILE *f = fopen(out_path, "wb");
    if (!f) return 1;
    Writer w{f};
    fwrite("LJS1", 4, 1, f);
    w.i32(1);
    const Camera &c = s.camera;
    w.m44(c.cam_to_world); w.m44(c.world_to_cam); w.m44(c.sample_to_cam); w.m44(c.cam_to_sample);
    w.i32(c.width); w.i32(c.height);
    if (auto *b = std::get_if<Box>(&c.filter)) { w.i32(LJ_FILTER_BOX); w.f32(b->width); }
    else if (auto *t = std::get_if<Tent>(&c.filter)) { w.i32(LJ_FILTER_TENT); w.f32(t->width); }
    else { w.i32(LJ_FILTER_GAUSSIAN); w.f32(std::get<Gaussian>(c.filter).stddev); }
    w.i32(c.medium_id);
    w.i32((int)s.options.integrator); w.i32(s.options.samples_per_pixel); w.i32(s.options.max_depth);
    w.i32(s.options.rr_depth); w.i32(s.options.vol_path_version); w.i32(s.options.max_null_collisions);
    int n3 = (int)s.texture_pool.image3s.size(), n1 = (int)s.texture_pool.image1s.size();
    w.i32(n3 + n1); w.i32((int)s.materials.size()); w.i32((int)s.shapes.size()); w.i32((int)s.lights.size());
    w.i32((int)s.media.size()); w.i32(s.envmap_light_id);
    for (const Mipmap3 &m : s.texture_pool.image3s) {
        const Image3 &im = m.images[0];
        w.i32(im.width); w.i32(im.height); w.i32(3);
        for (const Vector3 &p : im.data) w.v3(p);
    }
    for (const Mipmap1 &m : s.texture_pool.image1s) {
        const Image1 &im = m.images[0];
        w.i32(im.width); w.i32(im.height); w.i32(1);
        for (Real p : im.data) w.f32(p);
    }
    for (const Material &m : s.materials) std::visit(MatWriter{w, n3}, m);
    for (const Shape &sh : s.shapes) {
        if (auto *sp = std::get_if<Sphere>(&sh)) {
            w.i32(LJ_SHAPE_SPHERE); w.i32(sp->material_id); w.i32(sp->area_light_id); w.i32(sp->interior_medium_id); w.i32(sp->exterior_medium_id);
            w.v3(sp->position); w.f32(sp->radius);
            w.i32(0); w.i32(0); w.i32(0); w.i32(0);
        } else {
            const TriangleMesh &m = std::get<TriangleMesh>(sh);
            w.i32(LJ_SHAPE_MESH); w.i32(m.material_id); w.i32(m.area_light_id); w.i32(m.interior_medium_id); w.i32(m.exterior_medium_id);
            w.v3(Vector3{0, 0, 0}); w.f32(0);
            w.i32((int)m.positions.size()); w.i32((int)m.indices.size()); w.i32(m.normals.size() > 0); w.i32(m.uvs.size() > 0);
            for (auto &p : m.positions) w.v3(p);
            for (auto &t : m.indices) { w.i32(t[0]); w.i32(t[1]); w.i32(t[2]); }
            for (auto &n : m.normals) w.v3(n);
            for (auto &uv : m.uvs) { w.f32(uv.x); w.f32(uv.y); }
        }
    }
    for (const Light &l : s.lights) {
        if (auto *a = std::get_if<DiffuseAreaLight>(&l)) {
            w.i32(LJ_LIGHT_AREA); w.i32(a->shape_id); w.v3(a->intensity);
            write_empty_texture(w);
            Matrix4x4 id = Matrix4x4::identity();
            w.m44(id); w.m44(id); w.f32(1);
        } else {
            const Envmap &e = std::get<Envmap>(l);
            w.i32(LJ_LIGHT_ENVMAP); w.i32(-1); w.v3(Vector3{0, 0, 0});
            write_texture(w, e.values, n3);
            w.m44(e.to_world); w.m44(e.to_local); w.f32(e.scale);
        }
    }
    for (const Medium &m : s.media) {
        PhaseFunction pf = get_phase_function(m);
        int ptype = std::get_if<HenyeyGreenstein>(&pf) ? LJ_PHASE_HG : LJ_PHASE_ISOTROPIC;
        double g = ptype == LJ_PHASE_HG ? std::get<HenyeyGreenstein>(pf).g : 0;
        if (auto *hm = std::get_if<HomogeneousMedium>(&m)) {
            w.i32(LJ_MEDIUM_HOMOGENEOUS); w.i32(ptype); w.f32(g); w.v3(hm->sigma_a); w.v3(hm->sigma_s);
        } else {
            const HeterogeneousMedium &het = std::get<HeterogeneousMedium>(m);
            w.i32(LJ_MEDIUM_HETEROGENEOUS); w.i32(ptype); w.f32(g); w.v3(Vector3{0, 0, 0}); w.v3(Vector3{0, 0, 0});
            write_volume(w, het.albedo);
            write_volume(w, het.density);
        }
    }
    fclose(f);
    return 0;
}

void ljo_scene_info(void *h, double *bsphere /*radius, cx, cy, cz*/, double *eps, int *counts /*shapes, lights, materials*/) {
    const Scene &s = *(const Scene *)h;
    bsphere[0] = s.bounds.radius; bsphere[1] = s.bounds.center.x; bsphere[2] = s.bounds.center.y; bsphere[3] = s.bounds.center.z;
    *eps = get_shadow_epsilon(s);
    counts[0] = (int)s.shapes.size(); counts[1] = (int)s.lights.size(); counts[2] = (int)s.materials.size();
}
void ljo_light_table(void *h, double *pmf, double *cdf) {
    const Scene &s = *(const Scene *)h;
    for (size_t i = 0; i < s.light_dist.pmf.size(); i++) pmf[i] = s.light_dist.pmf[i];
    for (size_t i = 0; i < s.light_dist.cdf.size(); i++) cdf[i] = s.light_dist.cdf[i];
}

void ljo_trace_closest(void *h, const lj_ray *rays, int64_t n, lj_hit *hits) {
    const Scene &s = *(const Scene *)h;
    for (int64_t i = 0; i < n; i++) {
        Ray r = to_ray(rays[i]);
        std::optional<PathVertex> v = intersect(s, r);
        lj_hit o{(float)r.tfar, 0, 0, -1, -1};
        if (v) {
            // t as the reference sees it: float tfar written back by the ray cast
            Vector3 d = v->position - r.org;
            o.t = (float)(dot(d, r.dir) / dot(r.dir, r.dir));
            o.u = (float)v->st.x; o.v = (float)v->st.y; o.shape_id = v->shape_id; o.primitive_id = v->primitive_id;
        }
        hits[i] = o;
    }
}
void ljo_trace_any(void *h, const lj_ray *rays, int64_t n, uint8_t *occ) {
    const Scene &s = *(const Scene *)h;
    for (int64_t i = 0; i < n; i++) occ[i] = occluded(s, to_ray(rays[i])) ? 1 : 0;
}
void ljo_intersect(void *h, const lj_ray *rays, const float *rd, int64_t n, lj_vertex *out) {
    const Scene &s = *(const Scene *)h;
    for (int64_t i = 0; i < n; i++) {
        RayDifferential d{rd ? rd[2 * i] : 0.0, rd ? rd[2 * i + 1] : 0.0};
        std::optional<PathVertex> v = intersect(s, to_ray(rays[i]), d);
        if (v) out[i] = to_abi(*v);
        else { memset(&out[i], 0, sizeof(lj_vertex)); out[i].shape_id = out[i].primitive_id = out[i].material_id = out[i].interior_medium_id = out[i].exterior_medium_id = -1; }
    }
}
void ljo_bsdf(void *h, const lj_bsdf_query *q, int64_t n, lj_bsdf_result *out) {
    const Scene &s = *(const Scene *)h;
    for (int64_t i = 0; i < n; i++) {
        PathVertex v = from_abi(q[i].vertex);
        const Material &m = s.materials[v.material_id];
        Vector3 wi{q[i].dir_in[0], q[i].dir_in[1], q[i].dir_in[2]}, wo{q[i].dir_out[0], q[i].dir_out[1], q[i].dir_out[2]};
        TransportDirection td = q[i].transport == 0 ? TransportDirection::TO_LIGHT : TransportDirection::TO_VIEW;
        lj_bsdf_result r;
        memset(&r, 0, sizeof(r));
        Spectrum f = eval(m, wi, wo, v, s.texture_pool, td);
        r.f[0] = (float)f.x; r.f[1] = (float)f.y; r.f[2] = (float)f.z;
        r.pdf = (float)pdf_sample_bsdf(m, wi, wo, v, s.texture_pool, td);
        auto bs = sample_bsdf(m, wi, v, s.texture_pool, Vector2{q[i].rnd_uv[0], q[i].rnd_uv[1]}, (Real)q[i].rnd_w, td);
        if (bs) {
            r.sampled = 1;
            r.s_dir_out[0] = (float)bs->dir_out.x; r.s_dir_out[1] = (float)bs->dir_out.y; r.s_dir_out[2] = (float)bs->dir_out.z;
            r.s_eta = (float)bs->eta; r.s_roughness = (float)bs->roughness;
        }
        out[i] = r;
    }
}
void ljo_light(void *h, const lj_light_query *q, int64_t n, lj_light_result *out) {
    const Scene &s = *(const Scene *)h;
    for (int64_t i = 0; i < n; i++) {
        Vector3 ref{q[i].ref_point[0], q[i].ref_point[1], q[i].ref_point[2]};
        int id = sample_light(s, q[i].light_w);
        const Light &l = s.lights[id];
        PointAndNormal pn = sample_point_on_light(l, ref, Vector2{q[i].rnd_uv[0], q[i].rnd_uv[1]}, q[i].rnd_w, s);
        lj_light_result r;
        memset(&r, 0, sizeof(r));
        r.light_id = id;
        for (int c = 0; c < 3; c++) { r.position[c] = (float)pn.position[c]; r.normal[c] = (float)pn.normal[c]; }
        r.pmf = (float)light_pmf(s, id);
        r.pdf = (float)pdf_point_on_light(l, pn, ref, s);
        Vector3 dir_light = is_envmap(l) ? -pn.normal : normalize(pn.position - ref);
        Spectrum L = emission(l, -dir_light, Real(0), pn, s);
        r.emission[0] = (float)L.x; r.emission[1] = (float)L.y; r.emission[2] = (float)L.z;
        out[i] = r;
    }
}
// medium.h:25-31 / phase_function.h:18-29 on the reference's own objects (see lj_medium_batch)
void ljo_medium(void *h, const lj_medium_query *q, int64_t n, lj_medium_result *out) {
    const Scene &s = *(const Scene *)h;
    for (int64_t i = 0; i < n; i++) {
        const Medium &m = s.media[q[i].medium_id];
        Vector3 o{q[i].org[0], q[i].org[1], q[i].org[2]}, d{q[i].dir[0], q[i].dir[1], q[i].dir[2]};
        Ray ray{o, d, Real(0), (Real)q[i].tfar};
        Spectrum maj = get_majorant(m, ray);
        Vector3 p = o + Real(q[i].t) * d;
        Spectrum sa = get_sigma_a(m, p), ss = get_sigma_s(m, p);
        PhaseFunction ph = get_phase_function(m);
        lj_medium_result r;
        memset(&r, 0, sizeof(r));
        std::optional<Vector3> pd = sample_phase_function(ph, -d, Vector2{q[i].rnd[0], q[i].rnd[1]});
        for (int c = 0; c < 3; c++) {
            r.majorant[c] = (float)maj[c]; r.sigma_a[c] = (float)sa[c]; r.sigma_s[c] = (float)ss[c];
            if (pd) r.phase_dir[c] = (float)(*pd)[c];
        }
        if (pd) {
            r.phase_eval = (float)eval(ph, -d, *pd).x;
            r.phase_pdf = (float)pdf_sample_phase(ph, -d, *pd);
        }
        out[i] = r;
    }
}
int ljo_num_media(void *h) { return (int)((const Scene *)h)->media.size(); }
void ljo_camera_rays(void *h, const float *xy, int64_t n, lj_ray *rays) {
    const Scene &s = *(const Scene *)h;
    for (int64_t i = 0; i < n; i++) {
        Ray r = sample_primary(s.camera, Vector2{xy[2 * i], xy[2 * i + 1]});
        for (int c = 0; c < 3; c++) { rays[i].org[c] = (float)r.org[c]; rays[i].dir[c] = (float)r.dir[c]; }
        rays[i].tnear = (float)r.tnear; rays[i].tfar = (float)r.tfar;
    }
}
void ljo_texture(void *h, int material_id, const float *q, int64_t n, float *out) {
    const Scene &s = *(const Scene *)h;
    TextureSpectrum t = get_texture(s.materials[material_id]);
    for (int64_t i = 0; i < n; i++) {
        Spectrum v = eval(t, Vector2{q[3 * i], q[3 * i + 1]}, (Real)q[3 * i + 2], s.texture_pool);
        out[3 * i] = (float)v.x; out[3 * i + 1] = (float)v.y; out[3 * i + 2] = (float)v.z;
    }
}
// one mip level of the reference's own pyramid (mipmap.h make_mipmap), 3-channel images
int ljo_mip_level(void *h, int image3_id, int level, int *w, int *hh, float *data) {
    const Scene &s = *(const Scene *)h;
    const Mipmap3 &m = s.texture_pool.image3s[image3_id];
    if (level >= (int)m.images.size()) return (int)m.images.size();
    const Image3 &im = m.images[level];
    *w = im.width; *hh = im.height;
    if (data) for (size_t i = 0; i < im.data.size(); i++) { data[3 * i] = (float)im.data[i].x; data[3 * i + 1] = (float)im.data[i].y; data[3 * i + 2] = (float)im.data[i].z; }
    return (int)m.images.size();
}
void ljo_pcg32(uint64_t first_stream, uint64_t seed, int n_streams, int n_draws, uint32_t *out_u32, double *out_f64) {
    for (int i = 0; i < n_streams; i++) {
        pcg32_state a = init_pcg32(first_stream + i, seed), b = a;
        for (int k = 0; k < n_draws; k++) {
            if (out_u32) out_u32[(size_t)i * n_draws + k] = next_pcg32(a);
            if (out_f64) out_f64[(size_t)i * n_draws + k] = ljo_next_pcg32_real<double>(b);
        }
    }
}
// render() of the reference itself; out = w*h*3 fp32.  Returns seconds spent inside render().
double ljo_render(void *h, float *out) {
    const Scene &s = *(const Scene *)h;
    auto t0 = std::chrono::steady_clock::now();
    Image3 img = render(s);
    auto t1 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < img.data.size(); i++) { out[3 * i] = (float)img.data[i].x; out[3 * i + 1] = (float)img.data[i].y; out[3 * i + 2] = (float)img.data[i].z; }
    return std::chrono::duration<double>(t1 - t0).count();
}
void ljo_film_size(void *h, int *w, int *hh, int *spp) {
    const Scene &s = *(const Scene *)h;
    *w = s.camera.width; *hh = s.camera.height; *spp = s.options.samples_per_pixel;
}

}  // extern "C"
