// lj_path.h -- the per-sample radiance estimator of the reference's path_tracing.h:7-325, cut into
// wavefront stages.  One call of path_tracing() becomes: generate (path_tracing.h:10-14) ->
// [extend -> shade -> shadow]* where each shade() is "the second half of loop iteration k"
// (MIS-weighted emission at the vertex the extension ray found, :242-307, Russian roulette,
// :311-318) followed by "the first half of iteration k+1" (NEE sample, :94-207, BSDF sample,
// :209-237).  Per-path random numbers are drawn in exactly the order of the reference loop, from
// a private PCG32 stream.  Path records live in HBM as a structure of 16-byte arrays.
#pragma once
#include "lj_camera.h"
#include "lj_lights.h"
#include "lj_materials.h"
#include "lj_pcg.h"

namespace lj {

struct U4 { uint32_t x, y, z, w; };

// Structure-of-arrays path pool: every field is one 16-byte record per path slot.
struct PathPool {
    V4 *ray_o;  // org.xyz, tnear
    V4 *ray_d;  // dir.xyz, tfar
    V4 *hit;    // t, u, v, bits(prim)           (written by extend)
    V4 *thr;    // throughput.xyz (f/pdf of the last bounce folded in), last bsdf pdf (solid angle; <0 = camera ray)
    V4 *rad;    // radiance.xyz accumulated by this path, rr_prob for the pending vertex
    V4 *sh_d;   // shadow dir.xyz, tfar (<0: no shadow ray)
    V4 *sh_c;   // NEE contribution if unoccluded .xyz, -
    V4 *meta;   // bits(pixel), bits(nv | flags<<16), bits(rng lo), bits(rng hi)
    V4 *aux;    // eta_scale, ray spread, bits(sample index), bits(medium id)
    // volpath only (null for the path integrator): MIS caches of homework2.tex:521-558 and the NEE walk record
    V4 *vol0;   // multi_trans_dir_pdf.xyz, -
    V4 *vol1;   // multi_trans_nee_pdf.xyz, -
    V4 *vol2;   // nee_p_cache.xyz, -
    V4 *sh_o;   // walk origin.xyz, bits((medium + 1) & 0xffff | budget << 16)
    V4 *sh_pl;  // light point.xyz, bits(walk rng seed)
                // (volpath reuses sh_d.w as pdf_dir = pdf_scatter * G, < 0: no walk; sh_c.w as pdf_nee)
    // volpath, staged NEE walk (wavefront.cu k_walk_begin / k_walk_track): the walk's current segment in the layout the
    // closest-hit kernel reads (ray_o / ray_d / hit / meta of a "walk view" of the pool) and its running products
    V4 *w_o;    // segment origin.xyz, tnear
    V4 *w_d;    // direction.xyz, tfar
    V4 *w_hit;  // t, u, v, bits(prim) of the segment
    V4 *w_meta; // -, bits(kAlive while the walk is in flight), bits(rng lo), bits(rng hi)
    V4 *w_T;    // T_light.xyz, bits(index-matched surfaces crossed)
    V4 *w_pn;   // p_trans_nee.xyz, bits(current medium)
    V4 *w_pd;   // p_trans_dir.xyz, bits(budget)
    // one bit per slot, one word per warp of the shade kernel: the slot carries an NEE shadow ray / walk this wave
    uint32_t *sh_mask;
    // path integrator, scenes with Disney materials: the slots the first shade pass left for the second one (wavefront.cu)
    uint32_t *class_queue;
    int capacity;
};

constexpr uint32_t kAlive = 1u << 16;     // slot holds a path whose extension ray is pending
constexpr uint32_t kOccupied = 1u << 17;  // slot holds radiance not yet flushed to the film

struct PathState {
    V3 o; float tnear;
    V3 d; float tfar;
    Hit hit;
    V3 T; float pdf_sa;
    V3 L; float rr_prob;
    V3 sh_d; float sh_tfar;
    V3 sh_c;
    uint32_t pixel, nv, flags, sample;
    uint64_t rng_state;
    float eta_scale, spread;
    int medium;
    // volpath (lj_volpath.h); nv counts `bounces` there
    V3 mt_dir, mt_nee, nee_p;
    V3 sh_o, sh_pl;
    float sh_pdf_nee, sh_pdf_dir;
    int sh_medium;
    uint32_t sh_budget, sh_seed;
};

struct RenderParams {
    uint32_t spp_total;     // stream = path_stream(pixel * spp_total + sample)
    uint32_t sample_begin, sample_end;
    uint64_t seed;
    int width, height;
};

struct ShadeCounters { uint32_t bounces, shadow_rays, extend_rays, finished; };

LJ_HD Pcg path_rng(const PathState &s, const RenderParams &rp) {
    Pcg r;
    r.state = s.rng_state;
    r.inc = pcg_inc(path_stream((uint64_t)s.pixel * rp.spp_total + s.sample));
    return r;
}

// path_tracing.h:10-14 + state init (:44-53)
LJ_HD void generate_path(const DevScene &sc, const RenderParams &rp, uint32_t pixel, uint32_t sample, PathState &s) {
    Pcg rng = pcg_init(path_stream((uint64_t)pixel * rp.spp_total + sample), rp.seed);
    int x = pixel % rp.width, y = pixel / rp.width;
    float u0 = pcg_uniform(rng), u1 = pcg_uniform(rng);
    sample_primary_pixel(sc.camera, x, y, mk2(u0, u1), s.o, s.d);
    s.tnear = 0;
    s.tfar = LJ_INF;
    s.T = mk3(1);
    s.pdf_sa = -1;
    s.L = mk3(0);
    s.rr_prob = 1;
    s.sh_tfar = -1;
    s.sh_d = mk3(0);
    s.sh_c = mk3(0);
    s.pixel = pixel;
    s.sample = sample;
    s.nv = 1;
    s.flags = kAlive | kOccupied;
    s.rng_state = rng.state;
    s.eta_scale = 1;
    s.spread = init_ray_spread(rp.width, rp.height);
    s.medium = sc.camera.medium_id;
    s.hit.prim = kNoHit;
    s.hit.t = 0; s.hit.u = 0; s.hit.v = 0;
    s.mt_dir = mk3(1); s.mt_nee = mk3(1); s.nee_p = mk3(0);
    s.sh_pdf_dir = -1;
}

LJ_HD float mis_power(float pa, float pb) { return (pa * pa) / (pa * pa + pb * pb); }

// One shade step of the surface path tracer.  On return: s.flags has kAlive iff an extension ray
// was written; s.sh_tfar >= 0 iff a shadow ray was written.
// Material class of the vertex a path's extension ray found (kMatSmall for a miss: the environment is shaded there).
LJ_HD int path_material_class(const DevScene &sc, int prim) {
    if (prim == kNoHit) return kMatSmall;
    const int mid = sc.shapes[prim_shape_id(ld4(&sc.prims[prim].c))].material_id;
    return mid >= 0 && material_is_disney(sc.materials[mid].type) ? kMatDisney : kMatSmall;
}

// CLASS: kMatSmall (inlined dispatch of the three small materials) or kMatDisney (out-of-line Disney dispatchers).
template <int CLASS>
LJ_HD void shade_path(const DevScene &sc, const RenderParams &rp, PathState &s, ShadeCounters &cnt) {
    constexpr bool CALLS = CLASS == kMatDisney;
    constexpr int DISPATCH = CLASS == kMatLambert ? kMatSmall : CLASS;  // (dead code in the Lambertian-only kernel)
    Pcg rng = path_rng(s, rp);
    const bool primary = s.pdf_sa < 0;
    s.sh_tfar = -1;
    s.flags &= ~kAlive;

    if (s.hit.prim == kNoHit) {
        // path_tracing.h:17-28 (camera ray) / :284-307 (bsdf ray): environment map or nothing.
        if (sc.envmap_light_id >= 0) {
            const DevLight &env = sc.lights[sc.envmap_light_id];
            PointAndNormal pn;
            pn.position = mk3(0);
            pn.normal = -s.d;
            V3 Le = light_emission(sc, env, -s.d, s.spread, pn);
            if (primary) {
                s.L += Le;
            } else {
                float p2 = s.pdf_sa;  // G = 1
                float p1 = light_pmf(sc, sc.envmap_light_id) * pdf_point_on_light(sc, env, pn, s.o);
                s.L += s.T * Le * mis_power(p2, p1);
            }
        }
        cnt.finished++;
        return;
    }

    // intersect(): the camera ray carries the pixel's ray differential (path_tracing.h:14-16);
    // bsdf rays are intersected with the default RayDifferential{} (:237), i.e. footprint 0.
    Vertex vx = make_vertex(sc, s.o, s.d, s.hit, 0.f, primary ? s.spread : 0.f);
    const DevShape &shape = sc.shapes[vx.shape_id];
    V3 dir_view = -s.d;

    if (shape.area_light_id >= 0) {
        V3 Le = vertex_emission(sc, vx, dir_view);
        if (primary) {
            s.L += s.T * Le;  // :58-61
        } else {
            // :242-283 -- T already holds throughput * f / pdf_bsdf, so C2/p2 = T * Le
            float G = fabsf(dot(s.d, vx.geometric_normal)) / distance_squared(vx.position, s.o);
            float p2 = s.pdf_sa * G;
            PointAndNormal lp;
            lp.position = vx.position;
            lp.normal = vx.geometric_normal;
            float p1 = light_pmf(sc, shape.area_light_id) * pdf_point_on_light(sc, sc.lights[shape.area_light_id], lp, s.o);
            s.L += s.T * Le * mis_power(p2, p1);
        }
    }
    uint32_t nv = s.nv + 1;  // vertices on the path including vx
    if (!primary) {
        // :311-318 Russian roulette of the iteration that produced vx (its num_vertices == nv)
        if ((int)nv - 1 >= sc.options.rr_depth) {
            if (pcg_uniform(rng) > s.rr_prob) { cnt.finished++; s.rng_state = rng.state; return; }
            s.T = s.T / s.rr_prob;
        }
    }
    // loop condition of the next iteration (:66): num_vertices = nv + 1
    if (!(sc.options.max_depth == -1 || (int)nv + 1 <= sc.options.max_depth + 1)) {
        cnt.finished++; s.rng_state = rng.state; return;
    }
    cnt.bounces++;
    if (vx.material_id < 0) { cnt.finished++; s.rng_state = rng.state; return; }  // reference asserts (:165)
    const DevMaterial &mat = sc.materials[vx.material_id];
    // Lambertian fast path: the reflectance texture is fetched once per vertex instead of once per eval() call
    // (material.cpp evaluates it inside each of eval / pdf / sample; same values, lambertian.inl:1-50)
    // (CLASS == kMatLambert: the host checked that the scene has no other material -- the dispatchers are compiled out)
    const bool lambert = CLASS == kMatLambert || (CLASS != kMatDisney && mat.type == LJ_MAT_LAMBERTIAN);
    const V3 lambert_R = lambert ? mat_tex3(sc, mat, 0, vx) : mk3(0);
    MatCtx mc;
    mc.m = &mat;
    if (CLASS != kMatLambert) mc = mat_ctx<CLASS>(sc, mat, vx);

    // ---- next event estimation, :94-207
    float lu = pcg_uniform(rng), lv = pcg_uniform(rng);
    float light_w = pcg_uniform(rng), shape_w = pcg_uniform(rng);
    if (sc.num_lights > 0) {
        int light_id = sample_light(sc, light_w);
        const DevLight &light = sc.lights[light_id];
        PointAndNormal pl = sample_point_on_light(sc, light, vx.position, mk2(lu, lv), shape_w);
        V3 dir_light;
        float G, tfar;
        if (!light_is_envmap(light)) {
            dir_light = normalize(pl.position - vx.position);
            float dist = distance(pl.position, vx.position);
            tfar = (1 - sc.shadow_eps) * dist;
            G = fmaxf(-dot(dir_light, pl.normal), 0.f) / distance_squared(pl.position, vx.position);
        } else {
            dir_light = -pl.normal;
            tfar = LJ_INF;
            G = 1;
        }
        float p1 = light_pmf(sc, light_id) * pdf_point_on_light(sc, light, pl, vx.position);
        if (G > 0 && p1 > 0) {
            V3 f = lambert ? lambertian_eval(lambert_R, vx, dir_view, dir_light)
                           : (CALLS ? bsdf_eval_call(sc, mc, dir_view, dir_light, vx, 0) : bsdf_eval<DISPATCH>(sc, mc, dir_view, dir_light, vx, 0));
            V3 Le = light_emission(sc, light, -dir_light, 0.f, pl);
            float p2 = (lambert ? lambertian_pdf(vx, dir_view, dir_light)
                                : (CALLS ? bsdf_pdf_call(sc, mc, dir_view, dir_light, vx) : bsdf_pdf<DISPATCH>(sc, mc, dir_view, dir_light, vx))) * G;
            float w1 = mis_power(p1, p2);
            V3 c = s.T * (f * Le) * (G / p1 * w1);
            if (max3(c) > 0 || min3(c) < 0 || c.x != c.x || c.y != c.y || c.z != c.z) {
                s.sh_d = dir_light;
                s.sh_tfar = tfar;
                s.sh_c = c;
                cnt.shadow_rays++;
            }
        }
    }

    // ---- BSDF sampling, :209-259
    float bu = pcg_uniform(rng), bv = pcg_uniform(rng), bw = pcg_uniform(rng);
    s.rng_state = rng.state;
    BsdfSample bs;
    if (!(lambert ? lambertian_sample(vx, dir_view, mk2(bu, bv), bs)
                  : (CALLS ? bsdf_sample_call(sc, mc, dir_view, vx, mk2(bu, bv), bw, bs) : bsdf_sample<DISPATCH>(sc, mc, dir_view, vx, mk2(bu, bv), bw, bs)))) { cnt.finished++; return; }
    // ray_diff.radius stays 0 for the whole path upstream (only .spread is updated, :227-230)
    if (bs.eta == 0) {
        s.spread = spread_reflect(0.f, s.spread, vx.mean_curvature, bs.roughness);
    } else {
        s.spread = spread_refract(0.f, s.spread, vx.mean_curvature, bs.eta, bs.roughness);
        s.eta_scale /= (bs.eta * bs.eta);
    }
    V3 f = lambert ? lambertian_eval(lambert_R, vx, dir_view, bs.dir_out)
                   : (CALLS ? bsdf_eval_call(sc, mc, dir_view, bs.dir_out, vx, 0) : bsdf_eval<DISPATCH>(sc, mc, dir_view, bs.dir_out, vx, 0));
    float p2 = lambert ? lambertian_pdf(vx, dir_view, bs.dir_out)
                       : (CALLS ? bsdf_pdf_call(sc, mc, dir_view, bs.dir_out, vx) : bsdf_pdf<DISPATCH>(sc, mc, dir_view, bs.dir_out, vx));
    if (!(p2 > 0)) { cnt.finished++; return; }
    s.rr_prob = fminf(max3(s.T) / s.eta_scale, 0.95f);  // :313, evaluated with the pre-update throughput
    s.T = s.T * f / p2;
    s.pdf_sa = p2;
    s.o = vx.position;
    s.tnear = sc.isect_eps;
    s.d = bs.dir_out;
    s.tfar = LJ_INF;
    s.nv = nv;
    s.flags |= kAlive;
    cnt.extend_rays++;
}

// ---- SoA load / store -------------------------------------------------------------------------
LJ_HD void load_state(const PathPool &p, int i, PathState &s) {
    V4 a = p.ray_o[i], b = p.ray_d[i], h = p.hit[i], t = p.thr[i], r = p.rad[i], m = p.meta[i], x = p.aux[i];
    s.o = xyz(a); s.tnear = a.w;
    s.d = xyz(b); s.tfar = b.w;
    s.hit.t = h.x; s.hit.u = h.y; s.hit.v = h.z; s.hit.prim = (int)f2u(h.w);
    s.T = xyz(t); s.pdf_sa = t.w;
    s.L = xyz(r); s.rr_prob = r.w;
    s.pixel = f2u(m.x);
    uint32_t nf = f2u(m.y);
    s.nv = nf & 0xffffu;
    s.flags = nf & 0xffff0000u;
    s.rng_state = (uint64_t)f2u(m.z) | ((uint64_t)f2u(m.w) << 32);
    s.eta_scale = x.x; s.spread = x.y; s.sample = f2u(x.z); s.medium = (int)f2u(x.w);
    s.sh_tfar = -1; s.sh_d = mk3(0); s.sh_c = mk3(0);
}
LJ_HD void store_state(const PathPool &p, int i, const PathState &s, bool store_ray) {
    if (store_ray) {
        p.ray_o[i] = mk4(s.o, s.tnear);
        p.ray_d[i] = mk4(s.d, s.tfar);
    }
    p.thr[i] = mk4(s.T, s.pdf_sa);
    p.rad[i] = mk4(s.L, s.rr_prob);
    p.sh_d[i] = mk4(s.sh_d, s.sh_tfar);
    if (s.sh_tfar >= 0) p.sh_c[i] = mk4(s.sh_c, 0.f);
    p.meta[i] = mk4(u2f(s.pixel), u2f((s.nv & 0xffffu) | s.flags), u2f((uint32_t)s.rng_state), u2f((uint32_t)(s.rng_state >> 32)));
    p.aux[i] = mk4(s.eta_scale, s.spread, u2f(s.sample), u2f((uint32_t)s.medium));
}

// volpath: the common fields plus the MIS caches and the NEE walk record
LJ_HD void load_state_vol(const PathPool &p, int i, PathState &s) {
    load_state(p, i, s);
    s.mt_dir = xyz(p.vol0[i]); s.mt_nee = xyz(p.vol1[i]); s.nee_p = xyz(p.vol2[i]);
    s.sh_pdf_dir = -1;
}
LJ_HD void store_state_vol(const PathPool &p, int i, const PathState &s, bool store_ray) {
    if (store_ray) {
        p.ray_o[i] = mk4(s.o, s.tnear);
        p.ray_d[i] = mk4(s.d, s.tfar);
    }
    p.thr[i] = mk4(s.T, s.pdf_sa);
    p.rad[i] = mk4(s.L, s.rr_prob);
    p.meta[i] = mk4(u2f(s.pixel), u2f((s.nv & 0xffffu) | s.flags), u2f((uint32_t)s.rng_state), u2f((uint32_t)(s.rng_state >> 32)));
    p.aux[i] = mk4(s.eta_scale, s.spread, u2f(s.sample), u2f((uint32_t)s.medium));
    p.vol0[i] = mk4(s.mt_dir, 0.f); p.vol1[i] = mk4(s.mt_nee, 0.f); p.vol2[i] = mk4(s.nee_p, 0.f);
    p.sh_d[i] = mk4(s.sh_d, s.sh_pdf_dir);
    if (s.sh_pdf_dir >= 0) {
        p.sh_c[i] = mk4(s.sh_c, s.sh_pdf_nee);
        p.sh_o[i] = mk4(s.sh_o, u2f((uint32_t)((s.sh_medium + 1) & 0xffff) | (s.sh_budget << 16)));
        p.sh_pl[i] = mk4(s.sh_pl, u2f(s.sh_seed));
    }
}

}  // namespace lj
