// lj_volpath.h -- the volumetric radiance estimator in wavefront form.
// The public reference ships vol_path_tracing_1..5 / vol_path_tracing as stubs
// (src/vol_path_tracing.h:6-64); the estimator is specified by handouts/homework2.tex: main loop
// :341-398, update_medium :401-410, NEE through index-matched surfaces :459-510, MIS caches :521-558,
// chromatic delta tracking :713-758, ratio tracking :771-810, throughput update :814-816.  One call of
// vol_path_tracing() becomes generate -> [extend -> shade_vol -> walk]*:
//   shade_vol  free flight to the surface the extension ray found (or a scattering event before it),
//              MIS-weighted emission, index-matched pass-through, NEE sample, phase / BSDF sample, RR
//   walk       the NEE segment walk (lj::nee_walk_*): closest-hit segments through index-matched surfaces with
//              ratio tracking in between; runs inside the persistent traversal kernel (wavefront.cu)
// As in the reference's render.cpp:111-124 every `version` selects an estimator with the same
// expectation on the scenes written for it; the final one serves all of them.
#pragma once
#include "lj_media.h"
#include "lj_path.h"

namespace lj {

// homework2.tex:401-410
LJ_HD int update_medium(int interior, int exterior, V3 dir, V3 geometric_normal, int medium) {
    if (interior != exterior) medium = dot(dir, geometric_normal) > 0 ? exterior : interior;
    return medium;
}

constexpr uint32_t kWalkNoBudget = 0xffffu;
constexpr int kScatter = -2;  // hit.prim of a path whose free flight ended in a real collision (hit.t = its distance)

// Free flight of one path in whole-loop form (host simulation of k_flight and reference for its stepwise form):
// updates throughput and the MIS caches, records a scattering event in s.hit.
LJ_HD void flight_finish(PathState &s, bool scatter, float accum_t, V3 transmittance, V3 trans_dir_pdf, V3 trans_nee_pdf) {
    s.T = s.T * transmittance / avg3(trans_dir_pdf);
    s.mt_dir = s.mt_dir * trans_dir_pdf;
    s.mt_nee = s.mt_nee * trans_nee_pdf;
    if (scatter) { s.hit.t = accum_t; s.hit.u = 0; s.hit.v = 0; s.hit.prim = kScatter; }
}

// One shade step of the volumetric path tracer.  On return: s.flags has kAlive iff an extension ray was
// written; s.sh_pdf_dir >= 0 iff an NEE walk was written.
LJ_HD void shade_vol_path(const DevScene &sc, const RenderParams &rp, PathState &s, ShadeCounters &cnt) {
    Pcg rng = path_rng(s, rp);
    const bool never_scatter = s.pdf_sa < 0;
    const int max_depth = sc.options.max_depth;
    const int bounces = (int)s.nv - 1;  // generate_path starts nv at 1
    s.sh_pdf_dir = -1;
    s.flags &= ~kAlive;
    // The free flight over this ray (chromatic delta tracking, :713-758) and the throughput update (:814-816) were
    // done by k_flight (flight_path below): a real collision is recorded as hit.prim == kScatter at distance hit.t.
    const bool scatter = s.hit.prim == kScatter;
    const bool has_hit = s.hit.prim >= 0;
    const float accum_t = s.hit.t;
    if (!scatter && !has_hit) { cnt.finished++; s.rng_state = rng.state; return; }  // no envmaps in volpath (:196)

    Vertex vx;
    V3 p;
    if (scatter) {
        p = s.o + s.d * accum_t;
        vx.material_id = -1;
    } else {
        vx = make_vertex(sc, s.o, s.d, s.hit, 0.f, 0.f);  // ray differentials are off (:191-194)
        p = vx.position;
        // ---- emission of the surface reached, MIS against NEE (:538-558)
        const DevShape &shape = sc.shapes[vx.shape_id];
        if (shape.area_light_id >= 0) {
            V3 Le = vertex_emission(sc, vx, -s.d);
            if (never_scatter) {
                s.L += s.T * Le;
            } else {
                PointAndNormal lp;
                lp.position = vx.position;
                lp.normal = vx.geometric_normal;
                float pdf_nee = light_pmf(sc, shape.area_light_id) *
                                pdf_point_on_light(sc, sc.lights[shape.area_light_id], lp, s.nee_p) * avg3(s.mt_nee);
                float G = fabsf(dot(s.d, vx.geometric_normal)) / distance_squared(s.nee_p, vx.position);
                float pdf_dir = s.pdf_sa * avg3(s.mt_dir) * G;
                s.L += s.T * Le * mis_power(pdf_dir, pdf_nee);
            }
        }
    }
    if (max_depth != -1 && bounces == max_depth - 1) { cnt.finished++; s.rng_state = rng.state; return; }

    // ---- index-matched surface: pass through (:370-375)
    if (!scatter && vx.material_id == -1) {
        s.medium = update_medium(vx.interior_medium_id, vx.exterior_medium_id, s.d, vx.geometric_normal, s.medium);
        s.nv = (uint32_t)(bounces + 2);
        s.o = vx.position;
        s.tnear = sc.isect_eps;
        s.tfar = LJ_INF;
        s.flags |= kAlive;
        s.rng_state = rng.state;
        cnt.bounces++;
        cnt.extend_rays++;
        return;
    }
    cnt.bounces++;

    // ---- next event estimation from p (:459-510): everything but the segment walk
    const V3 dir_view = -s.d;
    V3 sigma_s = mk3(1);
    if (scatter) {
        V3 sa;
        medium_sigmas(sc.media[s.medium], p, sa, sigma_s);
    }
    // (one material context per vertex -- the principled BSDF's parameter textures are read once -- and out-of-line
    //  dispatchers: lj_materials.h)
    MatCtx mc;
    mc.m = nullptr;
    if (!scatter) mc = mat_ctx_all_call(sc, sc.materials[vx.material_id], vx);
    {
        float lu = pcg_uniform(rng), lv = pcg_uniform(rng);
        float light_w = pcg_uniform(rng), shape_w = pcg_uniform(rng);
        uint32_t walk_seed = pcg_next(rng);
        if (sc.num_lights > 0) {
            int light_id = sample_light(sc, light_w);
            const DevLight &light = sc.lights[light_id];
            PointAndNormal pl = sample_point_on_light(sc, light, p, mk2(lu, lv), shape_w);
            V3 dir_light = normalize(pl.position - p);
            float G = fmaxf(-dot(dir_light, pl.normal), 0.f) / distance_squared(pl.position, p);
            float pdf_nee = light_pmf(sc, light_id) * pdf_point_on_light(sc, light, pl, p);
            if (G > 0 && pdf_nee > 0) {
                V3 f;
                float pdf_scatter;
                if (scatter) {
                    const DevMedium &m = sc.media[s.medium];
                    f = sigma_s * phase_eval(m, dir_view, dir_light);
                    pdf_scatter = phase_pdf(m, dir_view, dir_light);
                } else {
                    f = bsdf_eval_all_call(sc, mc, dir_view, dir_light, vx, 0);
                    pdf_scatter = bsdf_pdf_all_call(sc, mc, dir_view, dir_light, vx);
                }
                V3 Le = light_emission(sc, light, -dir_light, 0.f, pl);
                V3 c = s.T * f * Le * (G / pdf_nee);
                if (max3(c) > 0 || min3(c) < 0 || c.x != c.x || c.y != c.y || c.z != c.z) {
                    s.sh_o = p;
                    s.sh_medium = scatter ? s.medium
                                          : update_medium(vx.interior_medium_id, vx.exterior_medium_id, dir_light, vx.geometric_normal, s.medium);
                    // the walk is blocked once bounces + shadow_bounces + 1 >= max_depth (:485-489)
                    int budget = max_depth == -1 ? (int)kWalkNoBudget : max_depth - bounces - 1;
                    s.sh_budget = (uint32_t)clampi(budget, 0, (int)kWalkNoBudget);
                    s.sh_d = dir_light;
                    s.sh_pl = pl.position;
                    s.sh_seed = walk_seed;
                    s.sh_c = c;
                    s.sh_pdf_nee = pdf_nee;
                    s.sh_pdf_dir = pdf_scatter * G;
                    cnt.shadow_rays++;
                }
            }
        }
    }

    // ---- sample the next direction (:376-387)
    V3 next_dir;
    if (scatter) {
        const DevMedium &m = sc.media[s.medium];
        float u0 = pcg_uniform(rng), u1 = pcg_uniform(rng);
        next_dir = phase_sample(m, dir_view, mk2(u0, u1));
        float pdf = phase_pdf(m, dir_view, next_dir);
        if (!(pdf > 0)) { cnt.finished++; s.rng_state = rng.state; return; }
        s.T = s.T * (sigma_s * (phase_eval(m, dir_view, next_dir) / pdf));
        s.pdf_sa = pdf;
        s.tnear = 0;
    } else {
        float bu = pcg_uniform(rng), bv = pcg_uniform(rng), bw = pcg_uniform(rng);
        BsdfSample bs;
        if (!bsdf_sample_all_call(sc, mc, dir_view, vx, mk2(bu, bv), bw, bs)) { cnt.finished++; s.rng_state = rng.state; return; }
        next_dir = bs.dir_out;
        if (bs.eta != 0) s.eta_scale /= (bs.eta * bs.eta);
        V3 f = bsdf_eval_all_call(sc, mc, dir_view, next_dir, vx, 0);
        float pdf = bsdf_pdf_all_call(sc, mc, dir_view, next_dir, vx);
        if (!(pdf > 0)) { cnt.finished++; s.rng_state = rng.state; return; }
        s.T = s.T * f / pdf;
        s.pdf_sa = pdf;
        s.medium = update_medium(vx.interior_medium_id, vx.exterior_medium_id, next_dir, vx.geometric_normal, s.medium);
        s.tnear = sc.isect_eps;
    }
    s.nee_p = p;
    s.mt_dir = mk3(1);
    s.mt_nee = mk3(1);

    // ---- Russian roulette (:388-396, eta-aware like path_tracing.h:311-318)
    if (bounces >= sc.options.rr_depth) {
        float rr_prob = fminf(max3(s.T) / s.eta_scale, 0.95f);
        if (pcg_uniform(rng) > rr_prob) { cnt.finished++; s.rng_state = rng.state; return; }
        s.T = s.T / rr_prob;
    }
    s.rng_state = rng.state;
    s.nv = (uint32_t)(bounces + 2);
    s.o = p;
    s.d = next_dir;
    s.tfar = LJ_INF;
    s.flags |= kAlive;
    cnt.extend_rays++;
}

// ---- the NEE segment walk (:459-510 with :771-810) ---------------------------------------------
// State a lane of the walk kernel keeps in registers between segments.
struct NeeWalk {
    V3 pc, dir, pl;       // current origin, direction to the light point, the light point
    V3 T_light, p_nee, p_dir;
    V3 c;                 // throughput * f * Le * G / pdf_nee (without transmittance)
    float pdf_nee, pdf_dir;
    int medium;
    uint32_t budget, shadow_bounces;
    Pcg rng;
};

// The walk's own PCG stream: a function of the PATH (pixel * spp + sample) and of one draw of the path's stream,
// never of the pool slot the path sits in -- renders are reproducible whatever the pool size or the split across GPUs.
LJ_HD Pcg walk_rng(uint64_t path_id, uint32_t walk_seed, uint64_t seed) {
    return pcg_init(path_stream(path_id) + (((uint64_t)walk_seed << 1) | 1ull), seed);
}

// tnear / tfar of the next segment
LJ_HD void nee_walk_segment(const DevScene &sc, const NeeWalk &w, float &tnear, float &tfar) {
    tnear = sc.shadow_eps;
    tfar = (1 - sc.shadow_eps) * distance_fixed(w.pc, w.pl);
}

// Geometric normal of a hit flipped to the shading-normal side, as intersect() leaves it
// (intersection.cpp:60-62), plus the ids update_medium needs.  Lighter than make_vertex: no frame, no uv.
LJ_HD void hit_medium_interface(const DevScene &sc, V3 org, V3 dir, const Hit &hit, V3 &ng, int &material_id, int &interior, int &exterior) {
    V4 pa = ld4(&sc.prims[hit.prim].a);
    V4 pc = ld4(&sc.prims[hit.prim].c);
    const DevShape &sh = sc.shapes[prim_shape_id(pc)];
    material_id = sh.material_id;
    interior = sh.interior_medium_id;
    exterior = sh.exterior_medium_id;
    if (prim_is_sphere(pc)) {
        ng = normalize((org + dir * hit.t) - xyz(pa));
        return;
    }
    V4 pb = ld4(&sc.prims[hit.prim].b);
    V3 A = xyz(pa), B = mk3(pa.w, pb.x, pb.y), C = mk3(pb.z, pb.w, pc.x);
    ng = normalize(cross(B - A, C - A));
    if (sh.has_normals) {
        const int *idx = sc.indices + 3 * (sh.tri_offset + prim_primitive_id(pc));
        V3 n0 = ld3(sc.normals, idx[0]), n1 = ld3(sc.normals, idx[1]), n2 = ld3(sc.normals, idx[2]);
        V3 sn = (1 - hit.u - hit.v) * n0 + hit.u * n1 + hit.v * n2;
        if (dot(ng, sn) < 0) ng = -ng;
    }
}

// Length of the segment just traversed (to the surface hit, or to the light point).
LJ_HD float nee_walk_next_t(const NeeWalk &w, const Hit &hit) { return hit.prim != kNoHit ? hit.t : distance_fixed(w.pc, w.pl); }

// After ratio tracking over the segment: opaque / index-matched test.  Returns true when the walk is over;
// `contribution` is then what the path's radiance gains (zero if blocked).
LJ_HD bool nee_walk_decide(const DevScene &sc, NeeWalk &w, const Hit &hit, V3 &contribution) {
    contribution = mk3(0);
    if (hit.prim != kNoHit) {
        V3 ng;
        int material_id, interior, exterior;
        hit_medium_interface(sc, w.pc, w.dir, hit, ng, material_id, interior, exterior);
        if (material_id >= 0) return true;  // blocked by an opaque surface
        w.shadow_bounces++;
        if (w.budget != kWalkNoBudget && w.shadow_bounces >= w.budget) return true;  // depth limit: treated as blocked
        if (!(max3(w.T_light) > 0)) return true;
        w.medium = update_medium(interior, exterior, w.dir, ng, w.medium);
        w.pc = w.pc + w.dir * hit.t;
        return false;
    }
    if (max3(w.T_light) > 0) {
        float pdf_nee = w.pdf_nee * avg3(w.p_nee);
        float pdf_dir = w.pdf_dir * avg3(w.p_dir);
        // (power heuristic with its sum of squares rounded in a fixed order, see dot_fixed)
        if (pdf_nee > 0) contribution = w.c * w.T_light * ((pdf_nee * pdf_nee) / sum_squares_fixed(pdf_nee, pdf_dir) / avg3(w.p_nee));
    }
    return true;
}

}  // namespace lj
