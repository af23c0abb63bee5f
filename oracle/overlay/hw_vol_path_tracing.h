// oracle/overlay/hw_vol_path_tracing.h -- TEST INFRASTRUCTURE (CPU oracle), never shipped.
//
// The public reference ships vol_path_tracing_1..5 / vol_path_tracing as homework stubs that return
// zero (src/vol_path_tracing.h:6-64).  This overlay restates the final volumetric path tracer of the
// course handout handouts/homework2.tex in the reference's own style (double precision, the reference's
// intersect / emission / light / medium / phase-function API), so the otherwise unmodified reference
// renderer can produce images for the volpath configs.  PARITY UNPINNED: no reference code, test or
// golden vector pins it; the external pins are the handout pseudo-code cited below and its LDR figures.
//
// One estimator serves every `version`: the earlier versions of the handout are restrictions of the
// final one and have the same expectation on the scenes written for them (v1: sigma_s = 0, maxDepth 1;
// v2: maxDepth 2; v3: no NEE, same integral; v4/v5: monochromatic homogeneous media).
//   main loop                      homework2.tex:341-398
//   update_medium                  :401-410
//   NEE through index-matched
//   surfaces                       :459-510, MIS caches :521-558
//   chromatic delta tracking       :713-758, ratio tracking :771-810, throughput update :814-816
// Substituted for src/vol_path_tracing.h in the generated oracle/_ref/gen/render_hw.cpp.
#pragma once

namespace hwvol {

inline int update_medium(const PathVertex &isect, const Vector3 &dir, int medium) {
    if (isect.interior_medium_id != isect.exterior_medium_id) {
        medium = dot(dir, isect.geometric_normal) > 0 ? isect.exterior_medium_id : isect.interior_medium_id;
    }
    return medium;
}

inline Spectrum vexp(const Spectrum &s) { return Spectrum{exp(s.x), exp(s.y), exp(s.z)}; }
inline Real vavg(const Spectrum &s) { return (s.x + s.y + s.z) / 3; }

// Free flight over [0, t_hit] by chromatic delta tracking.  Returns true on a real collision at accum_t.
inline bool free_flight(const Scene &scene, const Medium &medium, const Ray &ray, Real t_hit, pcg32_state &rng,
                        Spectrum &transmittance, Spectrum &trans_dir_pdf, Spectrum &trans_nee_pdf, Real &accum_t) {
    Spectrum majorant = get_majorant(medium, ray);
    Real u = next_pcg32_real<Real>(rng);
    int channel = std::clamp(int(u * 3), 0, 2);
    Real max_maj = max(majorant);
    accum_t = 0;
    int iteration = 0;
    while (true) {
        if (majorant[channel] <= 0) break;
        if (iteration >= scene.options.max_null_collisions) break;
        Real t = -log(1 - next_pcg32_real<Real>(rng)) / majorant[channel];
        Real dt = t_hit - accum_t;
        accum_t = min(accum_t + t, t_hit);
        if (t < dt) {
            Vector3 p = ray.org + accum_t * ray.dir;
            Spectrum sigma_t = get_sigma_a(medium, p) + get_sigma_s(medium, p);
            Spectrum real_prob = sigma_t / majorant;
            Spectrum e = vexp(-majorant * t);
            if (next_pcg32_real<Real>(rng) < real_prob[channel]) {
                transmittance *= e / max_maj;
                trans_dir_pdf *= e * majorant * real_prob / max_maj;
                return true;
            }
            Spectrum sigma_n = majorant - sigma_t;
            transmittance *= e * sigma_n / max_maj;
            trans_dir_pdf *= e * majorant * (Real(1) - real_prob) / max_maj;
            trans_nee_pdf *= e * majorant / max_maj;
        } else {
            Spectrum e = vexp(-majorant * dt);
            transmittance *= e;
            trans_dir_pdf *= e;
            trans_nee_pdf *= e;
            break;
        }
        iteration++;
    }
    return false;
}

// Ratio tracking over one shadow segment of length next_t.
inline void ratio_track(const Scene &scene, const Medium &medium, const Ray &ray, Real next_t, pcg32_state &rng,
                        Spectrum &T_light, Spectrum &p_trans_nee, Spectrum &p_trans_dir) {
    Spectrum majorant = get_majorant(medium, ray);
    Real u = next_pcg32_real<Real>(rng);
    int channel = std::clamp(int(u * 3), 0, 2);
    Real max_maj = max(majorant);
    Real accum_t = 0;
    int iteration = 0;
    while (true) {
        if (majorant[channel] <= 0) break;
        if (iteration >= scene.options.max_null_collisions) break;
        Real t = -log(1 - next_pcg32_real<Real>(rng)) / majorant[channel];
        Real dt = next_t - accum_t;
        accum_t = min(accum_t + t, next_t);
        if (t < dt) {
            Vector3 p = ray.org + accum_t * ray.dir;
            Spectrum sigma_t = get_sigma_a(medium, p) + get_sigma_s(medium, p);
            Spectrum sigma_n = majorant - sigma_t;
            Spectrum real_prob = sigma_t / majorant;
            Spectrum e = vexp(-majorant * t);
            T_light *= e * sigma_n / max_maj;
            p_trans_nee *= e * majorant / max_maj;
            p_trans_dir *= e * majorant * (Real(1) - real_prob) / max_maj;
            if (max(T_light) <= 0) break;
        } else {
            Spectrum e = vexp(-majorant * dt);
            T_light *= e;
            p_trans_nee *= e;
            p_trans_dir *= e;
            break;
        }
        iteration++;
    }
}

inline Spectrum L(const Scene &scene, int x, int y, pcg32_state &rng) {
    int w = scene.camera.width, h = scene.camera.height;
    Vector2 screen_pos((x + next_pcg32_real<Real>(rng)) / w, (y + next_pcg32_real<Real>(rng)) / h);
    Ray ray = sample_primary(scene.camera, screen_pos);
    RayDifferential ray_diff = RayDifferential{Real(0), Real(0)};  // homework2.tex:191-194
    int current_medium = scene.camera.medium_id;

    Spectrum throughput = make_const_spectrum(1);
    Spectrum radiance = make_zero_spectrum();
    int bounces = 0;
    Real dir_pdf = 0;
    Vector3 nee_p_cache{0, 0, 0};
    Spectrum multi_trans_dir_pdf = make_const_spectrum(1);
    Spectrum multi_trans_nee_pdf = make_const_spectrum(1);
    bool never_scatter = true;
    Real eta_scale = 1;
    const int max_depth = scene.options.max_depth;
    const int rr_depth = scene.options.rr_depth;

    while (true) {
        bool scatter = false;
        std::optional<PathVertex> vertex_ = intersect(scene, ray, ray_diff);
        Real t_hit = vertex_ ? distance(ray.org, vertex_->position) : infinity<Real>();
        Spectrum transmittance = make_const_spectrum(1);
        Spectrum trans_dir_pdf = make_const_spectrum(1), trans_nee_pdf = make_const_spectrum(1);
        Vector3 p = vertex_ ? vertex_->position : ray.org;
        if (current_medium >= 0) {
            Real accum_t = 0;
            scatter = free_flight(scene, scene.media[current_medium], ray, t_hit, rng,
                                  transmittance, trans_dir_pdf, trans_nee_pdf, accum_t);
            if (scatter) p = ray.org + accum_t * ray.dir;
        }
        throughput *= transmittance / vavg(trans_dir_pdf);
        multi_trans_dir_pdf *= trans_dir_pdf;
        multi_trans_nee_pdf *= trans_nee_pdf;
        if (!scatter && !vertex_) break;  // left the scene; no environment maps in volpath (homework2.tex:196)

        // ---- emission of the surface we reached, MIS against NEE (:538-558)
        if (!scatter && is_light(scene.shapes[vertex_->shape_id])) {
            Spectrum Le = emission(*vertex_, -ray.dir, scene);
            if (never_scatter) {
                radiance += throughput * Le;
            } else {
                int light_id = get_area_light_id(scene.shapes[vertex_->shape_id]);
                PointAndNormal light_point{vertex_->position, vertex_->geometric_normal};
                Real pdf_nee = light_pmf(scene, light_id) *
                               pdf_point_on_light(scene.lights[light_id], light_point, nee_p_cache, scene) *
                               vavg(multi_trans_nee_pdf);
                Real G = fabs(dot(ray.dir, vertex_->geometric_normal)) / distance_squared(nee_p_cache, vertex_->position);
                Real pdf_dir = dir_pdf * vavg(multi_trans_dir_pdf) * G;
                Real wgt = (pdf_dir * pdf_dir) / (pdf_dir * pdf_dir + pdf_nee * pdf_nee);
                radiance += throughput * Le * wgt;
            }
        }
        if (max_depth != -1 && bounces == max_depth - 1) break;

        // ---- index-matched surface: pass through (:370-375)
        if (!scatter && vertex_->material_id == -1) {
            current_medium = update_medium(*vertex_, ray.dir, current_medium);
            bounces++;
            ray = Ray{vertex_->position, ray.dir, get_intersection_epsilon(scene), infinity<Real>()};
            continue;
        }

        // ---- next event estimation from p (:459-510, :771-810)
        const Vector3 dir_view = -ray.dir;
        const Medium *medium = current_medium >= 0 ? &scene.media[current_medium] : nullptr;
        Spectrum sigma_s = scatter ? get_sigma_s(*medium, p) : make_const_spectrum(1);
        {
            Vector2 light_uv{next_pcg32_real<Real>(rng), next_pcg32_real<Real>(rng)};
            Real light_w = next_pcg32_real<Real>(rng);
            Real shape_w = next_pcg32_real<Real>(rng);
            int light_id = sample_light(scene, light_w);
            const Light &light = scene.lights[light_id];
            PointAndNormal pl = sample_point_on_light(light, p, light_uv, shape_w, scene);
            Vector3 dir_light = normalize(pl.position - p);
            Spectrum T_light = make_const_spectrum(1), p_trans_nee = make_const_spectrum(1), p_trans_dir = make_const_spectrum(1);
            int shadow_medium = scatter ? current_medium : update_medium(*vertex_, dir_light, current_medium);
            int shadow_bounces = 0;
            Vector3 pc = p;
            bool blocked = false;
            while (true) {
                Real dist_light = distance(pc, pl.position);
                Ray shadow_ray{pc, dir_light, get_shadow_epsilon(scene), (1 - get_shadow_epsilon(scene)) * dist_light};
                std::optional<PathVertex> isect = intersect(scene, shadow_ray, RayDifferential{Real(0), Real(0)});
                Real next_t = isect ? distance(pc, isect->position) : dist_light;
                if (shadow_medium >= 0) {
                    ratio_track(scene, scene.media[shadow_medium], shadow_ray, next_t, rng, T_light, p_trans_nee, p_trans_dir);
                }
                if (!isect) break;
                if (isect->material_id >= 0) { blocked = true; break; }
                shadow_bounces++;
                if (max_depth != -1 && bounces + shadow_bounces + 1 >= max_depth) { blocked = true; break; }
                shadow_medium = update_medium(*isect, dir_light, shadow_medium);
                pc = isect->position;
            }
            if (!blocked && max(T_light) > 0) {
                Real G = max(-dot(dir_light, pl.normal), Real(0)) / distance_squared(pl.position, p);
                Real pdf_nee = light_pmf(scene, light_id) * pdf_point_on_light(light, pl, p, scene) * vavg(p_trans_nee);
                if (G > 0 && pdf_nee > 0) {
                    Spectrum Le = emission(light, -dir_light, Real(0), pl, scene);
                    Spectrum f;
                    Real pdf_scatter;
                    if (scatter) {
                        PhaseFunction phase = get_phase_function(*medium);
                        f = eval(phase, dir_view, dir_light) * sigma_s;
                        pdf_scatter = pdf_sample_phase(phase, dir_view, dir_light);
                    } else {
                        const Material &mat = scene.materials[vertex_->material_id];
                        f = eval(mat, dir_view, dir_light, *vertex_, scene.texture_pool);
                        pdf_scatter = pdf_sample_bsdf(mat, dir_view, dir_light, *vertex_, scene.texture_pool);
                    }
                    Real pdf_dir = pdf_scatter * G * vavg(p_trans_dir);
                    Real wgt = (pdf_nee * pdf_nee) / (pdf_nee * pdf_nee + pdf_dir * pdf_dir);
                    radiance += throughput * T_light * f * Le * (G / pdf_nee * wgt);
                }
            }
        }

        // ---- sample the next direction
        Vector3 next_dir;
        if (scatter) {
            PhaseFunction phase = get_phase_function(*medium);
            Vector2 rnd{next_pcg32_real<Real>(rng), next_pcg32_real<Real>(rng)};
            std::optional<Vector3> d = sample_phase_function(phase, dir_view, rnd);
            if (!d) break;
            next_dir = *d;
            Real pdf = pdf_sample_phase(phase, dir_view, next_dir);
            if (pdf <= 0) break;
            throughput *= eval(phase, dir_view, next_dir) * sigma_s / pdf;
            dir_pdf = pdf;
            ray = Ray{p, next_dir, Real(0), infinity<Real>()};
        } else {
            const Material &mat = scene.materials[vertex_->material_id];
            Vector2 rnd{next_pcg32_real<Real>(rng), next_pcg32_real<Real>(rng)};
            Real rnd_w = next_pcg32_real<Real>(rng);
            std::optional<BSDFSampleRecord> bs = sample_bsdf(mat, dir_view, *vertex_, scene.texture_pool, rnd, rnd_w);
            if (!bs) break;
            next_dir = bs->dir_out;
            if (bs->eta != 0) eta_scale /= (bs->eta * bs->eta);
            Spectrum f = eval(mat, dir_view, next_dir, *vertex_, scene.texture_pool);
            Real pdf = pdf_sample_bsdf(mat, dir_view, next_dir, *vertex_, scene.texture_pool);
            if (pdf <= 0) break;
            throughput *= f / pdf;
            dir_pdf = pdf;
            current_medium = update_medium(*vertex_, next_dir, current_medium);
            ray = Ray{p, next_dir, get_intersection_epsilon(scene), infinity<Real>()};
        }
        never_scatter = false;
        nee_p_cache = p;
        multi_trans_dir_pdf = make_const_spectrum(1);
        multi_trans_nee_pdf = make_const_spectrum(1);

        // ---- Russian roulette (:388-396, with the eta-aware probability of path_tracing.h:311-318)
        if (bounces >= rr_depth) {
            Real rr_prob = min(max((1 / eta_scale) * throughput), Real(0.95));
            if (next_pcg32_real<Real>(rng) > rr_prob) break;
            throughput /= rr_prob;
        }
        bounces++;
    }
    return radiance;
}

}  // namespace hwvol

Spectrum vol_path_tracing_1(const Scene &scene, int x, int y, pcg32_state &rng) { return hwvol::L(scene, x, y, rng); }
Spectrum vol_path_tracing_2(const Scene &scene, int x, int y, pcg32_state &rng) { return hwvol::L(scene, x, y, rng); }
Spectrum vol_path_tracing_3(const Scene &scene, int x, int y, pcg32_state &rng) { return hwvol::L(scene, x, y, rng); }
Spectrum vol_path_tracing_4(const Scene &scene, int x, int y, pcg32_state &rng) { return hwvol::L(scene, x, y, rng); }
Spectrum vol_path_tracing_5(const Scene &scene, int x, int y, pcg32_state &rng) { return hwvol::L(scene, x, y, rng); }
Spectrum vol_path_tracing(const Scene &scene, int x, int y, pcg32_state &rng) { return hwvol::L(scene, x, y, rng); }
