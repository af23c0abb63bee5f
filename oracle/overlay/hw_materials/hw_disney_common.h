// oracle/overlay/hw_materials/hw_disney_common.h -- TEST INFRASTRUCTURE (CPU oracle), never shipped.
//
// The public reference ships the six Disney materials as homework stubs that return zero
// (src/materials/disney_*.inl).  This overlay restates the lobes from the course handout
// handouts/homework1.tex in the reference's own style (double precision, reference helpers from
// microfacet.h / frame.h), so the unmodified reference renderer can produce images for the Disney
// configs.  PARITY UNPINNED: no reference code, test or golden vector pins these formulas; the only
// external pins are the handout equations cited per function and the handout's LDR figures.
// Included (through the generated oracle/_ref/gen/material_hw.cpp) after src/material.cpp's own
// definitions of eval_op / pdf_sample_bsdf_op / sample_bsdf_op and sample_cos_hemisphere.
#pragma once
#include "microfacet.h"

namespace hw {

inline Real sqr(Real x) { return x * x; }
inline Real pow5(Real x) { Real x2 = x * x; return x2 * x2 * x; }

// homework1.tex:204-210
inline void disney_alphas(Real roughness, Real anisotropic, Real &ax, Real &ay) {
    Real aspect = sqrt(1 - Real(0.9) * anisotropic);
    ax = fmax(Real(1e-4), roughness * roughness / aspect);
    ay = fmax(Real(1e-4), roughness * roughness * aspect);
}
// homework1.tex:197-201
inline Real ggx_aniso_D(const Vector3 &hl, Real ax, Real ay) {
    Real t = sqr(hl.x / ax) + sqr(hl.y / ay) + sqr(hl.z);
    return 1 / (c_PI * ax * ay * t * t);
}
// homework1.tex:213-220
inline Real smith_aniso_G1(const Vector3 &wl, Real ax, Real ay) {
    Real Lambda = (sqrt(1 + (sqr(wl.x * ax) + sqr(wl.y * ay)) / sqr(wl.z)) - 1) / 2;
    return 1 / (1 + Lambda);
}
// Heitz 2018 with two alphas (homework1.tex:251); same construction as microfacet.h:85-114.
inline Vector3 sample_visible_normals_aniso(const Vector3 &local_dir_in, Real ax, Real ay, const Vector2 &rnd) {
    if (local_dir_in.z < 0) return -sample_visible_normals_aniso(-local_dir_in, ax, ay, rnd);
    Vector3 hemi_dir_in = normalize(Vector3{ax * local_dir_in.x, ay * local_dir_in.y, local_dir_in.z});
    Real r = sqrt(rnd.x);
    Real phi = 2 * c_PI * rnd.y;
    Real t1 = r * cos(phi);
    Real t2 = r * sin(phi);
    Real s = (1 + hemi_dir_in.z) / 2;
    t2 = (1 - s) * sqrt(1 - t1 * t1) + s * t2;
    Vector3 disk_N{t1, t2, sqrt(max(Real(0), 1 - t1 * t1 - t2 * t2))};
    Frame hemi_frame(hemi_dir_in);
    Vector3 hemi_N = to_world(hemi_frame, disk_N);
    return normalize(Vector3{ax * hemi_N.x, ay * hemi_N.y, max(Real(0), hemi_N.z)});
}
inline Spectrum tint_of(const Spectrum &base) {  // homework1.tex:417
    Real l = luminance(base);
    return l > 0 ? base / l : make_const_spectrum(1);
}
inline bool below(const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {
    return dot(v.geometric_normal, wi) < 0 || dot(v.geometric_normal, wo) < 0;
}
inline Frame reflective_frame(const PathVertex &v, const Vector3 &wi) {
    Frame f = v.shading_frame;
    if (dot(f.n, wi) < 0) f = -f;
    return f;
}
inline Frame transmissive_frame(const PathVertex &v, const Vector3 &wi) {
    Frame f = v.shading_frame;
    if (dot(f.n, wi) * dot(v.geometric_normal, wi) < 0) f = -f;
    return f;
}

// ---- diffuse + subsurface, homework1.tex:100-135
inline Spectrum diffuse_eval(const Spectrum &base, Real roughness, Real subsurface, const PathVertex &v,
                             const Vector3 &wi, const Vector3 &wo) {
    if (below(v, wi, wo)) return make_zero_spectrum();
    Frame f = reflective_frame(v, wi);
    Vector3 h = normalize(wi + wo);
    Real n_in = fabs(dot(f.n, wi)), n_out = fabs(dot(f.n, wo)), h_out = fabs(dot(h, wo));
    Real FD90 = Real(0.5) + 2 * roughness * h_out * h_out;
    Real FD_in = 1 + (FD90 - 1) * pow5(1 - n_in), FD_out = 1 + (FD90 - 1) * pow5(1 - n_out);
    Spectrum base_diffuse = base * (FD_in * FD_out * n_out / c_PI);
    Real FSS90 = roughness * h_out * h_out;
    Real FSS_in = 1 + (FSS90 - 1) * pow5(1 - n_in), FSS_out = 1 + (FSS90 - 1) * pow5(1 - n_out);
    Real denom = n_in + n_out;
    Real ss_term = denom > 0 ? (FSS_in * FSS_out * (1 / denom - Real(0.5)) + Real(0.5)) : Real(0.5);
    Spectrum ss = base * (Real(1.25) * ss_term * n_out / c_PI);
    return (1 - subsurface) * base_diffuse + subsurface * ss;
}
inline Real cosine_pdf(const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {  // lambertian.inl:21-35
    if (below(v, wi, wo)) return 0;
    Frame f = reflective_frame(v, wi);
    return fmax(dot(f.n, wo), Real(0)) / c_PI;
}
inline std::optional<BSDFSampleRecord> cosine_sample(const PathVertex &v, const Vector3 &wi, const Vector2 &rnd) {
    if (dot(v.geometric_normal, wi) < 0) return {};
    Frame f = reflective_frame(v, wi);
    return BSDFSampleRecord{to_world(f, sample_cos_hemisphere(rnd)), Real(0), Real(1)};  // lambertian.inl:37-50
}

// ---- metal, homework1.tex:182-222 (F0 supplied so the full BSDF can pass its modified Fresnel)
inline bool metal_geom(Real ax, Real ay, const PathVertex &v, const Vector3 &wi, const Vector3 &wo,
                       Vector3 &h, Real &D, Real &G_in, Real &G_out, Real &n_in) {
    if (below(v, wi, wo)) return false;
    Frame f = reflective_frame(v, wi);
    h = normalize(wi + wo);
    n_in = dot(f.n, wi);
    if (dot(f.n, wo) <= 0 || dot(f.n, h) <= 0 || n_in <= 0) return false;
    D = ggx_aniso_D(to_local(f, h), ax, ay);
    G_in = smith_aniso_G1(to_local(f, wi), ax, ay);
    G_out = smith_aniso_G1(to_local(f, wo), ax, ay);
    return true;
}
inline Spectrum metal_eval(const Spectrum &F0, Real ax, Real ay, const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {
    Vector3 h; Real D, Gi, Go, n_in;
    if (!metal_geom(ax, ay, v, wi, wo, h, D, Gi, Go, n_in)) return make_zero_spectrum();
    Spectrum F = schlick_fresnel(F0, fabs(dot(h, wo)));
    return F * (D * Gi * Go / (4 * fabs(n_in)));
}
inline Real metal_pdf(Real ax, Real ay, const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {
    Vector3 h; Real D, Gi, Go, n_in;
    if (!metal_geom(ax, ay, v, wi, wo, h, D, Gi, Go, n_in)) return 0;
    return D * Gi / (4 * fabs(n_in));
}
inline std::optional<BSDFSampleRecord> metal_sample(Real ax, Real ay, Real roughness, const PathVertex &v,
                                                    const Vector3 &wi, const Vector2 &rnd) {
    if (dot(v.geometric_normal, wi) < 0) return {};
    Frame f = reflective_frame(v, wi);
    Vector3 h = to_world(f, sample_visible_normals_aniso(to_local(f, wi), ax, ay, rnd));
    return BSDFSampleRecord{normalize(-wi + 2 * dot(wi, h) * h), Real(0), roughness};
}

// ---- clearcoat, homework1.tex:267-327
inline Real clearcoat_alpha(Real gloss) { return (1 - gloss) * Real(0.1) + gloss * Real(0.001); }
inline Real clearcoat_D(Real a2, Real hlz) { return (a2 - 1) / (c_PI * log(a2) * (1 + (a2 - 1) * hlz * hlz)); }
inline Real clearcoat_eval(Real gloss, const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {
    if (below(v, wi, wo)) return 0;
    Frame f = reflective_frame(v, wi);
    Vector3 h = normalize(wi + wo);
    Real n_in = dot(f.n, wi);
    if (dot(f.n, wo) <= 0 || dot(f.n, h) <= 0 || n_in <= 0) return 0;
    Real ag = clearcoat_alpha(gloss);
    Real Fc = schlick_fresnel(Real(0.04), fabs(dot(h, wo)));  // R0(eta = 1.5)
    Real Dc = clearcoat_D(ag * ag, dot(f.n, h));
    Real Gc = smith_aniso_G1(to_local(f, wi), Real(0.25), Real(0.25)) * smith_aniso_G1(to_local(f, wo), Real(0.25), Real(0.25));
    return Fc * Dc * Gc / (4 * fabs(n_in));
}
inline Real clearcoat_pdf(Real gloss, const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {
    if (below(v, wi, wo)) return 0;
    Frame f = reflective_frame(v, wi);
    Vector3 h = normalize(wi + wo);
    Real n_h = dot(f.n, h);
    if (dot(f.n, wo) <= 0 || n_h <= 0) return 0;
    Real ag = clearcoat_alpha(gloss);
    return clearcoat_D(ag * ag, n_h) * fabs(n_h) / (4 * fabs(dot(h, wo)));
}
inline std::optional<BSDFSampleRecord> clearcoat_sample(Real gloss, const PathVertex &v, const Vector3 &wi, const Vector2 &rnd) {
    if (dot(v.geometric_normal, wi) < 0) return {};
    Frame f = reflective_frame(v, wi);
    Real ag = clearcoat_alpha(gloss);
    Real a2 = ag * ag;
    Real cos_e = sqrt(std::clamp((1 - pow(a2, 1 - rnd.x)) / (1 - a2), Real(0), Real(1)));
    Real sin_e = sqrt(fmax(Real(0), 1 - cos_e * cos_e));
    Real az = 2 * c_PI * rnd.y;
    Vector3 h = to_world(f, Vector3{sin_e * cos(az), sin_e * sin(az), cos_e});
    return BSDFSampleRecord{normalize(-wi + 2 * dot(wi, h) * h), Real(0), sqrt(ag)};
}

// ---- glass, homework1.tex:343-366 (roughdielectric.inl with anisotropic D / G and tinted lobes)
inline Spectrum glass_eval(const Spectrum &Cr, const Spectrum &Ct, Real ax, Real ay, Real mat_eta, const PathVertex &v,
                           const Vector3 &wi, const Vector3 &wo, TransportDirection dir) {
    bool reflect = dot(v.geometric_normal, wi) * dot(v.geometric_normal, wo) > 0;
    Frame f = transmissive_frame(v, wi);
    Real eta = dot(v.geometric_normal, wi) > 0 ? mat_eta : 1 / mat_eta;
    Vector3 h = reflect ? normalize(wi + wo) : normalize(wi + wo * eta);
    if (dot(h, f.n) < 0) h = -h;
    Real h_dot_in = dot(h, wi);
    Real F = fresnel_dielectric(h_dot_in, eta);
    Real D = ggx_aniso_D(to_local(f, h), ax, ay);
    Real G = smith_aniso_G1(to_local(f, wi), ax, ay) * smith_aniso_G1(to_local(f, wo), ax, ay);
    if (reflect) return Cr * ((F * D * G) / (4 * fabs(dot(f.n, wi))));
    Real eta_factor = dir == TransportDirection::TO_LIGHT ? (1 / (eta * eta)) : 1;  // roughdielectric.inl:64
    Real h_dot_out = dot(h, wo);
    Real sqrt_denom = h_dot_in + eta * h_dot_out;
    return Ct * ((eta_factor * (1 - F) * D * G * eta * eta * fabs(h_dot_out * h_dot_in)) /
                 (fabs(dot(f.n, wi)) * sqrt_denom * sqrt_denom));
}
inline Real glass_pdf(Real ax, Real ay, Real mat_eta, const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {
    bool reflect = dot(v.geometric_normal, wi) * dot(v.geometric_normal, wo) > 0;
    Frame f = transmissive_frame(v, wi);
    Real eta = dot(v.geometric_normal, wi) > 0 ? mat_eta : 1 / mat_eta;
    Vector3 h = reflect ? normalize(wi + wo) : normalize(wi + wo * eta);
    if (dot(h, f.n) < 0) h = -h;
    Real h_dot_in = dot(h, wi);
    Real F = fresnel_dielectric(h_dot_in, eta);
    Real D = ggx_aniso_D(to_local(f, h), ax, ay);
    Real G_in = smith_aniso_G1(to_local(f, wi), ax, ay);
    if (reflect) return (F * D * G_in) / (4 * fabs(dot(f.n, wi)));
    Real h_dot_out = dot(h, wo);
    Real sqrt_denom = h_dot_in + eta * h_dot_out;
    Real dh_dout = eta * eta * h_dot_out / (sqrt_denom * sqrt_denom);
    return (1 - F) * D * G_in * fabs(dh_dout * h_dot_in / dot(f.n, wi));
}
inline std::optional<BSDFSampleRecord> glass_sample(Real ax, Real ay, Real roughness, Real mat_eta, const PathVertex &v,
                                                    const Vector3 &wi, const Vector2 &rnd, Real w) {
    Real eta = dot(v.geometric_normal, wi) > 0 ? mat_eta : 1 / mat_eta;
    Frame f = transmissive_frame(v, wi);
    Vector3 h = to_world(f, sample_visible_normals_aniso(to_local(f, wi), ax, ay, rnd));
    if (dot(h, f.n) < 0) h = -h;
    Real h_dot_in = dot(h, wi);
    Real F = fresnel_dielectric(h_dot_in, eta);
    if (w <= F) return BSDFSampleRecord{normalize(-wi + 2 * dot(wi, h) * h), Real(0), roughness};
    Real h_dot_out_sq = 1 - (1 - h_dot_in * h_dot_in) / (eta * eta);
    if (h_dot_out_sq <= 0) return {};
    if (h_dot_in < 0) h = -h;
    Real h_dot_out = sqrt(h_dot_out_sq);
    return BSDFSampleRecord{-wi / eta + (fabs(h_dot_in) / eta - h_dot_out) * h, eta, roughness};
}

// ---- sheen, homework1.tex:412-425
inline Spectrum sheen_eval(const Spectrum &base, Real sheen_tint, const PathVertex &v, const Vector3 &wi, const Vector3 &wo) {
    if (below(v, wi, wo)) return make_zero_spectrum();
    Frame f = reflective_frame(v, wi);
    Vector3 h = normalize(wi + wo);
    Spectrum Csheen = make_const_spectrum(1 - sheen_tint) + sheen_tint * tint_of(base);
    return Csheen * (pow5(1 - fabs(dot(h, wo))) * fabs(dot(f.n, wo)));
}

// ---- the full BSDF, homework1.tex:461-548
struct Params {
    Spectrum base;
    Real st, metallic, subsurface, specular, roughness, stint, aniso, sheen, sheen_tint, cc, cc_gloss, eta;
};
inline Params params_of(const DisneyBSDF &b, const PathVertex &v, const TexturePool &pool) {
    Params p;
    p.base = eval(b.base_color, v.uv, v.uv_screen_size, pool);
    p.st = eval(b.specular_transmission, v.uv, v.uv_screen_size, pool);
    p.metallic = eval(b.metallic, v.uv, v.uv_screen_size, pool);
    p.subsurface = eval(b.subsurface, v.uv, v.uv_screen_size, pool);
    p.specular = eval(b.specular, v.uv, v.uv_screen_size, pool);
    p.roughness = std::clamp(eval(b.roughness, v.uv, v.uv_screen_size, pool), Real(0.01), Real(1));
    p.stint = eval(b.specular_tint, v.uv, v.uv_screen_size, pool);
    p.aniso = eval(b.anisotropic, v.uv, v.uv_screen_size, pool);
    p.sheen = eval(b.sheen, v.uv, v.uv_screen_size, pool);
    p.sheen_tint = eval(b.sheen_tint, v.uv, v.uv_screen_size, pool);
    p.cc = eval(b.clearcoat, v.uv, v.uv_screen_size, pool);
    p.cc_gloss = eval(b.clearcoat_gloss, v.uv, v.uv_screen_size, pool);
    p.eta = b.eta;
    return p;
}
inline Spectrum C0_of(const Params &p) {  // homework1.tex:478-482
    Real r0 = sqr(p.eta - 1) / sqr(p.eta + 1);
    Spectrum Ks = make_const_spectrum(1 - p.stint) + p.stint * tint_of(p.base);
    return (p.specular * r0 * (1 - p.metallic)) * Ks + p.metallic * p.base;
}
inline void weights_of(const Params &p, Real &dw, Real &mw, Real &gw, Real &cw) {  // homework1.tex:534-541
    dw = (1 - p.metallic) * (1 - p.st);
    mw = 1 - p.st * (1 - p.metallic);
    gw = (1 - p.metallic) * p.st;
    cw = Real(0.25) * p.cc;
}

}  // namespace hw
