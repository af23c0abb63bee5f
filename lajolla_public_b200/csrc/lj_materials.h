// lj_materials.h -- eval / pdf_sample_bsdf / sample_bsdf for the nine Material alternatives.
// fp32 restatement of the reference's material.cpp:4-11,90-123, microfacet.h:23-114,
// materials/lambertian.inl, roughplastic.inl, roughdielectric.inl.  The six Disney alternatives are
// homework stubs upstream (materials/disney_*.inl return zero); they follow handouts/homework1.tex
// (equation cites inline) and keep the frame-flip / below-surface early-outs of the stub files.
// eval() returns BSDF * |n . dir_out| (material.h:119-131).  Directions point away from the surface.
#pragma once
#include "lj_shapes.h"
#include "lj_texture.h"

namespace lj {

struct BsdfSample { V3 dir_out; float eta, roughness; };  // material.h:133-138

// material.cpp:4-11
LJ_HD V3 sample_cos_hemisphere(V2 u) {
    float phi = kTwoPi * u.x;
    float tmp = sqrtf(clampf(1 - u.y, 0.f, 1.f));
    return mk3(cosf(phi) * tmp, sinf(phi) * tmp, sqrtf(clampf(u.y, 0.f, 1.f)));
}

LJ_HD float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }

// microfacet.h:23-27 (the pow(max(1-c,0),5))
LJ_HD float schlick_weight(float cos_theta) { return pow5(fmaxf(1 - cos_theta, 0.f)); }
LJ_HD V3 schlick_fresnel3(V3 F0, float cos_theta) { return F0 + (mk3(1) - F0) * schlick_weight(cos_theta); }
LJ_HD float schlick_fresnel1(float F0, float cos_theta) { return F0 + (1 - F0) * schlick_weight(cos_theta); }

// microfacet.h:34-55
LJ_HD float fresnel_dielectric2(float n_dot_i, float n_dot_t, float eta) {
    float rs = (n_dot_i - eta * n_dot_t) / (n_dot_i + eta * n_dot_t);
    float rp = (eta * n_dot_i - n_dot_t) / (eta * n_dot_i + n_dot_t);
    return (rs * rs + rp * rp) / 2;
}
LJ_HD float fresnel_dielectric(float n_dot_i, float eta) {
    float n_dot_t_sq = 1 - (1 - n_dot_i * n_dot_i) / (eta * eta);
    if (n_dot_t_sq < 0) return 1;
    return fresnel_dielectric2(fabsf(n_dot_i), sqrtf(n_dot_t_sq), eta);
}

// microfacet.h:57-81
// The reference evaluates D = a2 / (pi (1 + (a2 - 1) (n.h)^2)^2) in double.  In fp32 that denominator cancels when
// the half vector is near the normal and alpha is small (veach_mi's alpha = 0.005: a2 = 2.5e-5); a cancellation-free
// rewrite (sin^2 + a2 cos^2 from the tangential components) is NOT equivalent either, because the fp32 frame the
// reference is handed is unit / orthogonal only to 1e-7 and its formula sees |n|^2 - 1 at full weight (relative
// 4e-4 on D at n.h = 0.9998, profiles/r02_bsdf_tail.txt).  So the reference's own expression is evaluated:
// 1 + (a2 - 1) (n.h)^2 = ((s.s - (n.s)^2) + a2 (n.s)^2) / s.s with s = wi + eta wo (eta = 1 for reflection, the relative
// IOR for refraction, roughdielectric.inl:33-40); the numerator's cancellation is done in fp64 (a dozen DFMAs), the
// quotient in fp32 -- no fp64 division.
// `ndh` is the caller's fp32 n.h: where the fp32 denominator is not small (t >= 0.05: relative error <= 5e-6) it is
// used as it is, and the fp64 path only runs for the near-specular configurations that need it.
LJ_HD float GTR2_alpha(V3 n, V3 wi, V3 wo, float eta, float alpha, float ndh) {
    {
        float a2f = alpha * alpha, tf = 1 + (a2f - 1) * ndh * ndh;
        if (tf >= 0.05f) return a2f / (kPi * tf * tf);
    }
    double a2 = (double)alpha * (double)alpha;
    double sx = (double)wi.x + (double)eta * wo.x, sy = (double)wi.y + (double)eta * wo.y, sz = (double)wi.z + (double)eta * wo.z;
    double ns = sx * n.x + sy * n.y + sz * n.z, ss = sx * sx + sy * sy + sz * sz;
    float t = (float)((ss - ns * ns) + a2 * (ns * ns)) / (float)ss;
    return (float)a2 / (kPi * t * t);
}
LJ_HD float GTR2(V3 n, V3 wi, V3 wo, float roughness, float ndh) { return GTR2_alpha(n, wi, wo, 1.f, roughness * roughness, ndh); }
// 1 - F for the refraction branch.  Near the critical angle F -> 1 and fp32's 1 - F keeps no digits, and h.wi itself comes
// out of the cancelling sum s = wi + eta wo.  The two delicate quantities are quotients whose numerators are formed in
// fp64: cos^2 of the incident angle (s.wi)^2 / s.s, and of the transmitted angle (eta^2 s.s - (s.s - (s.wi)^2)) /
// (eta^2 s.s); the rest is fp32, in the stable product form 1 - ((a - b) / (a + b))^2 = 4 a b / (a + b)^2 for both
// polarisations (microfacet.h:34-55).  (volpath_test5_2: relative error 0.06 -> 1.5e-4.)
LJ_HD float dielectric_transmittance(V3 wi, V3 wo, float eta) {
    double sx = (double)wi.x + (double)eta * wo.x, sy = (double)wi.y + (double)eta * wo.y, sz = (double)wi.z + (double)eta * wo.z;
    double ss = sx * sx + sy * sy + sz * sz, sw = sx * wi.x + sy * wi.y + sz * wi.z;
    double e2 = (double)eta * (double)eta;
    float ssf = (float)ss;
    if (!(ssf > 0)) return 0.f;
    float ci2 = (float)(sw * sw) / ssf;
    float ct2 = (float)(e2 * ss - (ss - sw * sw)) / ((float)e2 * ssf);
    if (ct2 < 0) return 0.f;  // total internal reflection
    float ci = sqrtf(ci2), ct = sqrtf(ct2);
    float a = ci, b = eta * ct, c = eta * ci, d = ct;
    float ts = (a + b) != 0 ? 4 * a * b / ((a + b) * (a + b)) : 0.f, tp = (c + d) != 0 ? 4 * c * d / ((c + d) * (c + d)) : 0.f;
    return (ts + tp) / 2;
}
// |wi + eta wo|^2.  With h = +-normalize(wi + eta wo) this IS (h.wi + eta h.wo)^2, the denominator of the refraction
// Jacobian (roughdielectric.inl:66-70), without the cancellation of the two dot products near the configuration
// wi = -eta wo where the half vector degenerates.
LJ_HD float half_len2(V3 wi, V3 wo, float eta) {
    double sx = (double)wi.x + (double)eta * wo.x, sy = (double)wi.y + (double)eta * wo.y, sz = (double)wi.z + (double)eta * wo.z;
    return (float)(sx * sx + sy * sy + sz * sz);
}
// The (unnormalised) half vector wi + eta wo in the shading frame.  Near the mirror / straight-through configurations
// its tangential components are differences of nearly equal numbers: they are accumulated in fp64 so that what is left
// after the cancellation still carries fp32's digits.  ggx_aniso_D divides by the squared length.
LJ_HD V3 half_local(const Frame &f, V3 wi, V3 wo, float eta) {
    double sx = (double)wi.x + (double)eta * wo.x, sy = (double)wi.y + (double)eta * wo.y, sz = (double)wi.z + (double)eta * wo.z;
    return mk3((float)(sx * f.x.x + sy * f.x.y + sz * f.x.z), (float)(sx * f.y.x + sy * f.y.y + sz * f.y.z), (float)(sx * f.n.x + sy * f.n.y + sz * f.n.z));
}
LJ_HD float smith_masking_gtr2(V3 v, float roughness) {
    float alpha = roughness * roughness;
    float a2 = alpha * alpha;
    float Lambda = (-1 + sqrtf(1 + (v.x * v.x * a2 + v.y * v.y * a2) / (v.z * v.z))) / 2;
    return 1 / (1 + Lambda);
}

// microfacet.h:85-114 (Heitz 2018), generalised to (alpha_x, alpha_y) for the Disney lobes
// (homework1.tex:251); alpha_x == alpha_y reproduces the reference's isotropic routine.
LJ_HD V3 sample_visible_normals(V3 local_dir_in, float ax, float ay, V2 u) {
    bool flipped = local_dir_in.z < 0;
    if (flipped) local_dir_in = -local_dir_in;
    V3 hemi = normalize(mk3(ax * local_dir_in.x, ay * local_dir_in.y, local_dir_in.z));
    float r = sqrtf(u.x);
    float phi = 2 * kPi * u.y;
    float t1 = r * cosf(phi);
    float t2 = r * sinf(phi);
    float s = (1 + hemi.z) / 2;
    t2 = (1 - s) * sqrtf(1 - t1 * t1) + s * t2;
    V3 disk = mk3(t1, t2, sqrtf(fmaxf(0.f, 1 - t1 * t1 - t2 * t2)));
    Frame hf = make_frame(hemi);
    V3 hn = to_world(hf, disk);
    V3 n = normalize(mk3(ax * hn.x, ay * hn.y, fmaxf(0.f, hn.z)));
    return flipped ? -n : n;
}

// Frame conventions shared by every lobe.
LJ_HD Frame frame_reflective(const Vertex &vx, V3 dir_in) {  // lambertian.inl:10-13
    Frame f = vx.shading_frame;
    if (dot(f.n, dir_in) < 0) f = flip(f);
    return f;
}
LJ_HD Frame frame_transmissive(const Vertex &vx, V3 dir_in) {  // roughdielectric.inl:6-9
    Frame f = vx.shading_frame;
    if (dot(f.n, dir_in) * dot(vx.geometric_normal, dir_in) < 0) f = flip(f);
    return f;
}
LJ_HD bool below_surface(const Vertex &vx, V3 dir_in, V3 dir_out) {  // lambertian.inl:2-6
    return dot(vx.geometric_normal, dir_in) < 0 || dot(vx.geometric_normal, dir_out) < 0;
}

// ============================================================== Lambertian (lambertian.inl)
LJ_HD V3 lambertian_eval(V3 R, const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return mk3(0);
    Frame f = frame_reflective(vx, wi);
    return R * (fmaxf(dot(f.n, wo), 0.f) / kPi);
}
LJ_HD float lambertian_pdf(const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return 0;
    Frame f = frame_reflective(vx, wi);
    return fmaxf(dot(f.n, wo), 0.f) / kPi;
}
LJ_HD bool lambertian_sample(const Vertex &vx, V3 wi, V2 u, BsdfSample &s) {
    if (dot(vx.geometric_normal, wi) < 0) return false;
    Frame f = frame_reflective(vx, wi);
    s.dir_out = to_world(f, sample_cos_hemisphere(u));
    s.eta = 0;
    s.roughness = 1;
    return true;
}

// ============================================================== RoughPlastic (roughplastic.inl)
LJ_HD V3 roughplastic_eval(V3 Kd, V3 Ks, float roughness, float eta, const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return mk3(0);
    Frame f = frame_reflective(vx, wi);
    V3 h = normalize(wi + wo);
    float n_dot_h = dot(f.n, h), n_dot_in = dot(f.n, wi), n_dot_out = dot(f.n, wo);
    if (n_dot_out <= 0 || n_dot_h <= 0) return mk3(0);
    roughness = clampf(roughness, 0.01f, 1.f);
    float F_o = fresnel_dielectric(dot(h, wo), eta);
    float D = GTR2(f.n, wi, wo, roughness, n_dot_h);
    float G = smith_masking_gtr2(to_local(f, wi), roughness) * smith_masking_gtr2(to_local(f, wo), roughness);
    V3 spec = Ks * ((G * F_o * D) / (4 * n_dot_in * n_dot_out));
    float F_i = fresnel_dielectric(dot(h, wi), eta);
    V3 diff = Kd * ((1 - F_o) * (1 - F_i) / kPi);
    return (spec + diff) * n_dot_out;
}
LJ_HD float roughplastic_pdf(V3 Kd, V3 Ks, float roughness, const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return 0;
    Frame f = frame_reflective(vx, wi);
    V3 h = normalize(wi + wo);
    float n_dot_in = dot(f.n, wi), n_dot_out = dot(f.n, wo), n_dot_h = dot(f.n, h);
    if (n_dot_out <= 0 || n_dot_h <= 0) return 0;
    float lS = luminance(Ks), lR = luminance(Kd);
    if (lS + lR <= 0) return 0;
    roughness = clampf(roughness, 0.01f, 1.f);
    float spec_prob = lS / (lS + lR);
    float diff_prob = 1 - spec_prob;
    float G = smith_masking_gtr2(to_local(f, wi), roughness);
    float D = GTR2(f.n, wi, wo, roughness, n_dot_h);
    spec_prob *= (G * D) / (4 * n_dot_in);
    diff_prob *= n_dot_out / kPi;
    return spec_prob + diff_prob;
}
LJ_HD bool roughplastic_sample(V3 Kd, V3 Ks, float roughness, const Vertex &vx, V3 wi, V2 u, float w, BsdfSample &s) {
    if (dot(vx.geometric_normal, wi) < 0) return false;
    Frame f = frame_reflective(vx, wi);
    float lS = luminance(Ks), lR = luminance(Kd);
    if (lS + lR <= 0) return false;
    float spec_prob = lS / (lS + lR);
    if (w < spec_prob) {
        V3 li = to_local(f, wi);
        roughness = clampf(roughness, 0.01f, 1.f);
        float alpha = roughness * roughness;
        V3 h = to_world(f, sample_visible_normals(li, alpha, alpha, u));
        s.dir_out = normalize(-wi + 2 * dot(wi, h) * h);
        s.eta = 0;
        s.roughness = roughness;
    } else {
        s.dir_out = to_world(f, sample_cos_hemisphere(u));
        s.eta = 0;
        s.roughness = 1;
    }
    return true;
}

// ===================================================== RoughDielectric (roughdielectric.inl)
// Shared with DisneyGlass: Cr / Ct are the reflection / transmission tints, (ax, ay) the GGX
// alphas, D and G evaluated by the anisotropic forms below (isotropic when ax == ay).
LJ_HD float ggx_aniso_D(V3 hl, float ax, float ay) {  // homework1.tex:197-201; hl need not be normalised
    float t = (hl.x * hl.x / (ax * ax) + hl.y * hl.y / (ay * ay) + hl.z * hl.z) / (hl.x * hl.x + hl.y * hl.y + hl.z * hl.z);
    return 1 / (kPi * ax * ay * t * t);
}
LJ_HD float smith_aniso_G1(V3 wl, float ax, float ay) {  // homework1.tex:213-220
    float Lambda = (sqrtf(1 + ((wl.x * ax) * (wl.x * ax) + (wl.y * ay) * (wl.y * ay)) / (wl.z * wl.z)) - 1) / 2;
    return 1 / (1 + Lambda);
}

LJ_HD V3 dielectric_eval(V3 Cr, V3 Ct, float ax, float ay, float mat_eta, const Vertex &vx, V3 wi, V3 wo, int transport) {
    bool reflect = dot(vx.geometric_normal, wi) * dot(vx.geometric_normal, wo) > 0;
    Frame f = frame_transmissive(vx, wi);
    float eta = dot(vx.geometric_normal, wi) > 0 ? mat_eta : 1 / mat_eta;
    V3 h = reflect ? normalize(wi + wo) : normalize(wi + wo * eta);
    if (dot(h, f.n) < 0) h = -h;
    float h_dot_in = dot(h, wi);
    float F = fresnel_dielectric(h_dot_in, eta);
    // isotropic (RoughDielectric, roughdielectric.inl:60: GTR2 of n.h): the reference's expression in fp64, see GTR2_alpha
    float D = ax == ay ? GTR2_alpha(f.n, wi, wo, reflect ? 1.f : eta, ax, dot(f.n, h)) : ggx_aniso_D(half_local(f, wi, wo, reflect ? 1.f : eta), ax, ay);
    float G = smith_aniso_G1(to_local(f, wi), ax, ay) * smith_aniso_G1(to_local(f, wo), ax, ay);
    if (reflect) return Cr * ((F * D * G) / (4 * fabsf(dot(f.n, wi))));
    float eta_factor = transport == 0 ? (1 / (eta * eta)) : 1;  // roughdielectric.inl:64
    float h_dot_out = dot(h, wo);
    return Ct * ((eta_factor * dielectric_transmittance(wi, wo, eta) * D * G * eta * eta * fabsf(h_dot_out * h_dot_in)) /
                 (fabsf(dot(f.n, wi)) * half_len2(wi, wo, eta)));
}
LJ_HD float dielectric_pdf(float ax, float ay, float mat_eta, const Vertex &vx, V3 wi, V3 wo) {
    bool reflect = dot(vx.geometric_normal, wi) * dot(vx.geometric_normal, wo) > 0;
    Frame f = frame_transmissive(vx, wi);
    float eta = dot(vx.geometric_normal, wi) > 0 ? mat_eta : 1 / mat_eta;
    V3 h = reflect ? normalize(wi + wo) : normalize(wi + wo * eta);
    if (dot(h, f.n) < 0) h = -h;
    float h_dot_in = dot(h, wi);
    float F = fresnel_dielectric(h_dot_in, eta);
    float D = ax == ay ? GTR2_alpha(f.n, wi, wo, reflect ? 1.f : eta, ax, dot(f.n, h)) : ggx_aniso_D(half_local(f, wi, wo, reflect ? 1.f : eta), ax, ay);
    float G_in = smith_aniso_G1(to_local(f, wi), ax, ay);
    if (reflect) return (F * D * G_in) / (4 * fabsf(dot(f.n, wi)));
    float h_dot_out = dot(h, wo);
    float dh_dout = eta * eta * h_dot_out / half_len2(wi, wo, eta);
    return dielectric_transmittance(wi, wo, eta) * D * G_in * fabsf(dh_dout * h_dot_in / dot(f.n, wi));
}
LJ_HD bool dielectric_sample(float ax, float ay, float roughness, float mat_eta, const Vertex &vx, V3 wi, V2 u, float w, BsdfSample &s) {
    float eta = dot(vx.geometric_normal, wi) > 0 ? mat_eta : 1 / mat_eta;
    Frame f = frame_transmissive(vx, wi);
    V3 h = to_world(f, sample_visible_normals(to_local(f, wi), ax, ay, u));
    if (dot(h, f.n) < 0) h = -h;
    float h_dot_in = dot(h, wi);
    float F = fresnel_dielectric(h_dot_in, eta);
    if (w <= F) {
        s.dir_out = normalize(-wi + 2 * dot(wi, h) * h);
        s.eta = 0;
        s.roughness = roughness;
        return true;
    }
    float h_dot_out_sq = 1 - (1 - h_dot_in * h_dot_in) / (eta * eta);
    if (h_dot_out_sq <= 0) return false;
    if (h_dot_in < 0) h = -h;
    float h_dot_out = sqrtf(h_dot_out_sq);
    s.dir_out = -wi / eta + (fabsf(h_dot_in) / eta - h_dot_out) * h;
    s.eta = eta;
    s.roughness = roughness;
    return true;
}

// ========================================================================== Disney lobes
LJ_HD void disney_alphas(float roughness, float anisotropic, float &ax, float &ay) {  // homework1.tex:204-210
    float aspect = sqrtf(1 - 0.9f * anisotropic);
    ax = fmaxf(1e-4f, roughness * roughness / aspect);
    ay = fmaxf(1e-4f, roughness * roughness * aspect);
}
LJ_HD V3 disney_tint(V3 base) {  // homework1.tex:417
    float l = luminance(base);
    return l > 0 ? base / l : mk3(1);
}

// diffuse + subsurface, homework1.tex:100-135
LJ_HD V3 disney_diffuse_eval(V3 base, float roughness, float subsurface, const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return mk3(0);
    Frame f = frame_reflective(vx, wi);
    V3 h = normalize(wi + wo);
    float n_in = fabsf(dot(f.n, wi)), n_out = fabsf(dot(f.n, wo));
    float h_out = fabsf(dot(h, wo));
    float FD90 = 0.5f + 2 * roughness * h_out * h_out;
    float FD_in = 1 + (FD90 - 1) * pow5(1 - n_in), FD_out = 1 + (FD90 - 1) * pow5(1 - n_out);
    V3 base_diffuse = base * (FD_in * FD_out * n_out / kPi);
    float FSS90 = roughness * h_out * h_out;
    float FSS_in = 1 + (FSS90 - 1) * pow5(1 - n_in), FSS_out = 1 + (FSS90 - 1) * pow5(1 - n_out);
    float denom = n_in + n_out;
    float ss_term = denom > 0 ? (FSS_in * FSS_out * (1 / denom - 0.5f) + 0.5f) : 0.5f;
    V3 ss = base * (1.25f * ss_term * n_out / kPi);
    return (1 - subsurface) * base_diffuse + subsurface * ss;
}
// cosine-hemisphere sampling (homework1.tex:166): pdf and sample are Lambertian's.

// metal, homework1.tex:182-222; F is supplied so the full BSDF can pass its modified Fresnel.
LJ_HD bool disney_metal_geom(float ax, float ay, const Vertex &vx, V3 wi, V3 wo, V3 &h, float &D, float &G_in, float &G_out, float &n_in) {
    if (below_surface(vx, wi, wo)) return false;
    Frame f = frame_reflective(vx, wi);
    h = normalize(wi + wo);
    n_in = dot(f.n, wi);
    if (dot(f.n, wo) <= 0 || dot(f.n, h) <= 0 || n_in <= 0) return false;
    D = ggx_aniso_D(to_local(f, h), ax, ay);
    G_in = smith_aniso_G1(to_local(f, wi), ax, ay);
    G_out = smith_aniso_G1(to_local(f, wo), ax, ay);
    return true;
}
LJ_HD V3 disney_metal_eval(V3 F0, float ax, float ay, const Vertex &vx, V3 wi, V3 wo) {
    V3 h; float D, Gi, Go, n_in;
    if (!disney_metal_geom(ax, ay, vx, wi, wo, h, D, Gi, Go, n_in)) return mk3(0);
    V3 F = schlick_fresnel3(F0, fabsf(dot(h, wo)));
    return F * (D * Gi * Go / (4 * fabsf(n_in)));
}
LJ_HD float disney_metal_pdf(float ax, float ay, const Vertex &vx, V3 wi, V3 wo) {
    V3 h; float D, Gi, Go, n_in;
    if (!disney_metal_geom(ax, ay, vx, wi, wo, h, D, Gi, Go, n_in)) return 0;
    return D * Gi / (4 * fabsf(n_in));
}
LJ_HD bool disney_metal_sample(float ax, float ay, float roughness, const Vertex &vx, V3 wi, V2 u, BsdfSample &s) {
    if (dot(vx.geometric_normal, wi) < 0) return false;
    Frame f = frame_reflective(vx, wi);
    V3 h = to_world(f, sample_visible_normals(to_local(f, wi), ax, ay, u));
    s.dir_out = normalize(-wi + 2 * dot(wi, h) * h);
    s.eta = 0;
    s.roughness = roughness;
    return true;
}

// clearcoat, homework1.tex:267-327
LJ_HD float clearcoat_alpha(float gloss) { return (1 - gloss) * 0.1f + gloss * 0.001f; }
LJ_HD float clearcoat_D(float a2, float hlz) { return (a2 - 1) / (kPi * logf(a2) * (1 + (a2 - 1) * hlz * hlz)); }
LJ_HD float disney_clearcoat_eval(float gloss, const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return 0;
    Frame f = frame_reflective(vx, wi);
    V3 h = normalize(wi + wo);
    float n_in = dot(f.n, wi);
    if (dot(f.n, wo) <= 0 || dot(f.n, h) <= 0 || n_in <= 0) return 0;
    float ag = clearcoat_alpha(gloss);
    float Fc = schlick_fresnel1(0.04f, fabsf(dot(h, wo)));  // R0(eta = 1.5)
    float Dc = clearcoat_D(ag * ag, dot(f.n, h));
    float Gc = smith_aniso_G1(to_local(f, wi), 0.25f, 0.25f) * smith_aniso_G1(to_local(f, wo), 0.25f, 0.25f);
    return Fc * Dc * Gc / (4 * fabsf(n_in));
}
LJ_HD float disney_clearcoat_pdf(float gloss, const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return 0;
    Frame f = frame_reflective(vx, wi);
    V3 h = normalize(wi + wo);
    float n_h = dot(f.n, h);
    if (dot(f.n, wo) <= 0 || n_h <= 0) return 0;
    float ag = clearcoat_alpha(gloss);
    return clearcoat_D(ag * ag, n_h) * fabsf(n_h) / (4 * fabsf(dot(h, wo)));
}
LJ_HD bool disney_clearcoat_sample(float gloss, const Vertex &vx, V3 wi, V2 u, BsdfSample &s) {
    if (dot(vx.geometric_normal, wi) < 0) return false;
    Frame f = frame_reflective(vx, wi);
    float ag = clearcoat_alpha(gloss);
    float a2 = ag * ag;
    float cos_e = sqrtf(clampf((1 - powf(a2, 1 - u.x)) / (1 - a2), 0.f, 1.f));
    float sin_e = sqrtf(fmaxf(0.f, 1 - cos_e * cos_e));
    float az = 2 * kPi * u.y;
    V3 h = to_world(f, mk3(sin_e * cosf(az), sin_e * sinf(az), cos_e));
    s.dir_out = normalize(-wi + 2 * dot(wi, h) * h);
    s.eta = 0;
    s.roughness = sqrtf(ag);
    return true;
}

// sheen, homework1.tex:412-425 (cosine-hemisphere sampled)
LJ_HD V3 disney_sheen_eval(V3 base, float sheen_tint, const Vertex &vx, V3 wi, V3 wo) {
    if (below_surface(vx, wi, wo)) return mk3(0);
    Frame f = frame_reflective(vx, wi);
    V3 h = normalize(wi + wo);
    V3 Csheen = mk3(1 - sheen_tint) + sheen_tint * disney_tint(base);
    return Csheen * (pow5(1 - fabsf(dot(h, wo))) * fabsf(dot(f.n, wo)));
}

// ========================================================================== dispatch
struct MatParams {  // textures of one material evaluated at one vertex
    V3 c0, c1;      // slot 0 / slot 1 colours
    float p[kNumTexSlots];
};

LJ_HD V3 mat_tex3(const DevScene &sc, const DevMaterial &m, int slot, const Vertex &vx) {
    return eval_tex3(sc, m.tex[slot], vx.uv, vx.uv_screen_size);
}
LJ_HD float mat_tex1(const DevScene &sc, const DevMaterial &m, int slot, const Vertex &vx) {
    return eval_tex1(sc, m.tex[slot], vx.uv, vx.uv_screen_size);
}

struct DisneyParams {
    V3 base;
    float st, metallic, subsurface, specular, roughness, stint, aniso, sheen, sheen_tint, cc, cc_gloss, eta;
};
LJ_HD DisneyParams disney_params(const DevScene &sc, const DevMaterial &m, const Vertex &vx) {
    DisneyParams p;
    p.base = mat_tex3(sc, m, 0, vx);
    p.subsurface = mat_tex1(sc, m, 1, vx);
    p.roughness = clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f);
    p.aniso = mat_tex1(sc, m, 3, vx);
    p.cc_gloss = mat_tex1(sc, m, 4, vx);
    p.sheen_tint = mat_tex1(sc, m, 5, vx);
    p.st = mat_tex1(sc, m, 6, vx);
    p.metallic = mat_tex1(sc, m, 7, vx);
    p.specular = mat_tex1(sc, m, 8, vx);
    p.stint = mat_tex1(sc, m, 9, vx);
    p.sheen = mat_tex1(sc, m, 10, vx);
    p.cc = mat_tex1(sc, m, 11, vx);
    p.eta = m.eta;
    return p;
}
// modified metal Fresnel base C0, homework1.tex:478-482
LJ_HD V3 disney_C0(const DisneyParams &p) {
    float r0 = (p.eta - 1) * (p.eta - 1) / ((p.eta + 1) * (p.eta + 1));
    V3 Ks = mk3(1 - p.stint) + p.stint * disney_tint(p.base);
    return (p.specular * r0 * (1 - p.metallic)) * Ks + p.metallic * p.base;
}
LJ_HD void disney_weights(const DisneyParams &p, float &dw, float &mw, float &gw, float &cw) {  // homework1.tex:534-541
    dw = (1 - p.metallic) * (1 - p.st);
    mw = 1 - p.st * (1 - p.metallic);
    gw = (1 - p.metallic) * p.st;
    cw = 0.25f * p.cc;
}

LJ_HD V3 disney_bsdf_eval(const DisneyParams &p, const Vertex &vx, V3 wi, V3 wo, int transport) {
    float ax, ay;
    disney_alphas(p.roughness, p.aniso, ax, ay);
    V3 glass = dielectric_eval(p.base, sqrt3(p.base), ax, ay, p.eta, vx, wi, wo, transport);
    float gscale = (1 - p.metallic) * p.st;
    if (dot(vx.geometric_normal, wi) <= 0) return gscale * glass;  // inside: glass only (:486-494)
    V3 f = gscale * glass;
    f += ((1 - p.st) * (1 - p.metallic)) * disney_diffuse_eval(p.base, p.roughness, p.subsurface, vx, wi, wo);
    f += ((1 - p.metallic) * p.sheen) * disney_sheen_eval(p.base, p.sheen_tint, vx, wi, wo);
    f += (1 - p.st * (1 - p.metallic)) * disney_metal_eval(disney_C0(p), ax, ay, vx, wi, wo);
    f += mk3(0.25f * p.cc * disney_clearcoat_eval(p.cc_gloss, vx, wi, wo));
    return f;
}
LJ_HD float disney_bsdf_pdf(const DisneyParams &p, const Vertex &vx, V3 wi, V3 wo) {
    float ax, ay;
    disney_alphas(p.roughness, p.aniso, ax, ay);
    float pg = dielectric_pdf(ax, ay, p.eta, vx, wi, wo);
    if (dot(vx.geometric_normal, wi) <= 0) return pg;
    float dw, mw, gw, cw;
    disney_weights(p, dw, mw, gw, cw);
    float total = dw + mw + gw + cw;
    if (total <= 0) return 0;
    float pdf = dw * lambertian_pdf(vx, wi, wo) + mw * disney_metal_pdf(ax, ay, vx, wi, wo) + gw * pg +
                cw * disney_clearcoat_pdf(p.cc_gloss, vx, wi, wo);
    return pdf / total;
}
LJ_HD bool disney_bsdf_sample(const DisneyParams &p, const Vertex &vx, V3 wi, V2 u, float w, BsdfSample &s) {
    float ax, ay;
    disney_alphas(p.roughness, p.aniso, ax, ay);
    if (dot(vx.geometric_normal, wi) <= 0) return dielectric_sample(ax, ay, p.roughness, p.eta, vx, wi, u, w, s);
    float dw, mw, gw, cw;
    disney_weights(p, dw, mw, gw, cw);
    float total = dw + mw + gw + cw;
    if (total <= 0) return false;
    dw /= total; mw /= total; gw /= total;
    if (w < dw) return lambertian_sample(vx, wi, u, s);
    if (w < dw + mw) return disney_metal_sample(ax, ay, p.roughness, vx, wi, u, s);
    if (w < dw + mw + gw) {
        float w2 = (w - (dw + mw)) / gw;  // rescale for the reflect/refract pick (homework1.tex:548)
        return dielectric_sample(ax, ay, p.roughness, p.eta, vx, wi, u, w2, s);
    }
    return disney_clearcoat_sample(p.cc_gloss, vx, wi, u, s);
}

// material.cpp:90-123: the three entry points, switch instead of std::visit.
// CLASS selects which alternatives are compiled in: kMatAll, kMatSmall (Lambertian, RoughPlastic, RoughDielectric) or
// kMatDisney (the six Disney alternatives).  Scenes with Disney materials shade in two passes, one kernel per class
// (wavefront.cu): the Disney lobes are ~10x the code of the small materials, and a kernel that holds both spends its
// time waiting for instructions (60 % of the stall samples "no instruction", profiles/r02g_disney_shade_ncu.txt).
enum { kMatAll = 0, kMatSmall = 1, kMatDisney = 2, kMatLambert = 3 /* scenes whose every material is Lambertian: shade_path only */ };
LJ_HD bool material_is_disney(int type) { return type >= LJ_MAT_DISNEY_DIFFUSE; }

// A material at one vertex.  For the principled BSDF the twelve parameter textures are evaluated ONCE here instead of
// inside each of the five eval / pdf / sample calls of a shade step.
struct MatCtx {
    const DevMaterial *m;
    DisneyParams dp;
};
template <int CLASS>
LJ_HD MatCtx mat_ctx(const DevScene &sc, const DevMaterial &m, const Vertex &vx) {
    MatCtx c;
    c.m = &m;
    if (CLASS != kMatSmall && m.type == LJ_MAT_DISNEY_BSDF) c.dp = disney_params(sc, m, vx);
    return c;
}

template <int CLASS>
LJ_HD V3 bsdf_eval(const DevScene &sc, const MatCtx &c, V3 wi, V3 wo, const Vertex &vx, int transport) {
    const DevMaterial &m = *c.m;
    if (CLASS != kMatDisney) {
        switch (m.type) {
            case 0: return lambertian_eval(mat_tex3(sc, m, 0, vx), vx, wi, wo);
            case 1: return roughplastic_eval(mat_tex3(sc, m, 0, vx), mat_tex3(sc, m, 1, vx), mat_tex1(sc, m, 2, vx), m.eta, vx, wi, wo);
            case 2: {
                float r = clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f);
                return dielectric_eval(mat_tex3(sc, m, 1, vx), mat_tex3(sc, m, 0, vx), r * r, r * r, m.eta, vx, wi, wo, transport);
            }
        }
    }
    if (CLASS != kMatSmall) {
        switch (m.type) {
            case 3: return disney_diffuse_eval(mat_tex3(sc, m, 0, vx), clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f), mat_tex1(sc, m, 1, vx), vx, wi, wo);
            case 4: {
                float ax, ay;
                disney_alphas(clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f), mat_tex1(sc, m, 3, vx), ax, ay);
                return disney_metal_eval(mat_tex3(sc, m, 0, vx), ax, ay, vx, wi, wo);
            }
            case 5: {
                float ax, ay;
                disney_alphas(clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f), mat_tex1(sc, m, 3, vx), ax, ay);
                V3 base = mat_tex3(sc, m, 0, vx);
                return dielectric_eval(base, sqrt3(base), ax, ay, m.eta, vx, wi, wo, transport);
            }
            case 6: return mk3(disney_clearcoat_eval(mat_tex1(sc, m, 4, vx), vx, wi, wo));
            case 7: return disney_sheen_eval(mat_tex3(sc, m, 0, vx), mat_tex1(sc, m, 5, vx), vx, wi, wo);
            case 8: return disney_bsdf_eval(c.dp, vx, wi, wo, transport);
        }
    }
    return mk3(0);
}

template <int CLASS>
LJ_HD float bsdf_pdf(const DevScene &sc, const MatCtx &c, V3 wi, V3 wo, const Vertex &vx) {
    const DevMaterial &m = *c.m;
    if (CLASS != kMatDisney) {
        switch (m.type) {
            case 0: return lambertian_pdf(vx, wi, wo);
            case 1: return roughplastic_pdf(mat_tex3(sc, m, 0, vx), mat_tex3(sc, m, 1, vx), mat_tex1(sc, m, 2, vx), vx, wi, wo);
            case 2: {
                float r = clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f);
                return dielectric_pdf(r * r, r * r, m.eta, vx, wi, wo);
            }
        }
    }
    if (CLASS != kMatSmall) {
        switch (m.type) {
            case 3: case 7: return lambertian_pdf(vx, wi, wo);
            case 4: case 5: {
                float ax, ay;
                disney_alphas(clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f), mat_tex1(sc, m, 3, vx), ax, ay);
                return m.type == 4 ? disney_metal_pdf(ax, ay, vx, wi, wo) : dielectric_pdf(ax, ay, m.eta, vx, wi, wo);
            }
            case 6: return disney_clearcoat_pdf(mat_tex1(sc, m, 4, vx), vx, wi, wo);
            case 8: return disney_bsdf_pdf(c.dp, vx, wi, wo);
        }
    }
    return 0;
}

template <int CLASS>
LJ_HD bool bsdf_sample(const DevScene &sc, const MatCtx &c, V3 wi, const Vertex &vx, V2 u, float w, BsdfSample &s) {
    const DevMaterial &m = *c.m;
    if (CLASS != kMatDisney) {
        switch (m.type) {
            case 0: return lambertian_sample(vx, wi, u, s);
            case 1: return roughplastic_sample(mat_tex3(sc, m, 0, vx), mat_tex3(sc, m, 1, vx), mat_tex1(sc, m, 2, vx), vx, wi, u, w, s);
            case 2: {
                float r = clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f);
                return dielectric_sample(r * r, r * r, r, m.eta, vx, wi, u, w, s);
            }
        }
    }
    if (CLASS != kMatSmall) {
        switch (m.type) {
            case 3: case 7: return lambertian_sample(vx, wi, u, s);
            case 4: case 5: {
                float r = clampf(mat_tex1(sc, m, 2, vx), 0.01f, 1.f);
                float ax, ay;
                disney_alphas(r, mat_tex1(sc, m, 3, vx), ax, ay);
                if (m.type == 4) return disney_metal_sample(ax, ay, r, vx, wi, u, s);
                return dielectric_sample(ax, ay, r, m.eta, vx, wi, u, w, s);
            }
            case 6: return disney_clearcoat_sample(mat_tex1(sc, m, 4, vx), vx, wi, u, s);
            case 8: return disney_bsdf_sample(c.dp, vx, wi, u, w, s);
        }
    }
    return false;
}

// The query seam and the volpath integrator take a material and evaluate everything (one context per call).
LJ_HD V3 bsdf_eval(const DevScene &sc, const DevMaterial &m, V3 wi, V3 wo, const Vertex &vx, int transport) {
    return bsdf_eval<kMatAll>(sc, mat_ctx<kMatAll>(sc, m, vx), wi, wo, vx, transport);
}
LJ_HD float bsdf_pdf(const DevScene &sc, const DevMaterial &m, V3 wi, V3 wo, const Vertex &vx) {
    return bsdf_pdf<kMatAll>(sc, mat_ctx<kMatAll>(sc, m, vx), wi, wo, vx);
}
LJ_HD bool bsdf_sample(const DevScene &sc, const DevMaterial &m, V3 wi, const Vertex &vx, V2 u, float w, BsdfSample &s) {
    return bsdf_sample<kMatAll>(sc, mat_ctx<kMatAll>(sc, m, vx), wi, vx, u, w, s);
}

// Out-of-line copies of the Disney-class dispatchers (one per translation unit instead of one per call site): inlined
// five times, the Disney lobes made the shade kernel 1.5 MB of SASS.
LJ_HD_CALL V3 bsdf_eval_call(const DevScene &sc, const MatCtx &c, V3 wi, V3 wo, const Vertex &vx, int transport) {
    return bsdf_eval<kMatDisney>(sc, c, wi, wo, vx, transport);
}
LJ_HD_CALL float bsdf_pdf_call(const DevScene &sc, const MatCtx &c, V3 wi, V3 wo, const Vertex &vx) {
    return bsdf_pdf<kMatDisney>(sc, c, wi, wo, vx);
}
LJ_HD_CALL bool bsdf_sample_call(const DevScene &sc, const MatCtx &c, V3 wi, const Vertex &vx, V2 u, float w, BsdfSample &s) {
    return bsdf_sample<kMatDisney>(sc, c, wi, vx, u, w, s);
}

// The same for every material (the volpath integrator's shade kernel: surfaces are the rare event there, and the five
// inlined all-material dispatchers made k_shade_vol 1.6 MB of SASS -- 7 of every 8 stall cycles waiting for instructions)
LJ_HD_CALL MatCtx mat_ctx_all_call(const DevScene &sc, const DevMaterial &m, const Vertex &vx) { return mat_ctx<kMatAll>(sc, m, vx); }
LJ_HD_CALL V3 bsdf_eval_all_call(const DevScene &sc, const MatCtx &c, V3 wi, V3 wo, const Vertex &vx, int transport) {
    return bsdf_eval<kMatAll>(sc, c, wi, wo, vx, transport);
}
LJ_HD_CALL float bsdf_pdf_all_call(const DevScene &sc, const MatCtx &c, V3 wi, V3 wo, const Vertex &vx) {
    return bsdf_pdf<kMatAll>(sc, c, wi, wo, vx);
}
LJ_HD_CALL bool bsdf_sample_all_call(const DevScene &sc, const MatCtx &c, V3 wi, const Vertex &vx, V2 u, float w, BsdfSample &s) {
    return bsdf_sample<kMatAll>(sc, c, wi, vx, u, w, s);
}

}  // namespace lj
