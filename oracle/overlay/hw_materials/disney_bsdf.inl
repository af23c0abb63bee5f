// TEST INFRASTRUCTURE (CPU oracle overlay, parity unpinned): handouts/homework1.tex:461-548.
#include "hw_disney_common.h"

Spectrum eval_op::operator()(const DisneyBSDF &bsdf) const {
    hw::Params p = hw::params_of(bsdf, vertex, texture_pool);
    Real ax, ay;
    hw::disney_alphas(p.roughness, p.aniso, ax, ay);
    Spectrum glass = hw::glass_eval(p.base, sqrt(p.base), ax, ay, p.eta, vertex, dir_in, dir_out, dir);
    Real gscale = (1 - p.metallic) * p.st;
    if (dot(vertex.geometric_normal, dir_in) <= 0) return gscale * glass;  // inside the object: glass only (:486-494)
    Spectrum f = gscale * glass;
    f += ((1 - p.st) * (1 - p.metallic)) * hw::diffuse_eval(p.base, p.roughness, p.subsurface, vertex, dir_in, dir_out);
    f += ((1 - p.metallic) * p.sheen) * hw::sheen_eval(p.base, p.sheen_tint, vertex, dir_in, dir_out);
    f += (1 - p.st * (1 - p.metallic)) * hw::metal_eval(hw::C0_of(p), ax, ay, vertex, dir_in, dir_out);
    f += make_const_spectrum(Real(0.25) * p.cc * hw::clearcoat_eval(p.cc_gloss, vertex, dir_in, dir_out));
    return f;
}
Real pdf_sample_bsdf_op::operator()(const DisneyBSDF &bsdf) const {
    hw::Params p = hw::params_of(bsdf, vertex, texture_pool);
    Real ax, ay;
    hw::disney_alphas(p.roughness, p.aniso, ax, ay);
    Real pg = hw::glass_pdf(ax, ay, p.eta, vertex, dir_in, dir_out);
    if (dot(vertex.geometric_normal, dir_in) <= 0) return pg;
    Real dw, mw, gw, cw;
    hw::weights_of(p, dw, mw, gw, cw);
    Real total = dw + mw + gw + cw;
    if (total <= 0) return 0;
    return (dw * hw::cosine_pdf(vertex, dir_in, dir_out) + mw * hw::metal_pdf(ax, ay, vertex, dir_in, dir_out) + gw * pg +
            cw * hw::clearcoat_pdf(p.cc_gloss, vertex, dir_in, dir_out)) / total;
}
std::optional<BSDFSampleRecord> sample_bsdf_op::operator()(const DisneyBSDF &bsdf) const {
    hw::Params p = hw::params_of(bsdf, vertex, texture_pool);
    Real ax, ay;
    hw::disney_alphas(p.roughness, p.aniso, ax, ay);
    if (dot(vertex.geometric_normal, dir_in) <= 0)
        return hw::glass_sample(ax, ay, p.roughness, p.eta, vertex, dir_in, rnd_param_uv, rnd_param_w);
    Real dw, mw, gw, cw;
    hw::weights_of(p, dw, mw, gw, cw);
    Real total = dw + mw + gw + cw;
    if (total <= 0) return {};
    dw /= total; mw /= total; gw /= total;
    Real w = rnd_param_w;
    if (w < dw) return hw::cosine_sample(vertex, dir_in, rnd_param_uv);
    if (w < dw + mw) return hw::metal_sample(ax, ay, p.roughness, vertex, dir_in, rnd_param_uv);
    if (w < dw + mw + gw)  // rescale w for the reflect / refract decision (homework1.tex:548)
        return hw::glass_sample(ax, ay, p.roughness, p.eta, vertex, dir_in, rnd_param_uv, (w - (dw + mw)) / gw);
    return hw::clearcoat_sample(p.cc_gloss, vertex, dir_in, rnd_param_uv);
}
TextureSpectrum get_texture_op::operator()(const DisneyBSDF &bsdf) const { return bsdf.base_color; }
