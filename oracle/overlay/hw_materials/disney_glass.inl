// TEST INFRASTRUCTURE (CPU oracle overlay, parity unpinned): handouts/homework1.tex:343-396.
#include "hw_disney_common.h"

Spectrum eval_op::operator()(const DisneyGlass &bsdf) const {
    Real ax, ay;
    hw::disney_alphas(std::clamp(eval(bsdf.roughness, vertex.uv, vertex.uv_screen_size, texture_pool), Real(0.01), Real(1)),
                      eval(bsdf.anisotropic, vertex.uv, vertex.uv_screen_size, texture_pool), ax, ay);
    Spectrum base = eval(bsdf.base_color, vertex.uv, vertex.uv_screen_size, texture_pool);
    return hw::glass_eval(base, sqrt(base), ax, ay, bsdf.eta, vertex, dir_in, dir_out, dir);
}
Real pdf_sample_bsdf_op::operator()(const DisneyGlass &bsdf) const {
    Real ax, ay;
    hw::disney_alphas(std::clamp(eval(bsdf.roughness, vertex.uv, vertex.uv_screen_size, texture_pool), Real(0.01), Real(1)),
                      eval(bsdf.anisotropic, vertex.uv, vertex.uv_screen_size, texture_pool), ax, ay);
    return hw::glass_pdf(ax, ay, bsdf.eta, vertex, dir_in, dir_out);
}
std::optional<BSDFSampleRecord> sample_bsdf_op::operator()(const DisneyGlass &bsdf) const {
    Real roughness = std::clamp(eval(bsdf.roughness, vertex.uv, vertex.uv_screen_size, texture_pool), Real(0.01), Real(1));
    Real ax, ay;
    hw::disney_alphas(roughness, eval(bsdf.anisotropic, vertex.uv, vertex.uv_screen_size, texture_pool), ax, ay);
    return hw::glass_sample(ax, ay, roughness, bsdf.eta, vertex, dir_in, rnd_param_uv, rnd_param_w);
}
TextureSpectrum get_texture_op::operator()(const DisneyGlass &bsdf) const { return bsdf.base_color; }
