// TEST INFRASTRUCTURE (CPU oracle overlay, parity unpinned): handouts/homework1.tex:182-251.
#include "hw_disney_common.h"

Spectrum eval_op::operator()(const DisneyMetal &bsdf) const {
    Real ax, ay;
    hw::disney_alphas(std::clamp(eval(bsdf.roughness, vertex.uv, vertex.uv_screen_size, texture_pool), Real(0.01), Real(1)),
                      eval(bsdf.anisotropic, vertex.uv, vertex.uv_screen_size, texture_pool), ax, ay);
    return hw::metal_eval(eval(bsdf.base_color, vertex.uv, vertex.uv_screen_size, texture_pool), ax, ay, vertex, dir_in, dir_out);
}
Real pdf_sample_bsdf_op::operator()(const DisneyMetal &bsdf) const {
    Real ax, ay;
    hw::disney_alphas(std::clamp(eval(bsdf.roughness, vertex.uv, vertex.uv_screen_size, texture_pool), Real(0.01), Real(1)),
                      eval(bsdf.anisotropic, vertex.uv, vertex.uv_screen_size, texture_pool), ax, ay);
    return hw::metal_pdf(ax, ay, vertex, dir_in, dir_out);
}
std::optional<BSDFSampleRecord> sample_bsdf_op::operator()(const DisneyMetal &bsdf) const {
    Real roughness = std::clamp(eval(bsdf.roughness, vertex.uv, vertex.uv_screen_size, texture_pool), Real(0.01), Real(1));
    Real ax, ay;
    hw::disney_alphas(roughness, eval(bsdf.anisotropic, vertex.uv, vertex.uv_screen_size, texture_pool), ax, ay);
    return hw::metal_sample(ax, ay, roughness, vertex, dir_in, rnd_param_uv);
}
TextureSpectrum get_texture_op::operator()(const DisneyMetal &bsdf) const { return bsdf.base_color; }
