// TEST INFRASTRUCTURE (CPU oracle overlay, parity unpinned): handouts/homework1.tex:100-166.
#include "hw_disney_common.h"

Spectrum eval_op::operator()(const DisneyDiffuse &bsdf) const {
    Spectrum base = eval(bsdf.base_color, vertex.uv, vertex.uv_screen_size, texture_pool);
    Real roughness = std::clamp(eval(bsdf.roughness, vertex.uv, vertex.uv_screen_size, texture_pool), Real(0.01), Real(1));
    Real subsurface = eval(bsdf.subsurface, vertex.uv, vertex.uv_screen_size, texture_pool);
    return hw::diffuse_eval(base, roughness, subsurface, vertex, dir_in, dir_out);
}
Real pdf_sample_bsdf_op::operator()(const DisneyDiffuse &bsdf) const { return hw::cosine_pdf(vertex, dir_in, dir_out); }
std::optional<BSDFSampleRecord> sample_bsdf_op::operator()(const DisneyDiffuse &bsdf) const {
    return hw::cosine_sample(vertex, dir_in, rnd_param_uv);
}
TextureSpectrum get_texture_op::operator()(const DisneyDiffuse &bsdf) const { return bsdf.base_color; }
