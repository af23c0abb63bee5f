// TEST INFRASTRUCTURE (CPU oracle overlay, parity unpinned): handouts/homework1.tex:267-327.
#include "hw_disney_common.h"

Spectrum eval_op::operator()(const DisneyClearcoat &bsdf) const {
    return make_const_spectrum(hw::clearcoat_eval(eval(bsdf.clearcoat_gloss, vertex.uv, vertex.uv_screen_size, texture_pool), vertex, dir_in, dir_out));
}
Real pdf_sample_bsdf_op::operator()(const DisneyClearcoat &bsdf) const {
    return hw::clearcoat_pdf(eval(bsdf.clearcoat_gloss, vertex.uv, vertex.uv_screen_size, texture_pool), vertex, dir_in, dir_out);
}
std::optional<BSDFSampleRecord> sample_bsdf_op::operator()(const DisneyClearcoat &bsdf) const {
    return hw::clearcoat_sample(eval(bsdf.clearcoat_gloss, vertex.uv, vertex.uv_screen_size, texture_pool), vertex, dir_in, rnd_param_uv);
}
TextureSpectrum get_texture_op::operator()(const DisneyClearcoat &bsdf) const { return make_constant_spectrum_texture(make_zero_spectrum()); }
