// TEST INFRASTRUCTURE (CPU oracle overlay, parity unpinned): handouts/homework1.tex:412-451.
#include "hw_disney_common.h"

Spectrum eval_op::operator()(const DisneySheen &bsdf) const {
    return hw::sheen_eval(eval(bsdf.base_color, vertex.uv, vertex.uv_screen_size, texture_pool),
                          eval(bsdf.sheen_tint, vertex.uv, vertex.uv_screen_size, texture_pool), vertex, dir_in, dir_out);
}
Real pdf_sample_bsdf_op::operator()(const DisneySheen &bsdf) const { return hw::cosine_pdf(vertex, dir_in, dir_out); }
std::optional<BSDFSampleRecord> sample_bsdf_op::operator()(const DisneySheen &bsdf) const {
    return hw::cosine_sample(vertex, dir_in, rnd_param_uv);
}
TextureSpectrum get_texture_op::operator()(const DisneySheen &bsdf) const { return bsdf.base_color; }
