// lj_texture.h -- Texture<T> evaluation and mip-mapped lookup in device memory.
// fp32 restatement of the reference's texture.h:117-163 and mipmap.h:50-88.
#pragma once
#include "lj_scene_dev.h"

namespace lj {

LJ_HD V3 texel3(const DevScene &sc, const DevImage &img, int level, int x, int y) {
    V4 t = ld4(&sc.texels3[img.offset[level] + y * img.w[level] + x]);
    return xyz(t);
}
LJ_HD float texel1(const DevScene &sc, const DevImage &img, int level, int x, int y) {
    return sc.texels1[img.offset[level] + y * img.w[level] + x];
}

// mipmap.h:50-71.  TRAP: int(u) truncates toward zero, so u in (-0.5, 0) extrapolates; kept.
template <int CH>
LJ_HD V3 mip_lookup_level(const DevScene &sc, const DevImage &img, float u, float v, int level) {
    int w = img.w[level], h = img.h[level];
    u = u * w - 0.5f;
    v = v * h - 0.5f;
    int ufi = modulo_i((int)u, w), vfi = modulo_i((int)v, h);
    int uci = modulo_i(ufi + 1, w), vci = modulo_i(vfi + 1, h);
    float uo = u - ufi, vo = v - vfi;
    V3 ff, fc, cf, cc;
    if (CH == 3) {
        ff = texel3(sc, img, level, ufi, vfi); fc = texel3(sc, img, level, ufi, vci);
        cf = texel3(sc, img, level, uci, vfi); cc = texel3(sc, img, level, uci, vci);
    } else {
        ff = mk3(texel1(sc, img, level, ufi, vfi)); fc = mk3(texel1(sc, img, level, ufi, vci));
        cf = mk3(texel1(sc, img, level, uci, vfi)); cc = mk3(texel1(sc, img, level, uci, vci));
    }
    return ff * ((1 - uo) * (1 - vo)) + fc * ((1 - uo) * vo) + cf * (uo * (1 - vo)) + cc * (uo * vo);
}

// mipmap.h:73-88
template <int CH>
LJ_HD V3 mip_lookup(const DevScene &sc, const DevImage &img, float u, float v, float level) {
    int last = img.levels - 1;
    if (level <= 0) return mip_lookup_level<CH>(sc, img, u, v, 0);
    if (level < (float)last) {
        int fl = clampi((int)floorf(level), 0, last);
        int cl = clampi(fl + 1, 0, last);
        float off = level - fl;
        return mip_lookup_level<CH>(sc, img, u, v, fl) * (1 - off) + mip_lookup_level<CH>(sc, img, u, v, cl) * off;
    }
    return mip_lookup_level<CH>(sc, img, u, v, last);
}

// texture.h:132-139 -- the level the image texture is sampled at (also the mipmapLevel aux view).
LJ_HD float texture_level(const DevTexture &t, const DevImage &img, float footprint) {
    float scaled = (float)(img.w[0] > img.h[0] ? img.w[0] : img.h[0]) * fmaxf(t.uscale, t.vscale) * footprint;
    return log2f(fmaxf(scaled, 1e-8f));
}

// texture.h:117-159
template <int CH>
LJ_HD V3 eval_texture(const DevScene &sc, const DevTexture &t, V2 uv, float footprint) {
    if (t.kind == 0) return mk3(t.v0[0], t.v0[1], t.v0[2]);
    float lu = modulo_f(uv.x * t.uscale + t.uoffset, 1.0f);
    float lv = modulo_f(uv.y * t.vscale + t.voffset, 1.0f);
    if (t.kind == 1) {
        const DevImage &img = (CH == 3 ? sc.images3 : sc.images1)[t.image_id];
        return mip_lookup<CH>(sc, img, lu, lv, texture_level(t, img, footprint));
    }
    int x = 2 * modulo_i((int)(lu * 2), 2) - 1;
    int y = 2 * modulo_i((int)(lv * 2), 2) - 1;
    if (x * y == 1) return mk3(t.v0[0], t.v0[1], t.v0[2]);
    return mk3(t.v1[0], t.v1[1], t.v1[2]);
}

LJ_HD V3 eval_tex3(const DevScene &sc, const DevTexture &t, V2 uv, float footprint) {
    return eval_texture<3>(sc, t, uv, footprint);
}
LJ_HD float eval_tex1(const DevScene &sc, const DevTexture &t, V2 uv, float footprint) {
    return eval_texture<1>(sc, t, uv, footprint).x;
}

}  // namespace lj
