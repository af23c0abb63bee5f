// lj_image_io.cpp -- see lj_image_io.h.  The decoders restate the published algorithms of the libraries the reference
// links (stb_image 2.x for JPEG/PNG/HDR, tinyexr / OpenEXR for EXR) closely enough that the decoded texels are the
// same numbers the reference's textures hold: baseline JPEG with stb's integer IDCT, its h2v2 "fancy" chroma
// upsampling and its fixed-point YCbCr conversion (incl. the different rounding of the 8-pixel SSE2 groups and the
// scalar row tail); PNG via zlib; OpenEXR scanline files with NO / RLE / ZIPS / ZIP / PIZ compression.
#include "lj_image_io.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace ljhost {

namespace {

[[noreturn]] void fail(const std::string &what) { throw std::runtime_error(what); }

std::vector<unsigned char> read_file(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) fail("cannot open " + path);
    std::vector<unsigned char> buf;
    unsigned char tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return buf;
}

std::string lower_ext(const std::string &path) {
    size_t dot = path.find_last_of('.');
    size_t slash = path.find_last_of("/\\");
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
    std::string e = path.substr(dot);
    for (auto &c : e) c = (char)tolower((unsigned char)c);
    return e;
}

// ------------------------------------------------------------------------------------------------ JPEG (baseline)
const unsigned char kDezigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                                          6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                                          39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huffman {
    int maxcode[18];   // largest code of each length, left-justified to 16 bits, +1
    int delta[17];     // index of the first symbol of a length minus its first code
    unsigned char values[256];
    unsigned char size[257];
    unsigned short code[256];
    bool build(const int *count) {
        int k = 0;
        for (int i = 0; i < 16; i++)
            for (int j = 0; j < count[i]; j++) size[k++] = (unsigned char)(i + 1);
        size[k] = 0;
        int c = 0;
        k = 0;
        for (int j = 1; j <= 16; j++) {
            delta[j] = k - c;
            if (size[k] == j) {
                while (size[k] == j) code[k++] = (unsigned short)(c++);
                if (c - 1 >= (1 << j)) return false;
            }
            maxcode[j] = c << (16 - j);
            c <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        return true;
    }
};

struct JpegComp {
    int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0;
    int dc_pred = 0;
    int x = 0, y = 0, w2 = 0, h2 = 0;
    std::vector<unsigned char> data;
};

struct JpegDecoder {
    const unsigned char *p, *end;
    Huffman huff_dc[4], huff_ac[4];
    unsigned short dequant[4][64];
    JpegComp comp[4];
    int img_x = 0, img_y = 0, img_n = 0;
    int h_max = 1, v_max = 1, mcu_x = 0, mcu_y = 0, mcu_w = 0, mcu_h = 0;
    int restart_interval = 0, todo = 0;
    int scan_n = 0, order[4];
    int app14_transform = -1;
    bool jfif = false;
    // bit reader
    uint32_t code_buffer = 0;
    int code_bits = 0;
    unsigned char marker = 0xff;
    bool nomore = false;

    int get8() { return p < end ? *p++ : 0; }
    int get16() { int a = get8(); return (a << 8) | get8(); }

    void grow() {
        do {
            unsigned b = nomore ? 0 : (unsigned)get8();
            if (b == 0xff) {
                int c = get8();
                while (c == 0xff) c = get8();
                if (c != 0) { marker = (unsigned char)c; nomore = true; return; }
            }
            code_buffer |= b << (24 - code_bits);
            code_bits += 8;
        } while (code_bits <= 24);
    }
    int decode_symbol(const Huffman &h) {
        if (code_bits < 16) grow();
        unsigned temp = code_buffer >> 16;
        int k;
        for (k = 1; k <= 16; k++)
            if ((int)temp < h.maxcode[k]) break;
        if (k == 17 || k > code_bits) { code_bits -= 16; return -1; }
        int c = (int)((code_buffer >> (32 - k)) & ((1u << k) - 1)) + h.delta[k];
        if (c < 0 || c >= 256) return -1;
        code_bits -= k;
        code_buffer <<= k;
        return h.values[c];
    }
    // the n-bit value that follows a symbol, sign-extended the JPEG way
    int extend_receive(int n) {
        if (n == 0) return 0;
        if (code_bits < n) grow();
        int sgn = (int)(code_buffer >> 31);
        unsigned k = (code_buffer << n) | (code_buffer >> (32 - n));  // rotate left
        unsigned mask = (1u << n) - 1;
        code_buffer = k & ~mask;
        k &= mask;
        code_bits -= n;
        static const int bias[16] = {0, -1, -3, -7, -15, -31, -63, -127, -255, -511, -1023, -2047, -4095, -8191, -16383, -32767};
        return (int)k + (bias[n] & (sgn - 1));
    }
    bool decode_block(short data[64], JpegComp &c) {
        memset(data, 0, 64 * sizeof(short));
        int t = decode_symbol(huff_dc[c.hd]);
        if (t < 0 || t > 15) return false;
        int diff = t ? extend_receive(t) : 0;
        c.dc_pred += diff;
        data[0] = (short)(c.dc_pred * dequant[c.tq][0]);
        int k = 1;
        do {
            int rs = decode_symbol(huff_ac[c.ha]);
            if (rs < 0) return false;
            int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xf0) break;  // end of block
                k += 16;
            } else {
                k += r;
                unsigned zig = kDezigzag[k++];
                data[zig] = (short)(extend_receive(s) * dequant[c.tq][zig]);  // (the table is stored de-zigzagged)
            }
        } while (k < 64);
        return true;
    }
};

inline unsigned char clamp8(int x) { return (unsigned char)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// integer inverse DCT (the jidctint-style factorisation stb uses; 12-bit fixed-point constants)
#define LJ_F2F(x) ((int)(((x) * 4096 + 0.5)))
#define LJ_FSH(x) ((x) * 4096)
#define LJ_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)            \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;  \
    p2 = s2; p3 = s6;                                         \
    p1 = (p2 + p3) * LJ_F2F(0.5411961f);                      \
    t2 = p1 + p3 * LJ_F2F(-1.847759065f);                     \
    t3 = p1 + p2 * LJ_F2F(0.765366865f);                      \
    p2 = s0; p3 = s4;                                         \
    t0 = LJ_FSH(p2 + p3); t1 = LJ_FSH(p2 - p3);               \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;   \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                       \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;   \
    p5 = (p3 + p4) * LJ_F2F(1.175875602f);                    \
    t0 = t0 * LJ_F2F(0.298631336f);                           \
    t1 = t1 * LJ_F2F(2.053119869f);                           \
    t2 = t2 * LJ_F2F(3.072711026f);                           \
    t3 = t3 * LJ_F2F(1.501321110f);                           \
    p1 = p5 + p1 * LJ_F2F(-0.899976223f);                     \
    p2 = p5 + p2 * LJ_F2F(-2.562915447f);                     \
    p3 = p3 * LJ_F2F(-1.961570560f);                          \
    p4 = p4 * LJ_F2F(-0.390180644f);                          \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

void idct_block(unsigned char *out, int out_stride, const short data[64]) {
    int val[64], *v = val;
    const short *d = data;
    for (int i = 0; i < 8; ++i, ++d, ++v) {
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            int dcterm = d[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dcterm;
        } else {
            LJ_IDCT_1D(d[0], d[8], d[16], d[24], d[32], d[40], d[48], d[56])
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;
            v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
        }
    }
    v = val;
    unsigned char *o = out;
    for (int i = 0; i < 8; ++i, v += 8, o += out_stride) {
        LJ_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17);
        o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17);
        o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
    }
}

// chroma upsampling kernels: one output row from the two nearest input rows
void resample_row_1(unsigned char *out, const unsigned char *near, const unsigned char *, int w, int) { memcpy(out, near, (size_t)w); }
void resample_row_v2(unsigned char *out, const unsigned char *near, const unsigned char *far, int w, int) {
    for (int i = 0; i < w; i++) out[i] = (unsigned char)((3 * near[i] + far[i] + 2) >> 2);
}
void resample_row_h2(unsigned char *out, const unsigned char *in, const unsigned char *, int w, int) {
    if (w == 1) { out[0] = out[1] = in[0]; return; }
    out[0] = in[0];
    out[1] = (unsigned char)((in[0] * 3 + in[1] + 2) >> 2);
    int i;
    for (i = 1; i < w - 1; i++) {
        int n = 3 * in[i] + 2;
        out[i * 2] = (unsigned char)((n + in[i - 1]) >> 2);
        out[i * 2 + 1] = (unsigned char)((n + in[i + 1]) >> 2);
    }
    out[i * 2] = (unsigned char)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
    out[i * 2 + 1] = in[w - 1];
}
void resample_row_hv2(unsigned char *out, const unsigned char *near, const unsigned char *far, int w, int) {
    if (w == 1) { out[0] = out[1] = (unsigned char)((3 * near[0] + far[0] + 2) >> 2); return; }
    int t1 = 3 * near[0] + far[0];
    out[0] = (unsigned char)((t1 + 2) >> 2);
    for (int i = 1; i < w; i++) {
        int t0 = t1;
        t1 = 3 * near[i] + far[i];
        out[i * 2 - 1] = (unsigned char)((3 * t0 + t1 + 8) >> 4);
        out[i * 2] = (unsigned char)((3 * t1 + t0 + 8) >> 4);
    }
    out[w * 2 - 1] = (unsigned char)((t1 + 2) >> 2);
}
void resample_row_generic(unsigned char *out, const unsigned char *near, const unsigned char *, int w, int hs) {
    for (int i = 0; i < w; i++)
        for (int j = 0; j < hs; j++) out[i * hs + j] = near[i];
}

// YCbCr -> RGB for one row.  stb runs groups of 8 pixels through a 16-bit SSE2 path whose rounding differs from the
// scalar 20-bit fixed-point code used for the last (count % 8) pixels; both are reproduced (the reference's textures
// are what an x86-64 build of stb_image produced).
inline int mulhi16(int a, int b) { return (int)(((int32_t)(int16_t)a * (int32_t)(int16_t)b) >> 16); }
void ycbcr_to_rgb_row(unsigned char *out, const unsigned char *y, const unsigned char *pcb, const unsigned char *pcr, int count, int step) {
    int i = 0;
    const int cr_const0 = (short)(1.40200f * 4096.0f + 0.5f), cr_const1 = -(short)(0.71414f * 4096.0f + 0.5f);
    const int cb_const0 = -(short)(0.34414f * 4096.0f + 0.5f), cb_const1 = (short)(1.77200f * 4096.0f + 0.5f);
    if (step == 4 || step == 3) {
        for (; i + 7 < count; i += 8) {
            for (int k = 0; k < 8; k++) {
                int yws = (((int)y[i + k] << 8) | 128) >> 4;  // logical shift of the unsigned 16-bit word
                int crw = (int)(int16_t)((((int)pcr[i + k] ^ 0x80) & 0xff) << 8);
                int cbw = (int)(int16_t)((((int)pcb[i + k] ^ 0x80) & 0xff) << 8);
                int cr0 = mulhi16(cr_const0, crw), cb0 = mulhi16(cb_const0, cbw);
                int cb1 = mulhi16(cbw, cb_const1), cr1 = mulhi16(crw, cr_const1);
                int rws = (int16_t)(cr0 + yws), gws = (int16_t)((int16_t)(cb0 + yws) + cr1), bws = (int16_t)(yws + cb1);
                out[0] = clamp8(rws >> 4); out[1] = clamp8(gws >> 4); out[2] = clamp8(bws >> 4);
                if (step == 4) out[3] = 255;
                out += step;
            }
        }
    }
    for (; i < count; ++i) {
        int y_fixed = (y[i] << 20) + (1 << 19);
        int cr = pcr[i] - 128, cb = pcb[i] - 128;
        int r = y_fixed + cr * (((int)(1.40200f * 4096.0f + 0.5f)) << 8);
        int g = y_fixed + (cr * -(((int)(0.71414f * 4096.0f + 0.5f)) << 8)) + ((cb * -(((int)(0.34414f * 4096.0f + 0.5f)) << 8)) & 0xffff0000);
        int b = y_fixed + cb * (((int)(1.77200f * 4096.0f + 0.5f)) << 8);
        r >>= 20; g >>= 20; b >>= 20;
        out[0] = clamp8(r); out[1] = clamp8(g); out[2] = clamp8(b);
        if (step == 4) out[3] = 255;
        out += step;
    }
}

unsigned char compute_y(int r, int g, int b) { return (unsigned char)(((r * 77) + (g * 150) + (29 * b)) >> 8); }

}  // namespace

std::vector<unsigned char> decode_jpeg(const std::vector<unsigned char> &file, int &w, int &h, int &comps, int want_comps) {
    JpegDecoder z;
    memset(z.dequant, 0, sizeof(z.dequant));
    z.p = file.data();
    z.end = file.data() + file.size();
    if (z.get8() != 0xff || z.get8() != 0xd8) fail("not a JPEG file");
    bool have_frame = false, decoded = false;
    auto next_marker = [&]() -> int {
        if (z.marker != 0xff) { int m = z.marker; z.marker = 0xff; return m; }
        int x = z.get8();
        if (x != 0xff) return 0xff;
        while (x == 0xff) x = z.get8();
        return x;
    };
    auto reset = [&]() {
        z.code_bits = 0; z.code_buffer = 0; z.nomore = false;
        for (int i = 0; i < 4; i++) z.comp[i].dc_pred = 0;
        z.marker = 0xff;
        z.todo = z.restart_interval ? z.restart_interval : 0x7fffffff;
    };
    for (;;) {
        int m = next_marker();
        if (m == 0xff) { if (z.p >= z.end) break; continue; }
        if (m == 0xd9) break;  // EOI
        if (m == 0xc2) fail("progressive JPEG is not supported");
        if (m == 0xc0 || m == 0xc1) {
            int len = z.get16();
            if (z.get8() != 8) fail("JPEG: only 8-bit samples are supported");
            z.img_y = z.get16(); z.img_x = z.get16(); z.img_n = z.get8();
            if (z.img_n != 1 && z.img_n != 3) fail("JPEG: unsupported component count");
            if (len != 8 + 3 * z.img_n) fail("JPEG: bad SOF length");
            for (int i = 0; i < z.img_n; i++) {
                z.comp[i].id = z.get8();
                int q = z.get8();
                z.comp[i].h = q >> 4; z.comp[i].v = q & 15;
                z.comp[i].tq = z.get8();
                if (!z.comp[i].h || z.comp[i].h > 4 || !z.comp[i].v || z.comp[i].v > 4 || z.comp[i].tq > 3) fail("JPEG: bad component header");
                z.h_max = std::max(z.h_max, z.comp[i].h);
                z.v_max = std::max(z.v_max, z.comp[i].v);
            }
            z.mcu_w = z.h_max * 8; z.mcu_h = z.v_max * 8;
            z.mcu_x = (z.img_x + z.mcu_w - 1) / z.mcu_w;
            z.mcu_y = (z.img_y + z.mcu_h - 1) / z.mcu_h;
            for (int i = 0; i < z.img_n; i++) {
                JpegComp &c = z.comp[i];
                c.x = (z.img_x * c.h + z.h_max - 1) / z.h_max;
                c.y = (z.img_y * c.v + z.v_max - 1) / z.v_max;
                c.w2 = z.mcu_x * c.h * 8;
                c.h2 = z.mcu_y * c.v * 8;
                c.data.assign((size_t)c.w2 * c.h2, 0);
            }
            have_frame = true;
        } else if (m == 0xc4) {  // DHT
            int len = z.get16() - 2;
            while (len > 0) {
                int q = z.get8(), tc = q >> 4, th = q & 15, sizes[16], n = 0;
                if (tc > 1 || th > 3) fail("JPEG: bad DHT header");
                for (int i = 0; i < 16; i++) { sizes[i] = z.get8(); n += sizes[i]; }
                if (n > 256) fail("JPEG: bad DHT header");
                Huffman &hf = tc == 0 ? z.huff_dc[th] : z.huff_ac[th];
                if (!hf.build(sizes)) fail("JPEG: bad code lengths");
                for (int i = 0; i < n; i++) hf.values[i] = (unsigned char)z.get8();
                len -= 17 + n;
            }
        } else if (m == 0xdb) {  // DQT
            int len = z.get16() - 2;
            while (len > 0) {
                int q = z.get8(), prec = q >> 4, t = q & 15;
                if (t > 3) fail("JPEG: bad DQT table");
                for (int i = 0; i < 64; i++) z.dequant[t][kDezigzag[i]] = (unsigned short)(prec ? z.get16() : z.get8());
                len -= prec ? 129 : 65;
            }
        } else if (m == 0xdd) {  // DRI
            z.get16();
            z.restart_interval = z.get16();
        } else if (m == 0xda) {  // SOS: one interleaved scan with all components (baseline files)
            if (!have_frame) fail("JPEG: scan before frame header");
            z.get16();
            z.scan_n = z.get8();
            if (z.scan_n < 1 || z.scan_n > z.img_n) fail("JPEG: bad SOS component count");
            for (int i = 0; i < z.scan_n; i++) {
                int id = z.get8(), q = z.get8(), which;
                for (which = 0; which < z.img_n; which++) if (z.comp[which].id == id) break;
                if (which == z.img_n) fail("JPEG: bad SOS component");
                z.comp[which].hd = q >> 4; z.comp[which].ha = q & 15;
                z.order[i] = which;
            }
            z.get8(); z.get8(); z.get8();  // spectral selection / successive approximation: fixed for baseline
            reset();
            short block[64];
            if (z.scan_n == 1) {
                JpegComp &c = z.comp[z.order[0]];
                int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
                for (int j = 0; j < bh; j++)
                    for (int i = 0; i < bw; i++) {
                        if (!z.decode_block(block, c)) fail("JPEG: bad huffman code");
                        idct_block(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, block);
                        if (--z.todo <= 0) {
                            if (z.code_bits < 24) z.grow();
                            if (!(z.marker >= 0xd0 && z.marker <= 0xd7)) goto scan_done;
                            reset();
                        }
                    }
            } else {
                for (int j = 0; j < z.mcu_y; j++)
                    for (int i = 0; i < z.mcu_x; i++) {
                        for (int k = 0; k < z.scan_n; k++) {
                            JpegComp &c = z.comp[z.order[k]];
                            for (int y = 0; y < c.v; y++)
                                for (int x = 0; x < c.h; x++) {
                                    int x2 = (i * c.h + x) * 8, y2 = (j * c.v + y) * 8;
                                    if (!z.decode_block(block, c)) fail("JPEG: bad huffman code");
                                    idct_block(c.data.data() + (size_t)c.w2 * y2 + x2, c.w2, block);
                                }
                        }
                        if (--z.todo <= 0) {
                            if (z.code_bits < 24) z.grow();
                            if (!(z.marker >= 0xd0 && z.marker <= 0xd7)) goto scan_done;
                            reset();
                        }
                    }
            }
        scan_done:
            decoded = true;
            if (z.marker == 0xff) {
                // skip any stuffing up to the next marker
                while (z.p < z.end) {
                    int x = z.get8();
                    if (x == 0xff) { int y = z.get8(); while (y == 0xff) y = z.get8(); if (y != 0) { z.marker = (unsigned char)y; break; } }
                }
            }
        } else if (m == 0xee) {  // APP14 "Adobe": colour transform flag
            int len = z.get16() - 2;
            const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
            bool ok = len >= 12;
            int k = 0;
            if (ok) {
                for (; k < 6; k++) if (z.get8() != tag[k]) { ok = false; k++; break; }
            }
            if (ok) { z.get8(); z.get16(); z.get16(); z.app14_transform = z.get8(); k = 12; }
            for (; k < len; k++) z.get8();
        } else if ((m >= 0xe0 && m <= 0xef) || m == 0xfe) {
            int len = z.get16() - 2;
            if (m == 0xe0 && len >= 5) {
                const char tag[5] = {'J', 'F', 'I', 'F', 0};
                bool ok = true;
                for (int k = 0; k < 5; k++) if (z.get8() != tag[k]) ok = false;
                z.jfif = ok;
                len -= 5;
            }
            for (int k = 0; k < len; k++) z.get8();
        } else if (m >= 0xd0 && m <= 0xd7) {
            // stray restart marker
        } else {
            int len = z.get16() - 2;
            for (int k = 0; k < len; k++) z.get8();
        }
    }
    if (!decoded) fail("JPEG: no image data");
    w = z.img_x; h = z.img_y;
    // three components are YCbCr unless the component ids spell RGB or an Adobe marker says "no transform"
    bool is_rgb = z.img_n == 3 && ((z.comp[0].id == 'R' && z.comp[1].id == 'G' && z.comp[2].id == 'B') || (z.app14_transform == 0 && !z.jfif));
    int n = want_comps ? want_comps : (z.img_n >= 3 ? 3 : 1);
    comps = n;
    int decode_n = (z.img_n == 3 && n < 3 && !is_rgb) ? 1 : z.img_n;
    typedef void (*ResampleFn)(unsigned char *, const unsigned char *, const unsigned char *, int, int);
    struct Resample { ResampleFn fn; const unsigned char *line0, *line1; int hs, vs, w_lores, ystep, ypos; std::vector<unsigned char> linebuf; } rs[4];
    for (int k = 0; k < decode_n; k++) {
        Resample &r = rs[k];
        r.linebuf.assign((size_t)z.img_x + 3, 0);
        r.hs = z.h_max / z.comp[k].h; r.vs = z.v_max / z.comp[k].v;
        r.ystep = r.vs >> 1;
        r.w_lores = (z.img_x + r.hs - 1) / r.hs;
        r.ypos = 0;
        r.line0 = r.line1 = z.comp[k].data.data();
        if (r.hs == 1 && r.vs == 1) r.fn = resample_row_1;
        else if (r.hs == 1 && r.vs == 2) r.fn = resample_row_v2;
        else if (r.hs == 2 && r.vs == 1) r.fn = resample_row_h2;
        else if (r.hs == 2 && r.vs == 2) r.fn = resample_row_hv2;
        else r.fn = resample_row_generic;
    }
    std::vector<unsigned char> output((size_t)n * z.img_x * z.img_y);
    const unsigned char *coutput[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int j = 0; j < z.img_y; j++) {
        unsigned char *out = output.data() + (size_t)n * z.img_x * j;
        for (int k = 0; k < decode_n; k++) {
            Resample &r = rs[k];
            bool y_bot = r.ystep >= (r.vs >> 1);
            const unsigned char *near = y_bot ? r.line1 : r.line0, *far = y_bot ? r.line0 : r.line1;
            if (r.fn == resample_row_1) coutput[k] = near;
            else { r.fn(r.linebuf.data(), near, far, r.w_lores, r.hs); coutput[k] = r.linebuf.data(); }
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < z.comp[k].y) r.line1 += z.comp[k].w2;
            }
        }
        if (n >= 3) {
            const unsigned char *y = coutput[0];
            if (z.img_n == 3) {
                if (is_rgb) for (int i = 0; i < z.img_x; i++) { out[0] = y[i]; out[1] = coutput[1][i]; out[2] = coutput[2][i]; if (n == 4) out[3] = 255; out += n; }
                else ycbcr_to_rgb_row(out, y, coutput[1], coutput[2], z.img_x, n);
            } else {
                for (int i = 0; i < z.img_x; i++) { out[0] = out[1] = out[2] = y[i]; if (n == 4) out[3] = 255; out += n; }
            }
        } else {
            if (is_rgb) {
                for (int i = 0; i < z.img_x; i++) { out[0] = compute_y(coutput[0][i], coutput[1][i], coutput[2][i]); if (n == 2) out[1] = 255; out += n; }
            } else {
                const unsigned char *y = coutput[0];
                for (int i = 0; i < z.img_x; i++) { out[0] = y[i]; if (n == 2) out[1] = 255; out += n; }
            }
        }
    }
    return output;
}

// ------------------------------------------------------------------------------------------------ zlib helpers
namespace {
std::vector<unsigned char> zlib_inflate(const unsigned char *src, size_t n, size_t expected) {
    std::vector<unsigned char> out(expected ? expected : n * 4 + 64);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) fail("inflateInit failed");
    zs.next_in = const_cast<unsigned char *>(src);
    zs.avail_in = (uInt)n;
    size_t produced = 0;
    for (;;) {
        if (produced == out.size()) {
            if (expected) { // output full: fine if the stream is done
                unsigned char dummy;
                zs.next_out = &dummy; zs.avail_out = 1;
                int r = inflate(&zs, Z_NO_FLUSH);
                if (r == Z_STREAM_END || zs.avail_out == 1) break;
                inflateEnd(&zs);
                fail("zlib stream longer than expected");
            }
            out.resize(out.size() * 2);
        }
        zs.next_out = out.data() + produced;
        zs.avail_out = (uInt)(out.size() - produced);
        int r = inflate(&zs, Z_NO_FLUSH);
        produced = out.size() - zs.avail_out;
        if (r == Z_STREAM_END) break;
        if (r != Z_OK) { inflateEnd(&zs); fail("zlib inflate failed"); }
        if (zs.avail_in == 0 && zs.avail_out != 0) break;
    }
    inflateEnd(&zs);
    out.resize(produced);
    return out;
}
std::vector<unsigned char> zlib_deflate(const unsigned char *src, size_t n) {
    uLongf bound = compressBound((uLong)n);
    std::vector<unsigned char> out(bound);
    if (compress2(out.data(), &bound, src, (uLong)n, Z_DEFAULT_COMPRESSION) != Z_OK) fail("zlib deflate failed");
    out.resize(bound);
    return out;
}
}  // namespace

// ------------------------------------------------------------------------------------------------ PNG
std::vector<unsigned char> decode_png(const std::vector<unsigned char> &file, int &w, int &h, int &comps) {
    static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) fail("not a PNG file");
    size_t p = 8;
    auto be32 = [&](size_t o) { return ((uint32_t)file[o] << 24) | ((uint32_t)file[o + 1] << 16) | ((uint32_t)file[o + 2] << 8) | file[o + 3]; };
    int depth = 0, color = 0, interlace = 0;
    std::vector<unsigned char> idat, palette, trns;
    w = h = 0;
    while (p + 8 <= file.size()) {
        uint32_t len = be32(p);
        std::string type((const char *)&file[p + 4], 4);
        size_t d = p + 8;
        if (d + len > file.size()) fail("PNG: truncated chunk");
        if (type == "IHDR") {
            w = (int)be32(d); h = (int)be32(d + 4); depth = file[d + 8]; color = file[d + 9]; interlace = file[d + 12];
        } else if (type == "PLTE") palette.assign(file.begin() + d, file.begin() + d + len);
        else if (type == "tRNS") trns.assign(file.begin() + d, file.begin() + d + len);
        else if (type == "IDAT") idat.insert(idat.end(), file.begin() + d, file.begin() + d + len);
        else if (type == "IEND") break;
        p = d + len + 4;
    }
    if (w <= 0 || h <= 0) fail("PNG: no IHDR");
    if (interlace) fail("PNG: interlaced files are not supported");
    if (depth != 8 && depth != 16 && !(color == 3 || color == 0)) fail("PNG: unsupported bit depth");
    int ch = color == 0 ? 1 : (color == 2 ? 3 : (color == 3 ? 1 : (color == 4 ? 2 : 4)));
    int bpp_bits = ch * depth;
    size_t stride = ((size_t)w * bpp_bits + 7) / 8;
    std::vector<unsigned char> raw = zlib_inflate(idat.data(), idat.size(), (stride + 1) * h);
    if (raw.size() < (stride + 1) * h) fail("PNG: not enough image data");
    int bpp = std::max(1, bpp_bits / 8);
    std::vector<unsigned char> img(stride * h);
    for (int y = 0; y < h; y++) {
        const unsigned char *src = &raw[(stride + 1) * y];
        unsigned char *cur = &img[stride * y];
        const unsigned char *prev = y ? &img[stride * (y - 1)] : nullptr;
        int filter = src[0];
        src++;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0, v = src[i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: { int pa = abs(b - c), pb = abs(a - c), pc = abs(a + b - 2 * c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: fail("PNG: bad filter");
            }
            cur[i] = (unsigned char)v;
        }
    }
    // to 8-bit channels
    std::vector<unsigned char> out;
    auto sample = [&](int y, int i) -> int {  // i-th sample of row y, scaled to 8 bits
        const unsigned char *row = &img[stride * y];
        if (depth == 8) return row[i];
        if (depth == 16) return row[2 * i];
        int per = 8 / depth, v = (row[i / per] >> (8 - depth * (i % per + 1))) & ((1 << depth) - 1);
        return color == 3 ? v : v * (255 / ((1 << depth) - 1));
    };
    if (color == 3) {
        comps = trns.empty() ? 3 : 4;
        out.resize((size_t)w * h * comps);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int idx = sample(y, x);
                for (int k = 0; k < 3; k++) out[((size_t)y * w + x) * comps + k] = (size_t)(3 * idx + k) < palette.size() ? palette[3 * idx + k] : 0;
                if (comps == 4) out[((size_t)y * w + x) * 4 + 3] = (size_t)idx < trns.size() ? trns[idx] : 255;
            }
    } else {
        comps = ch;
        out.resize((size_t)w * h * ch);
        for (int y = 0; y < h; y++)
            for (int i = 0; i < w * ch; i++) out[(size_t)y * w * ch + i] = (unsigned char)sample(y, i);
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ Radiance HDR
ImageF decode_hdr(const std::vector<unsigned char> &file) {
    size_t p = 0;
    auto line = [&]() { std::string s; while (p < file.size() && file[p] != '\n') s.push_back((char)file[p++]); p++; return s; };
    std::string first = line();
    if (first != "#?RADIANCE" && first != "#?RGBE") fail("not a Radiance HDR file");
    bool ok = false;
    for (;;) { std::string s = line(); if (s.empty()) break; if (s == "FORMAT=32-bit_rle_rgbe") ok = true; }
    if (!ok) fail("HDR: unsupported format");
    std::string dims = line();
    int h = 0, w = 0;
    if (sscanf(dims.c_str(), "-Y %d +X %d", &h, &w) != 2) fail("HDR: unsupported orientation");
    ImageF img;
    img.width = w; img.height = h; img.channels = 3;
    img.data.resize((size_t)w * h * 3);
    auto convert = [](float *out, const unsigned char *in) {
        if (in[3] != 0) { float f = (float)ldexp(1.0f, in[3] - (int)(128 + 8)); out[0] = in[0] * f; out[1] = in[1] * f; out[2] = in[2] * f; }
        else out[0] = out[1] = out[2] = 0;
    };
    std::vector<unsigned char> scan((size_t)w * 4);
    for (int j = 0; j < h; j++) {
        bool rle = w >= 8 && w < 32768 && p + 4 <= file.size() && file[p] == 2 && file[p + 1] == 2 && !(file[p + 2] & 0x80) && ((file[p + 2] << 8) | file[p + 3]) == w;
        if (!rle) {
            for (int i = 0; i < w; i++) { if (p + 4 > file.size()) fail("HDR: truncated"); convert(&img.data[((size_t)j * w + i) * 3], &file[p]); p += 4; }
            continue;
        }
        p += 4;
        for (int k = 0; k < 4; k++) {
            int i = 0;
            while (i < w) {
                if (p >= file.size()) fail("HDR: truncated");
                int count = file[p++];
                if (count > 128) { count -= 128; if (p >= file.size() || i + count > w) fail("HDR: corrupt"); unsigned char v = file[p++]; for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = v; }
                else { if (count == 0 || i + count > w || p + count > file.size()) fail("HDR: corrupt"); for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = file[p++]; }
            }
        }
        for (int i = 0; i < w; i++) convert(&img.data[((size_t)j * w + i) * 3], &scan[(size_t)i * 4]);
    }
    return img;
}

// ------------------------------------------------------------------------------------------------ half floats
unsigned short float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t e = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t m = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (unsigned short)(sign | 0x7c00u | (m ? 0x200u | (m >> 13) : 0));  // inf / nan
    if (e >= 31) return (unsigned short)(sign | 0x7c00u);  // overflow -> inf
    if (e <= 0) {
        if (e < -10) return (unsigned short)sign;  // underflow -> 0
        m |= 0x800000u;
        int shift = 14 - e;
        uint32_t r = m >> shift, rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (r & 1))) r++;
        return (unsigned short)(sign | r);
    }
    uint32_t r = (uint32_t)(e << 10) | (m >> 13), rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) r++;  // round to nearest even (may carry into the exponent)
    return (unsigned short)(sign | r);
}
float half_to_float(unsigned short h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ffu, x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { int k = 0; while (!(m & 0x400u)) { m <<= 1; k++; } m &= 0x3ffu; x = sign | (uint32_t)(127 - 15 - k + 1) << 23 | (m << 13); }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e - 15 + 127) << 23) | (m << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

// ------------------------------------------------------------------------------------------------ OpenEXR
namespace {

// --- PIZ: Huffman coding of 16-bit symbols (OpenEXR ImfHuf: canonical codes up to 58 bits, a 14-bit fast table, run-length symbol)
constexpr int HUF_ENCBITS = 16, HUF_DECBITS = 14, HUF_ENCSIZE = (1 << HUF_ENCBITS) + 1, HUF_DECSIZE = 1 << HUF_DECBITS, HUF_DECMASK = HUF_DECSIZE - 1;
constexpr int SHORT_ZEROCODE_RUN = 59, LONG_ZEROCODE_RUN = 63, SHORTEST_LONG_RUN = 2 + LONG_ZEROCODE_RUN - SHORT_ZEROCODE_RUN;

struct HufDec { int len = 0, lit = 0; std::vector<int> p; };

inline uint64_t huf_get_bits(int nbits, uint64_t &c, int &lc, const unsigned char *&in, const unsigned char *end) {
    while (lc < nbits) { if (in >= end) fail("EXR: PIZ table truncated"); c = (c << 8) | *in++; lc += 8; }
    lc -= nbits;
    return (c >> lc) & ((1ull << nbits) - 1);
}
void huf_canonical_table(std::vector<uint64_t> &hcode) {
    uint64_t n[59];
    memset(n, 0, sizeof(n));
    for (int i = 0; i < HUF_ENCSIZE; i++) n[hcode[i]] += 1;
    uint64_t c = 0;
    for (int i = 58; i > 0; --i) { uint64_t nc = (c + n[i]) >> 1; n[i] = c; c = nc; }
    for (int i = 0; i < HUF_ENCSIZE; i++) { int l = (int)hcode[i]; if (l > 0) hcode[i] = (uint64_t)l | (n[l]++ << 6); }
}
void huf_unpack_table(const unsigned char *&in, const unsigned char *end, int im, int iM, std::vector<uint64_t> &hcode) {
    hcode.assign(HUF_ENCSIZE, 0);
    uint64_t c = 0;
    int lc = 0;
    for (; im <= iM; im++) {
        uint64_t l = hcode[im] = huf_get_bits(6, c, lc, in, end);
        if (l == (uint64_t)LONG_ZEROCODE_RUN) {
            int zerun = (int)huf_get_bits(8, c, lc, in, end) + SHORTEST_LONG_RUN;
            if (im + zerun > iM + 1) fail("EXR: PIZ table overrun");
            while (zerun--) hcode[im++] = 0;
            im--;
        } else if (l >= (uint64_t)SHORT_ZEROCODE_RUN) {
            int zerun = (int)l - SHORT_ZEROCODE_RUN + 2;
            if (im + zerun > iM + 1) fail("EXR: PIZ table overrun");
            while (zerun--) hcode[im++] = 0;
            im--;
        }
    }
    huf_canonical_table(hcode);
}
void huf_build_dec(const std::vector<uint64_t> &hcode, int im, int iM, std::vector<HufDec> &hdec) {
    hdec.assign(HUF_DECSIZE, HufDec());
    for (; im <= iM; im++) {
        uint64_t c = hcode[im] >> 6;
        int l = (int)(hcode[im] & 63);
        if (c >> l) fail("EXR: PIZ bad code");
        if (l > HUF_DECBITS) {
            HufDec &pl = hdec[c >> (l - HUF_DECBITS)];
            if (pl.len) fail("EXR: PIZ bad code table");
            pl.lit++;
            pl.p.push_back(im);
        } else if (l) {
            size_t base = (size_t)(c << (HUF_DECBITS - l));
            for (uint64_t i = 1ull << (HUF_DECBITS - l); i > 0; i--, base++) {
                HufDec &pl = hdec[base];
                if (pl.len || !pl.p.empty()) fail("EXR: PIZ bad code table");
                pl.len = l;
                pl.lit = im;
            }
        }
    }
}
void huf_decode(const std::vector<uint64_t> &hcode, const std::vector<HufDec> &hdec, const unsigned char *in, const unsigned char *in_end, int ni, int rlc, size_t no, unsigned short *out) {
    uint64_t c = 0;
    int lc = 0;
    unsigned short *outb = out, *oe = out + no;
    const unsigned char *ie = in + (ni + 7) / 8;
    if (ie > in_end) fail("EXR: PIZ data truncated");
    auto get_char = [&]() { c = (c << 8) | *in++; lc += 8; };
    auto get_code = [&](int po) {
        if (po == rlc) {
            if (lc < 8) { if (in >= ie) fail("EXR: PIZ data truncated"); get_char(); }
            lc -= 8;
            int cs = (int)((c >> lc) & 0xff);
            if (out + cs > oe || out - 1 < outb) fail("EXR: PIZ run overflows");
            unsigned short s = out[-1];
            while (cs-- > 0) *out++ = s;
        } else if (out < oe) {
            *out++ = (unsigned short)po;
        } else fail("EXR: PIZ output overflows");
    };
    while (in < ie) {
        get_char();
        while (lc >= HUF_DECBITS) {
            const HufDec &pl = hdec[(c >> (lc - HUF_DECBITS)) & HUF_DECMASK];
            if (pl.len) {
                lc -= pl.len;
                get_code(pl.lit);
            } else {
                if (pl.p.empty()) fail("EXR: PIZ bad code");
                int j;
                for (j = 0; j < pl.lit; j++) {
                    int l = (int)(hcode[pl.p[j]] & 63);
                    while (lc < l && in < ie) get_char();
                    if (lc >= l && (hcode[pl.p[j]] >> 6) == ((c >> (lc - l)) & ((1ull << l) - 1))) {
                        lc -= l;
                        get_code(pl.p[j]);
                        break;
                    }
                }
                if (j == pl.lit) fail("EXR: PIZ bad code");
            }
        }
    }
    int i = (8 - ni) & 7;
    c >>= i;
    lc -= i;
    while (lc > 0) {
        const HufDec &pl = hdec[(c << (HUF_DECBITS - lc)) & HUF_DECMASK];
        if (!pl.len) fail("EXR: PIZ bad code");
        lc -= pl.len;
        get_code(pl.lit);
    }
    if ((size_t)(out - outb) != no) fail("EXR: PIZ short output");
}
void huf_uncompress(const unsigned char *data, size_t n, unsigned short *raw, size_t nraw) {
    if (n < 20) { if (nraw) fail("EXR: PIZ block too short"); return; }
    auto u32 = [&](size_t o) { return (uint32_t)data[o] | ((uint32_t)data[o + 1] << 8) | ((uint32_t)data[o + 2] << 16) | ((uint32_t)data[o + 3] << 24); };
    int im = (int)u32(0), iM = (int)u32(4), nbits = (int)u32(12);
    if (im < 0 || im >= HUF_ENCSIZE || iM < 0 || iM >= HUF_ENCSIZE) fail("EXR: PIZ bad header");
    const unsigned char *ptr = data + 20, *end = data + n;
    std::vector<uint64_t> freq;
    std::vector<HufDec> hdec;
    huf_unpack_table(ptr, end, im, iM, freq);
    if (nbits > 8 * (int)(end - ptr)) fail("EXR: PIZ bad bit count");
    huf_build_dec(freq, im, iM, hdec);
    huf_decode(freq, hdec, ptr, end, nbits, iM, nraw, raw);
}

// --- PIZ: inverse 2-D Haar-like wavelet (OpenEXR ImfWav), 14-bit and 16-bit variants
inline void wdec14(unsigned short l, unsigned short h, unsigned short &a, unsigned short &b) {
    short ls = (short)l, hs = (short)h;
    int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
    a = (unsigned short)(short)ai;
    b = (unsigned short)(short)(ai - hi);
}
inline void wdec16(unsigned short l, unsigned short h, unsigned short &a, unsigned short &b) {
    int m = l, d = h, bb = (m - (d >> 1)) & 0xffff, aa = (d + bb - (1 << 15)) & 0xffff;
    b = (unsigned short)bb;
    a = (unsigned short)aa;
}
void wav2_decode(unsigned short *in, int nx, int ox, int ny, int oy, unsigned short mx) {
    bool w14 = mx < (1 << 14);
    int n = nx > ny ? ny : nx, p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1; p2 = p; p >>= 1;
    while (p >= 1) {
        unsigned short *py = in, *ey = in + oy * (ny - p2);
        int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
        unsigned short i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            unsigned short *px = py, *ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                unsigned short *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
                if (w14) { wdec14(*px, *p10, i00, i10); wdec14(*p01, *p11, i01, i11); wdec14(i00, i01, *px, *p01); wdec14(i10, i11, *p10, *p11); }
                else { wdec16(*px, *p10, i00, i10); wdec16(*p01, *p11, i01, i11); wdec16(i00, i01, *px, *p01); wdec16(i10, i11, *p10, *p11); }
            }
            if (nx & p) {
                unsigned short *p10 = px + oy1;
                if (w14) wdec14(*px, *p10, i00, *p10); else wdec16(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {
            unsigned short *px = py, *ex = py + ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                unsigned short *p01 = px + ox1;
                if (w14) wdec14(*px, *p01, i00, *p01); else wdec16(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

struct ExrChannel { std::string name; int type = 1, xs = 1, ys = 1; };

// undo the byte predictor and the even/odd byte interleave of the ZIP / RLE codecs
std::vector<unsigned char> exr_unpredict(const std::vector<unsigned char> &t) {
    std::vector<unsigned char> tmp = t;
    for (size_t i = 1; i < tmp.size(); i++) tmp[i] = (unsigned char)(tmp[i - 1] + tmp[i] - 128);
    std::vector<unsigned char> out(tmp.size());
    size_t half = (tmp.size() + 1) / 2;
    for (size_t i = 0, s = 0; s < tmp.size(); i++) {
        out[s++] = tmp[i];
        if (s < tmp.size()) out[s++] = tmp[half + i];
    }
    return out;
}

}  // namespace

ImageF decode_exr(const std::vector<unsigned char> &file) {
    if (file.size() < 8 || file[0] != 0x76 || file[1] != 0x2f || file[2] != 0x31 || file[3] != 0x01) fail("not an OpenEXR file");
    uint32_t version = file[4] | (file[5] << 8) | (file[6] << 16) | (file[7] << 24);
    if (version & 0x200) fail("EXR: tiled files are not supported");
    if (version & 0x1800) fail("EXR: multi-part / deep files are not supported");
    size_t p = 8;
    auto cstr = [&]() { std::string s; while (p < file.size() && file[p]) s.push_back((char)file[p++]); p++; return s; };
    auto i32 = [&](size_t o) { return (int32_t)((uint32_t)file[o] | ((uint32_t)file[o + 1] << 8) | ((uint32_t)file[o + 2] << 16) | ((uint32_t)file[o + 3] << 24)); };
    std::vector<ExrChannel> channels;
    int compression = 0, xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    for (;;) {
        std::string name = cstr();
        if (name.empty()) break;
        std::string type = cstr();
        int size = i32(p);
        p += 4;
        size_t v = p;
        if (v + size > file.size()) fail("EXR: truncated header");
        if (name == "channels") {
            size_t q = v;
            while (q < v + size && file[q]) {
                ExrChannel c;
                while (file[q]) c.name.push_back((char)file[q++]);
                q++;
                c.type = i32(q); q += 4 + 4;  // pixel type, pLinear + reserved
                c.xs = i32(q); c.ys = i32(q + 4); q += 8;
                channels.push_back(c);
            }
        } else if (name == "compression") compression = file[v];
        else if (name == "dataWindow") { xmin = i32(v); ymin = i32(v + 4); xmax = i32(v + 8); ymax = i32(v + 12); }
        p = v + size;
    }
    int w = xmax - xmin + 1, h = ymax - ymin + 1;
    if (w <= 0 || h <= 0 || channels.empty()) fail("EXR: bad header");
    for (auto &c : channels) if (c.xs != 1 || c.ys != 1) fail("EXR: subsampled channels are not supported");
    int lines_per_block;
    switch (compression) {
        case 0: case 1: case 2: lines_per_block = 1; break;
        case 3: lines_per_block = 16; break;
        case 4: lines_per_block = 32; break;
        default: fail("EXR: unsupported compression " + std::to_string(compression));
    }
    int nblocks = (h + lines_per_block - 1) / lines_per_block;
    std::vector<uint64_t> offsets(nblocks);
    for (int i = 0; i < nblocks; i++) { uint64_t o = 0; for (int k = 0; k < 8; k++) o |= (uint64_t)file[p + k] << (8 * k); offsets[i] = o; p += 8; }
    std::vector<int> csize(channels.size());
    size_t line_bytes = 0;
    for (size_t c = 0; c < channels.size(); c++) { csize[c] = channels[c].type == 1 ? 2 : 4; line_bytes += (size_t)csize[c] * w; }
    // channel -> RGBA slot (tinyexr LoadEXR: looks the channels up by name; a single channel fills R, G, B)
    int idx_r = -1, idx_g = -1, idx_b = -1, idx_a = -1;
    for (size_t c = 0; c < channels.size(); c++) {
        if (channels[c].name == "R") idx_r = (int)c; else if (channels[c].name == "G") idx_g = (int)c;
        else if (channels[c].name == "B") idx_b = (int)c; else if (channels[c].name == "A") idx_a = (int)c;
    }
    if (channels.size() == 1) idx_r = idx_g = idx_b = 0;
    if (idx_r < 0 || idx_g < 0 || idx_b < 0) fail("EXR: R, G, B channels not found");
    ImageF img;
    img.width = w; img.height = h; img.channels = 4;
    img.data.assign((size_t)w * h * 4, 1.0f);
    for (int blk = 0; blk < nblocks; blk++) {
        size_t o = (size_t)offsets[blk];
        if (o + 8 > file.size()) fail("EXR: bad chunk offset");
        int y0 = i32(o) - ymin, dsize = i32(o + 4);
        const unsigned char *src = &file[o + 8];
        if (o + 8 + (size_t)dsize > file.size() || y0 < 0 || y0 >= h) fail("EXR: bad chunk");
        int nlines = std::min(lines_per_block, h - y0);
        size_t raw_size = line_bytes * nlines;
        std::vector<unsigned char> raw;
        if ((size_t)dsize == raw_size || compression == 0) {
            raw.assign(src, src + dsize);
        } else if (compression == 2 || compression == 3) {
            raw = exr_unpredict(zlib_inflate(src, (size_t)dsize, raw_size));
        } else if (compression == 1) {
            std::vector<unsigned char> t;
            for (int i = 0; i < dsize;) {
                int count = (signed char)src[i++];
                if (count < 0) { for (int k = 0; k < -count && i < dsize; k++) t.push_back(src[i++]); }
                else { unsigned char v = src[i++]; for (int k = 0; k <= count; k++) t.push_back(v); }
            }
            raw = exr_unpredict(t);
        } else {  // PIZ
            std::vector<unsigned short> tmp(raw_size / 2);
            std::vector<unsigned char> bitmap(8192, 0);
            unsigned short min_nz = (unsigned short)(src[0] | (src[1] << 8)), max_nz = (unsigned short)(src[2] | (src[3] << 8));
            size_t q = 4;
            if (max_nz >= 8192) fail("EXR: PIZ bad bitmap range");
            if (min_nz <= max_nz) { memcpy(&bitmap[min_nz], src + q, (size_t)(max_nz - min_nz + 1)); q += (size_t)(max_nz - min_nz + 1); }
            std::vector<unsigned short> lut(65536, 0);
            int k = 0;
            for (int i = 0; i < 65536; i++) if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (unsigned short)i;
            unsigned short max_value = (unsigned short)(k - 1);
            int length = i32(o + 8 + q);
            q += 4;
            if (length < 0 || q + (size_t)length > (size_t)dsize) fail("EXR: PIZ bad length");
            huf_uncompress(src + q, (size_t)length, tmp.data(), tmp.size());
            size_t start = 0;
            std::vector<size_t> cstart(channels.size());
            for (size_t c = 0; c < channels.size(); c++) {
                int size = csize[c] / 2;
                cstart[c] = start;
                for (int j = 0; j < size; j++) wav2_decode(&tmp[start + j], w, size, nlines, w * size, max_value);
                start += (size_t)w * nlines * size;
            }
            for (auto &v : tmp) v = lut[v];
            raw.resize(raw_size);
            size_t out = 0;
            for (int y = 0; y < nlines; y++)
                for (size_t c = 0; c < channels.size(); c++) {
                    size_t nwords = (size_t)w * (csize[c] / 2);
                    memcpy(&raw[out], &tmp[cstart[c] + nwords * y], nwords * 2);
                    out += nwords * 2;
                }
        }
        if (raw.size() < raw_size) fail("EXR: short scanline block");
        for (int y = 0; y < nlines; y++) {
            const unsigned char *line = &raw[line_bytes * y];
            size_t coff = 0;
            for (size_t c = 0; c < channels.size(); c++) {
                for (int slot = 0; slot < 4; slot++) {
                    int want = slot == 0 ? idx_r : (slot == 1 ? idx_g : (slot == 2 ? idx_b : idx_a));
                    if (want != (int)c) continue;
                    for (int x = 0; x < w; x++) {
                        float v;
                        const unsigned char *s = line + coff + (size_t)x * csize[c];
                        if (channels[c].type == 1) v = half_to_float((unsigned short)(s[0] | (s[1] << 8)));
                        else if (channels[c].type == 2) memcpy(&v, s, 4);
                        else { uint32_t u; memcpy(&u, s, 4); v = (float)u; }
                        img.data[((size_t)(y0 + y) * w + x) * 4 + slot] = v;
                    }
                }
                coff += (size_t)csize[c] * w;
            }
        }
    }
    return img;
}

// ------------------------------------------------------------------------------------------------ PFM
ImageF read_pfm(const std::string &path) {
    std::vector<unsigned char> f = read_file(path);
    size_t p = 0;
    auto token = [&]() { std::string s; while (p < f.size() && isspace(f[p])) p++; while (p < f.size() && !isspace(f[p])) s.push_back((char)f[p++]); return s; };
    std::string magic = token();
    int ch = magic == "PF" ? 3 : (magic == "Pf" ? 1 : 0);
    if (!ch) fail("not a PFM file: " + path);
    ImageF img;
    img.width = atoi(token().c_str()); img.height = atoi(token().c_str()); img.channels = ch;
    double scale = atof(token().c_str());
    p++;
    size_t n = (size_t)img.width * img.height * ch;
    if (p + n * 4 > f.size()) fail("PFM: truncated");
    img.data.resize(n);
    memcpy(img.data.data(), &f[p], n * 4);
    if (scale > 0) for (size_t i = 0; i < n; i++) { unsigned char *b = (unsigned char *)&img.data[i]; std::swap(b[0], b[3]); std::swap(b[1], b[2]); }
    return img;
}

// ------------------------------------------------------------------------------------------------ front doors
ImageF read_image(const std::string &path, int channels) {
    if (channels != 1 && channels != 3) fail("read_image: channels must be 1 or 3");
    std::string ext = lower_ext(path);
    ImageF out;
    out.channels = channels;
    if (ext == ".exr") {
        ImageF rgba = decode_exr(read_file(path));
        out.width = rgba.width; out.height = rgba.height;
        size_t n = (size_t)rgba.width * rgba.height;
        out.data.resize(n * channels);
        for (size_t i = 0; i < n; i++) {
            const float *s = &rgba.data[4 * i];
            if (channels == 3) { out.data[3 * i] = s[0]; out.data[3 * i + 1] = s[1]; out.data[3 * i + 2] = s[2]; }
            else out.data[i] = (float)(((double)s[0] + (double)s[1] + (double)s[2]) / 3);  // image.cpp:73-75, in double
        }
        return out;
    }
    if (ext == ".hdr" || ext == ".pic") {
        ImageF rgb = decode_hdr(read_file(path));
        out.width = rgb.width; out.height = rgb.height;
        size_t n = (size_t)rgb.width * rgb.height;
        out.data.resize(n * channels);
        for (size_t i = 0; i < n; i++) {
            const float *s = &rgb.data[3 * i];
            if (channels == 3) { out.data[3 * i] = s[0]; out.data[3 * i + 1] = s[1]; out.data[3 * i + 2] = s[2]; }
            else out.data[i] = (s[0] + s[1] + s[2]) / 3;  // stb's HDR -> 1 channel conversion averages
        }
        return out;
    }
    if (ext == ".pfm") {
        ImageF src = read_pfm(path);
        out.width = src.width; out.height = src.height;
        size_t n = (size_t)src.width * src.height;
        out.data.resize(n * channels);
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < channels; c++) out.data[i * channels + c] = src.data[i * src.channels + (src.channels == 3 ? (channels == 3 ? c : 0) : 0)];
        return out;
    }
    // 8-bit formats: decode, convert the channel count the way stb does, then gamma 2.2 -> linear (stbi_loadf)
    std::vector<unsigned char> file = read_file(path), px;
    int w = 0, h = 0, comps = 0;
    if (ext == ".jpg" || ext == ".jpeg") {
        px = decode_jpeg(file, w, h, comps, channels);
    } else if (ext == ".png") {
        std::vector<unsigned char> src = decode_png(file, w, h, comps);
        px.resize((size_t)w * h * channels);
        for (size_t i = 0; i < (size_t)w * h; i++) {
            const unsigned char *s = &src[i * comps];
            if (channels == 3) {
                if (comps >= 3) { px[3 * i] = s[0]; px[3 * i + 1] = s[1]; px[3 * i + 2] = s[2]; }
                else px[3 * i] = px[3 * i + 1] = px[3 * i + 2] = s[0];
            } else {
                px[i] = comps >= 3 ? compute_y(s[0], s[1], s[2]) : s[0];
            }
        }
    } else {
        fail("Unsupported image format: " + path);
    }
    out.width = w; out.height = h;
    out.data.resize((size_t)w * h * channels);
    for (size_t i = 0; i < out.data.size(); i++) out.data[i] = (float)(pow(px[i] / 255.0f, 2.2f) * 1.0f);
    return out;
}

void write_image(const std::string &path, int width, int height, const float *rgb) {
    std::string ext = lower_ext(path);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) fail("cannot write " + path);
    if (ext == ".pfm") {
        fprintf(f, "PF\n%d %d\n-1\n", width, height);
        fwrite(rgb, sizeof(float), (size_t)width * height * 3, f);
        fclose(f);
        return;
    }
    if (ext != ".exr") { fclose(f); fail("unsupported output format (use .exr or .pfm): " + path); }
    // header
    std::vector<unsigned char> hd;
    auto put = [&](const void *p, size_t n) { hd.insert(hd.end(), (const unsigned char *)p, (const unsigned char *)p + n); };
    auto put_i32 = [&](int32_t v) { put(&v, 4); };
    auto put_str = [&](const char *s) { put(s, strlen(s) + 1); };
    auto attr = [&](const char *name, const char *type, const void *data, int size) { put_str(name); put_str(type); put_i32(size); put(data, (size_t)size); };
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    put(magic, 8);
    {
        std::vector<unsigned char> ch;
        for (const char *nm : {"B", "G", "R"}) {
            ch.push_back((unsigned char)nm[0]); ch.push_back(0);
            int32_t v[4] = {1 /*HALF*/, 0, 1, 1};
            ch.insert(ch.end(), (unsigned char *)v, (unsigned char *)v + 16);
        }
        ch.push_back(0);
        attr("channels", "chlist", ch.data(), (int)ch.size());
    }
    const bool zip = width >= 16 && height >= 16;  // (tinyexr's SaveEXR leaves tiny images uncompressed)
    unsigned char comp = zip ? 3 : 0, line_order = 0;
    attr("compression", "compression", &comp, 1);
    int32_t win[4] = {0, 0, width - 1, height - 1};
    attr("dataWindow", "box2i", win, 16);
    attr("displayWindow", "box2i", win, 16);
    attr("lineOrder", "lineOrder", &line_order, 1);
    float one = 1.0f, zero2[2] = {0, 0};
    attr("pixelAspectRatio", "float", &one, 4);
    attr("screenWindowCenter", "v2f", zero2, 8);
    attr("screenWindowWidth", "float", &one, 4);
    hd.push_back(0);
    const int lpb = zip ? 16 : 1, nblocks = (height + lpb - 1) / lpb;
    std::vector<std::vector<unsigned char>> chunks(nblocks);
    for (int b = 0; b < nblocks; b++) {
        int y0 = b * lpb, nl = std::min(lpb, height - y0);
        std::vector<unsigned char> raw((size_t)nl * width * 3 * 2);
        size_t o = 0;
        for (int y = y0; y < y0 + nl; y++)
            for (int c = 2; c >= 0; c--)  // B, G, R planes per scanline
                for (int x = 0; x < width; x++) {
                    unsigned short hv = float_to_half(rgb[((size_t)y * width + x) * 3 + c]);
                    raw[o++] = (unsigned char)(hv & 0xff); raw[o++] = (unsigned char)(hv >> 8);
                }
        std::vector<unsigned char> data;
        if (zip) {
            std::vector<unsigned char> t(raw.size());
            size_t half = (raw.size() + 1) / 2, a = 0, bb = half;
            for (size_t i = 0; i < raw.size(); i++) { if (i & 1) t[bb++] = raw[i]; else t[a++] = raw[i]; }
            unsigned char prev = t[0];
            for (size_t i = 1; i < t.size(); i++) { unsigned char cur = t[i]; t[i] = (unsigned char)(cur - prev + 128); prev = cur; }
            data = zlib_deflate(t.data(), t.size());
            if (data.size() >= raw.size()) data = raw;
        } else data = raw;
        std::vector<unsigned char> &ck = chunks[b];
        int32_t yy = y0, sz = (int32_t)data.size();
        ck.insert(ck.end(), (unsigned char *)&yy, (unsigned char *)&yy + 4);
        ck.insert(ck.end(), (unsigned char *)&sz, (unsigned char *)&sz + 4);
        ck.insert(ck.end(), data.begin(), data.end());
    }
    fwrite(hd.data(), 1, hd.size(), f);
    uint64_t off = hd.size() + (uint64_t)nblocks * 8;
    for (int b = 0; b < nblocks; b++) { fwrite(&off, 8, 1, f); off += chunks[b].size(); }
    for (int b = 0; b < nblocks; b++) fwrite(chunks[b].data(), 1, chunks[b].size(), f);
    fclose(f);
}

}  // namespace ljhost
