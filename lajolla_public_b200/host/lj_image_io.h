// lj_image_io.h -- image files of the host front end: textures / environment maps in (JPEG, PNG, Radiance HDR,
// OpenEXR, PFM) and rendered images out (OpenEXR fp16, PFM).  Stands where the reference uses stb_image and tinyexr
// (image.cpp:27-173).  Conventions kept because the renderer's results depend on them:
// 8-bit images are linearised with gamma 2.2 exactly like stbi_loadf (image.cpp:43, stb's l2h gamma), a 1-channel
// read of an EXR averages R, G, B (image.cpp:73-75), EXR output is half precision (image.cpp:160-163).
#pragma once
#include <string>
#include <vector>

namespace ljhost {

struct ImageF {
    int width = 0, height = 0, channels = 0;  // channels interleaved, row-major, top row first
    std::vector<float> data;
};

// imread1 / imread3 (image.cpp:27-133): `channels` is 1 or 3.  Throws std::runtime_error.
ImageF read_image(const std::string &path, int channels);

// imwrite (image.cpp:135-173): ".pfm" -> little-endian PFM, bottom row... (the reference writes rows top-down with a
// negative scale, kept); ".exr" -> scanline OpenEXR, channels B G R as HALF, ZIP compressed.
void write_image(const std::string &path, int width, int height, const float *rgb);

// raw decoders, exposed for the tests
std::vector<unsigned char> decode_jpeg(const std::vector<unsigned char> &file, int &w, int &h, int &comps, int want_comps);
std::vector<unsigned char> decode_png(const std::vector<unsigned char> &file, int &w, int &h, int &comps);
ImageF decode_exr(const std::vector<unsigned char> &file);   // RGBA float (tinyexr LoadEXR layout)
ImageF decode_hdr(const std::vector<unsigned char> &file);   // RGB float
ImageF read_pfm(const std::string &path);

unsigned short float_to_half(float f);
float half_to_float(unsigned short h);

}  // namespace ljhost
