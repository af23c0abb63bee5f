// image_io_selftest.cpp -- TEST HELPER (tests/test_frontend.py): `read <image> <channels> <out.bin>` decodes an image
// file with the front end's readers into a raw dump (int32 w, h, c + floats); `write <in.bin> <image.exr|.pfm>` writes
// a raw RGB dump with the front end's writers.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <vector>

#include "lj_image_io.h"

int main(int argc, char **argv) {
    try {
        if (argc == 5 && !strcmp(argv[1], "read")) {
            ljhost::ImageF im = ljhost::read_image(argv[2], atoi(argv[3]));
            FILE *f = fopen(argv[4], "wb");
            if (!f) return 2;
            int hdr[3] = {im.width, im.height, im.channels};
            fwrite(hdr, 4, 3, f);
            fwrite(im.data.data(), 4, im.data.size(), f);
            fclose(f);
            return 0;
        }
        if (argc == 4 && !strcmp(argv[1], "write")) {
            FILE *f = fopen(argv[2], "rb");
            if (!f) return 2;
            int hdr[3];
            if (fread(hdr, 4, 3, f) != 3 || hdr[2] != 3) return 2;
            std::vector<float> px((size_t)hdr[0] * hdr[1] * 3);
            if (fread(px.data(), 4, px.size(), f) != px.size()) return 2;
            fclose(f);
            ljhost::write_image(argv[3], hdr[0], hdr[1], px.data());
            return 0;
        }
        fprintf(stderr, "usage: image_io_selftest read <image> <channels> <out.bin> | write <in.bin> <image>\n");
        return 2;
    } catch (std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
