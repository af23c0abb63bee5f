// main.cpp -- the `lajolla` command of the B200 build: lajolla [-t n] [-o out] scene.xml ...
// Same surface as the reference's main.cpp:11-50 (parse -> render -> write image, the same progress lines); the render
// runs in libljb200.so on the GPU(s).  Extra switches: --spp N (override sampleCount), --gpus N / --split spp|tiles
// (multi-GPU render inside the library), --dump-ljs FILE (write the flat scene description and exit: needs no GPU).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/lajolla_b200.h"
#include "lj_host_scene.h"
#include "lj_image_io.h"

int main(int argc, char *argv[]) {
    if (argc <= 1) {
        std::cout << "[Usage] ./lajolla [-t num_threads] [-o output_file_name] [--spp n] [--gpus n] [--split spp|tiles] [--dump-ljs file.ljs] filename.xml" << std::endl;
        return 0;
    }
    std::string outputfile, dump_path, split = "auto";
    std::vector<std::string> filenames;
    int spp = 0, gpus = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "missing value after " << a << std::endl; exit(1); } return argv[++i]; };
        if (a == "-t") next();  // host thread count of the reference's tile pool: accepted, the GPU path has no use for it
        else if (a == "-o") outputfile = next();
        else if (a == "--spp") spp = atoi(next().c_str());
        else if (a == "--gpus") gpus = atoi(next().c_str());
        else if (a == "--split") split = next();
        else if (a == "--dump-ljs") dump_path = next();
        else filenames.push_back(a);
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    bool initialised = false;
    for (const std::string &filename : filenames) {
        try {
            auto t0 = now();
            std::cout << "Parsing and constructing scene " << filename << "." << std::endl;
            ljhost::HostScene scene = ljhost::parse_scene_file(filename);
            ljhost::FlatScene flat;
            ljhost::to_flat(scene, flat);
            if (!dump_path.empty()) {
                ljhost::write_ljs(flat, dump_path);
                std::cout << "Done. Took " << secs(t0, now()) << " seconds." << std::endl;
                std::cout << "Scene description written to " << dump_path << std::endl;
                continue;
            }
            if (!initialised) {
                std::vector<int> ids;
                for (int g = 0; g < std::max(gpus, 1); g++) ids.push_back(g);
                if (lj_init(ids.data(), (int)ids.size()) != LJ_OK) { std::cerr << "lj_init: " << lj_last_error() << std::endl; return 1; }
                initialised = true;
            }
            lj_scene *dev_scene = nullptr;
            if (lj_scene_create(&flat.desc, &dev_scene) != LJ_OK) { std::cerr << "lj_scene_create: " << lj_last_error() << std::endl; return 1; }
            auto t1 = now();
            std::cout << "Done. Took " << secs(t0, t1) << " seconds." << std::endl;
            std::cout << "Rendering..." << std::endl;
            lj_render_opts opts;
            memset(&opts, 0, sizeof(opts));
            opts.spp = spp;
            opts.normalize = 1;
            opts.num_gpus = gpus;
            opts.split = split == "tiles" ? LJ_SPLIT_TILES : (split == "spp" ? LJ_SPLIT_SPP : LJ_SPLIT_AUTO);
            lj_stats stats;
            std::vector<float> img((size_t)scene.width * scene.height * 3);
            int rc = lj_render(dev_scene, &opts, img.data(), &stats);
            lj_scene_destroy(dev_scene);
            if (rc != LJ_OK) { std::cerr << "lj_render: " << lj_last_error() << std::endl; return 1; }
            if (outputfile.empty()) outputfile = scene.output_filename;
            auto t2 = now();
            std::cout << "Done. Took " << secs(t1, t2) << " seconds." << std::endl;
            std::cout << "  (" << stats.samples / 1e6 << " Msamples, " << (stats.closest_rays + stats.shadow_rays) / 1e6 << " Mrays, "
                      << stats.samples / (stats.render_ms * 1e3) << " Msamples/s on " << (stats.gpus_used > 0 ? stats.gpus_used : 1) << " GPU(s))" << std::endl;
            ljhost::write_image(outputfile, scene.width, scene.height, img.data());
            std::cout << "Image written to " << outputfile << std::endl;
        } catch (std::exception &e) {
            std::cerr << "Error: " << e.what() << std::endl;
            return 1;
        }
    }
    return 0;
}
