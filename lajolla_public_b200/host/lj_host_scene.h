// lj_host_scene.h -- the host front end's scene: what the reference's parse_scene() builds (parsers/parse_scene.cpp),
// held as the flat tables of include/lajolla_b200.h.  parse_scene_file() reads a Mitsuba-style XML file (the subset in
// SURVEY.md appendix A) and its assets; to_desc() hands the result to lj_scene_create; write_ljs() writes the same
// content as the `.ljs` container (lajolla_public_b200/ljs.py).  Host-only code: no CUDA here.
#pragma once
#include <string>
#include <vector>

#include "../../include/lajolla_b200.h"
#include "lj_math.h"
#include "lj_mesh_io.h"

namespace ljhost {

struct HostTexture {
    int kind = LJ_TEX_CONSTANT;
    int image_id = -1;     // index into the 3-channel or the 1-channel pool (see one_channel)
    bool one_channel = false;
    double value[3] = {0, 0, 0}, color1[3] = {0, 0, 0};
    double uscale = 1, vscale = 1, uoffset = 0, voffset = 0;
};
struct HostMaterial {
    int type = LJ_MAT_LAMBERTIAN;
    double eta = 1;
    HostTexture tex[LJ_NUM_TEX_SLOTS];
};
struct HostShape {
    int type = LJ_SHAPE_MESH;
    int material_id = -1, area_light_id = -1, interior_medium_id = -1, exterior_medium_id = -1;
    Vec3 center;
    double radius = 0;
    Mesh mesh;
};
struct HostLight {
    int type = LJ_LIGHT_AREA;
    int shape_id = -1;
    double intensity[3] = {0, 0, 0};
    HostTexture values;
    Mat4 to_world = Mat4::identity(), to_local = Mat4::identity();
    double scale = 1;
};
struct HostVolume {
    bool is_grid = false;
    double value[3] = {0, 0, 0};  // constant volumes
    double scale = 1;
    VolumeGrid grid;
};
struct HostMedium {
    int type = LJ_MEDIUM_HOMOGENEOUS;
    int phase_type = LJ_PHASE_ISOTROPIC;
    double phase_g = 0;
    double sigma_a[3] = {0, 0, 0}, sigma_s[3] = {0, 0, 0};
    HostVolume albedo, density;
};
struct HostImage { int width = 0, height = 0, channels = 0; std::vector<float> data; };

struct HostScene {
    // camera (camera.cpp:7-21)
    Mat4 cam_to_world = Mat4::identity(), world_to_cam = Mat4::identity(), sample_to_cam = Mat4::identity(), cam_to_sample = Mat4::identity();
    int width = 256, height = 256;
    int filter_type = LJ_FILTER_BOX;
    double filter_param = 1;
    int camera_medium_id = -1;
    // options (scene.h:24-31)
    int integrator = LJ_INT_PATH, samples_per_pixel = 4, max_depth = -1, rr_depth = 5, vol_path_version = 0, max_null_collisions = 1000;
    std::string output_filename = "image.exr";
    std::vector<HostImage> images3, images1;
    std::vector<HostMaterial> materials;
    std::vector<HostShape> shapes;
    std::vector<HostLight> lights;
    std::vector<HostMedium> media;
    int envmap_light_id = -1;
};

// parse_scene(filename) (parse_scene.cpp:1602-1617): asset paths are resolved relative to the XML file's directory.
HostScene parse_scene_file(const std::string &xml_path);

// The C-ABI description of a HostScene.  `FlatScene` owns every array the description points into.
struct FlatScene {
    lj_scene_desc desc;
    std::vector<lj_image_desc> images;
    std::vector<lj_material_desc> materials;
    std::vector<lj_shape_desc> shapes;
    std::vector<lj_light_desc> lights;
    std::vector<lj_medium_desc> media;
    std::vector<std::vector<float>> floats;
    std::vector<std::vector<int32_t>> ints;
};
void to_flat(const HostScene &scene, FlatScene &out);
void write_ljs(const FlatScene &flat, const std::string &path);

}  // namespace ljhost
