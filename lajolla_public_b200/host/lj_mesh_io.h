// lj_mesh_io.h -- triangle meshes and volume grids of the host front end: Wavefront OBJ, Mitsuba .serialized
// (zlib-compressed), PLY (ascii / binary little endian), Mitsuba .vol grids.  Stands where the reference has
// parsers/parse_obj.cpp, load_serialized.cpp, parse_ply.cpp (tinyply), volume.cpp:16-64 and shape_utils.h.
// Everything is kept in double until the flat description is written, like the reference's TriangleMesh (shape.h:41-52).
#pragma once
#include <string>
#include <vector>

#include "lj_math.h"

namespace ljhost {

struct Mesh {
    std::vector<Vec3> positions, normals;
    std::vector<Vec2> uvs;
    std::vector<int> indices;  // 3 per triangle
};

Mesh load_obj(const std::string &path, const Mat4 &to_world);                            // parse_obj.cpp:93-185
Mesh load_serialized(const std::string &path, int shape_index, const Mat4 &to_world);    // load_serialized.cpp:174-256
Mesh load_ply(const std::string &path, const Mat4 &to_world);                            // parse_ply.cpp:9-123
std::vector<Vec3> compute_vertex_normals(const std::vector<Vec3> &positions, const std::vector<int> &indices);  // shape_utils.h:15-50

struct VolumeGrid {  // volume.cpp:16-110 load_volume(filename, 3): always three channels in memory
    int res[3] = {0, 0, 0};
    float p_min[3], p_max[3];
    std::vector<float> data;  // 3 * nx * ny * nz, x fastest
    float max_data[3] = {0, 0, 0};
};
VolumeGrid load_volume(const std::string &path);

}  // namespace ljhost
