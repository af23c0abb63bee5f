// lj_scene_dev.h -- the scene as it lives in HBM: flat POD tables + pointers.
// Replaces the reference's std::vector<std::variant<...>> Scene (scene.h:39-88).
#pragma once
#include "lj_common.h"

namespace lj {

// ---- textures (texture.h:84-113, mipmap.h) -------------------------------------------------
constexpr int kMaxMipLevels = 8;  // mipmap.h:5

struct DevTexture {  // 64 B
    int kind;        // LJ_TEX_*
    int image_id;
    float v0[3];     // constant value / checker color0
    float v1[3];     // checker color1
    float uscale, vscale, uoffset, voffset;
    int _pad[4];
};

// One mip chain. 3-channel texels are stored as V4 (rgb, 0): one 16-byte load per texel.
struct DevImage {
    int levels;
    int channels;
    int w[kMaxMipLevels], h[kMaxMipLevels];
    int offset[kMaxMipLevels];  // texel offset of each level inside the pool of its channel count
};

// ---- materials (material.h:10-110) ------------------------------------------------------------
constexpr int kNumTexSlots = 12;
struct DevMaterial {
    int type;  // LJ_MAT_*
    float eta;
    int _pad[2];
    DevTexture tex[kNumTexSlots];
};

// ---- shapes (shape.h:26-54) -------------------------------------------------------------------
struct DevShape {
    int type;  // 0 sphere, 1 mesh
    int material_id, area_light_id, interior_medium_id, exterior_medium_id;
    float cx, cy, cz, radius;  // sphere
    int vertex_offset;         // into positions/normals/uvs pools
    int tri_offset;            // into indices pool (int3 per triangle, GLOBAL vertex ids)
    int num_tris;
    int has_normals, has_uvs;
    float total_area;          // triangle_mesh.inl:60-75
    int cdf_offset;            // into tri_cdf pool: num_tris+1 floats (table_dist.cpp:3-25)
};

// ---- BVH (replaces the Embree scene, scene.cpp:20-27) -----------------------------------------
// Primitive record, 48 B = 3 x 16-byte loads, stored in BVH leaf order.
//   triangle: a = (v0.xyz, v1.x)  b = (v1.yz, v2.xy)  c = (v2.z, bits(shape_id), bits(prim_id), bits(0))
//   sphere:   a = (center.xyz, r) b = unused           c = (0,   bits(shape_id), bits(0),       bits(1))
struct DevPrim { V4 a, b, c; };

// 8-wide compressed node, 80 B = 5 x 16-byte loads (Ylitie, Karras, Laine 2017 layout):
//   q0 = (origin.xyz, bits(ex | ey<<8 | ez<<16 | imask<<24))
//   q1 = bits(child_base, prim_base, meta[0..3], meta[4..7])
//   q2 = bits(qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7])
//   q3 = bits(qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7])
//   q4 = bits(qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7])
// Child box k = origin + q * 2^(e - 127) per axis.  meta[k]: 0 = empty slot; internal child = 0b001 in the
// top 3 bits and 24 + k below; leaf slot = primitive count in unary (1, 3, 7) in the top 3 bits and the index
// of its first primitive relative to prim_base below.  Internal children are stored contiguously from
// child_base in slot order (imask marks their slots), leaf primitives contiguously from prim_base.
struct DevNode8 { V4 q0, q1, q2, q3, q4; };

// ---- lights (light.h:14-27) --------------------------------------------------------------------
struct DevLight {
    int type;  // 0 area, 1 envmap
    int shape_id;
    float intensity[3];
    float scale;
    int _pad[2];
    DevTexture values;
    M44 to_world, to_local;
};

// TableDist2D of the envmap (table_dist.cpp:40-151)
struct DevTable2D {
    int width, height;
    float total_values;
    const float *cdf_rows;       // height*(width+1)
    const float *pdf_rows;       // height*width
    const float *cdf_marginals;  // height+1
    const float *pdf_marginals;  // height
};

// ---- media (medium.h, volume.h) ----------------------------------------------------------------
struct DevVolume {
    int is_grid;
    int res[3];
    float value[3];
    float p_min[3], p_max[3];
    float scale;
    float max_data[3];  // volume.h get_max_value before scale
    const V4 *data;     // rgb0 per voxel, (z*ny+y)*nx+x
    // Majorant grid of a density volume (null: none): the maximum of `data` over blocks of maj_block^3 voxel cells
    // (one node of dilation), rgb0 per block, (z*my+y)*mx+x -- local majorants for the tracking loops (lj_media.h)
    const V4 *maj;
    int maj_res[3];
    int maj_block;
};
struct DevMedium {
    int type;  // 0 homogeneous, 1 heterogeneous
    int phase_type;
    float phase_g;
    float sigma_a[3], sigma_s[3];
    DevVolume albedo, density;
};

struct DevCamera {
    M44 cam_to_world, sample_to_cam;
    int width, height;
    int filter_type;
    float filter_param;
    int medium_id;
};

struct DevOptions {
    int integrator, spp, max_depth, rr_depth, vol_path_version, max_null_collisions;
};

struct DevScene {
    DevCamera camera;
    DevOptions options;
    // geometry pools
    const float *positions;  // 3 floats per vertex
    const float *normals;    // 3 floats per vertex (valid where the mesh has_normals)
    const float *uvs;        // 2 floats per vertex
    const int *indices;      // 3 ints per triangle, global vertex ids
    const float *tri_cdf;
    const DevShape *shapes;
    int num_shapes;
    // BVH
    const DevPrim *prims;
    const DevNode8 *nodes8;  // root = node 0
    int num_prims;
    // shading tables
    const DevMaterial *materials;
    int num_materials;
    const DevImage *images1, *images3;
    const float *texels1;
    const V4 *texels3;
    const DevLight *lights;
    int num_lights;
    int envmap_light_id;
    const float *light_pmf, *light_cdf;  // scene.cpp:48-52
    DevTable2D envmap_dist;
    const DevMedium *media;
    int num_media;
    // scene.cpp:29-34, scene.h:99-105
    float bsphere_radius;
    V3 bsphere_center;
    float shadow_eps, isect_eps;
};

}  // namespace lj
