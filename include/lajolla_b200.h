/* lajolla_b200.h -- C ABI of the B200-native rendering hot path (libljb200.so).
 *
 * The reference (BachiLi/lajolla_public) has no FFI layer; the two seams this ABI replaces are
 * (SURVEY.md 8b):
 *   S1  Image3 render(const Scene&)                      src/render.h:9, src/render.cpp:155-170
 *       fed by parse_scene()                              src/parsers/parse_scene.h:9
 *   S2  the per-ray scene queries behind it               src/intersection.h:39-51 (intersect /
 *       occluded), src/material.h:119-163 (eval / sample_bsdf / pdf_sample_bsdf),
 *       src/light.h:29-51, src/camera.h:27-28, src/texture.h:161-163, src/pcg.h:22-68
 * Everything that crosses the boundary is plain C: POD structs, pointers and sizes.  The caller
 * owns every buffer it passes in; the library copies what it needs during lj_scene_create and owns
 * everything behind lj_scene*.  No function throws or aborts; each returns LJ_OK or an error code
 * and lj_last_error() gives the text.  The library needs a CUDA device: there is no CPU path.
 *
 * All geometry/shading arithmetic on the device is IEEE fp32 (the reference uses double for
 * shading and float for the Embree ray casts, src/lajolla.h:23, src/intersection.cpp:15-24).
 */
#ifndef LAJOLLA_B200_H
#define LAJOLLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LJ_OK 0
#define LJ_ERR_INVALID 1     /* bad argument / malformed description */
#define LJ_ERR_CUDA 2        /* CUDA runtime failure (text in lj_last_error) */
#define LJ_ERR_NO_DEVICE 3   /* no usable CUDA device: the library has no CPU fallback */
#define LJ_ERR_UNSUPPORTED 4

/* ---- enums: numeric values follow the reference's std::variant alternative order ---------- */
enum { LJ_TEX_CONSTANT = 0, LJ_TEX_IMAGE = 1, LJ_TEX_CHECKERBOARD = 2 };      /* texture.h:112-113 */
enum {                                                                        /* material.h:102-110 */
    LJ_MAT_LAMBERTIAN = 0, LJ_MAT_ROUGHPLASTIC = 1, LJ_MAT_ROUGHDIELECTRIC = 2,
    LJ_MAT_DISNEY_DIFFUSE = 3, LJ_MAT_DISNEY_METAL = 4, LJ_MAT_DISNEY_GLASS = 5,
    LJ_MAT_DISNEY_CLEARCOAT = 6, LJ_MAT_DISNEY_SHEEN = 7, LJ_MAT_DISNEY_BSDF = 8
};
enum { LJ_SHAPE_SPHERE = 0, LJ_SHAPE_MESH = 1 };                              /* shape.h:54 */
enum { LJ_LIGHT_AREA = 0, LJ_LIGHT_ENVMAP = 1 };                              /* light.h:27 */
enum { LJ_MEDIUM_HOMOGENEOUS = 0, LJ_MEDIUM_HETEROGENEOUS = 1 };              /* medium.h */
enum { LJ_PHASE_ISOTROPIC = 0, LJ_PHASE_HG = 1 };                             /* phase_function.h */
enum { LJ_FILTER_BOX = 0, LJ_FILTER_TENT = 1, LJ_FILTER_GAUSSIAN = 2 };       /* filter.h */
enum {                                                                        /* scene.h:14-22 */
    LJ_INT_DEPTH = 0, LJ_INT_SHADING_NORMAL = 1, LJ_INT_MEAN_CURVATURE = 2,
    LJ_INT_RAY_DIFFERENTIAL = 3, LJ_INT_MIPMAP_LEVEL = 4, LJ_INT_PATH = 5, LJ_INT_VOLPATH = 6
};

/* Texture slots inside lj_material_desc.tex[] (names follow material.h:10-98). */
enum {
    /* Lambertian */       LJ_SLOT_REFLECTANCE = 0,
    /* RoughPlastic */     LJ_SLOT_DIFFUSE_REFLECTANCE = 0, LJ_SLOT_SPECULAR_REFLECTANCE = 1, LJ_SLOT_ROUGHNESS = 2,
    /* RoughDielectric */  /* SPECULAR_REFLECTANCE = 1, ROUGHNESS = 2, */ LJ_SLOT_SPECULAR_TRANSMITTANCE = 0,
    /* Disney*  */         LJ_SLOT_BASE_COLOR = 0, /* ROUGHNESS = 2 */ LJ_SLOT_SUBSURFACE = 1, LJ_SLOT_ANISOTROPIC = 3,
                           LJ_SLOT_CLEARCOAT_GLOSS = 4, LJ_SLOT_SHEEN_TINT = 5, LJ_SLOT_SPECULAR_TRANSMISSION = 6,
                           LJ_SLOT_METALLIC = 7, LJ_SLOT_SPECULAR = 8, LJ_SLOT_SPECULAR_TINT = 9,
                           LJ_SLOT_SHEEN = 10, LJ_SLOT_CLEARCOAT = 11,
    LJ_NUM_TEX_SLOTS = 12
};

/* ---- flat scene description (what parse_scene()+Scene::Scene produce, scene.cpp:4-53) ------ */

/* One mip-level-0 image of the TexturePool (texture.h:9-15); the library builds the mip chain
 * on the device following mipmap.h:24-48. channels is 1 (Mipmap1) or 3 (Mipmap3). */
typedef struct lj_image_desc {
    int32_t width, height, channels;
    int32_t _pad;
    const float *data; /* row-major, top row first, channels interleaved */
} lj_image_desc;

/* Texture<T> (texture.h:84-113).  A Texture<Real> uses value[0]/color1[0]. */
typedef struct lj_texture_desc {
    int32_t kind;      /* LJ_TEX_* */
    int32_t image_id;  /* index into images of matching channel count, LJ_TEX_IMAGE only */
    float value[3];    /* constant value, or checkerboard color0 */
    float color1[3];   /* checkerboard color1 */
    float uscale, vscale, uoffset, voffset;
} lj_texture_desc;

typedef struct lj_material_desc { /* material.h:10-110 */
    int32_t type; /* LJ_MAT_* */
    float eta;    /* internal IOR / external IOR where the material has one */
    lj_texture_desc tex[LJ_NUM_TEX_SLOTS];
} lj_material_desc;

typedef struct lj_shape_desc { /* shape.h:26-54 */
    int32_t type; /* LJ_SHAPE_* */
    int32_t material_id, area_light_id, interior_medium_id, exterior_medium_id;
    /* sphere */
    float center[3];
    float radius;
    /* triangle mesh: world-space, already transformed (parse_scene.cpp applies toWorld) */
    int32_t num_vertices, num_triangles;
    const float *positions;  /* 3*num_vertices */
    const int32_t *indices;  /* 3*num_triangles */
    const float *normals;    /* 3*num_vertices or NULL */
    const float *uvs;        /* 2*num_vertices or NULL */
} lj_shape_desc;

typedef struct lj_light_desc { /* light.h:14-27 */
    int32_t type;      /* LJ_LIGHT_* */
    int32_t shape_id;  /* area light */
    float intensity[3];
    /* envmap */
    lj_texture_desc values;
    float to_world[16], to_local[16]; /* row-major 4x4 */
    float scale;
} lj_light_desc;

typedef struct lj_volume_desc { /* volume.h GridVolume<Spectrum> or ConstantVolume */
    int32_t is_grid;      /* 0: constant value[3] */
    int32_t res[3];       /* nx, ny, nz */
    float value[3];
    float p_min[3], p_max[3];
    float scale;          /* volume.h:93-95 */
    const float *data;    /* 3*nx*ny*nz, index (z*ny+y)*nx+x, RGB interleaved */
} lj_volume_desc;

typedef struct lj_medium_desc { /* medium.h, media/ *.inl */
    int32_t type;        /* LJ_MEDIUM_* */
    int32_t phase_type;  /* LJ_PHASE_* */
    float phase_g;
    float sigma_a[3], sigma_s[3]; /* homogeneous */
    lj_volume_desc albedo, density; /* heterogeneous */
} lj_medium_desc;

typedef struct lj_camera_desc { /* camera.h:10-24 */
    float cam_to_world[16], world_to_cam[16];
    float sample_to_cam[16], cam_to_sample[16];
    int32_t width, height;
    int32_t filter_type;  /* LJ_FILTER_* */
    float filter_param;   /* box/tent: width; gaussian: stddev */
    int32_t medium_id;
} lj_camera_desc;

typedef struct lj_options_desc { /* scene.h:24-31 */
    int32_t integrator; /* LJ_INT_* */
    int32_t samples_per_pixel;
    int32_t max_depth;
    int32_t rr_depth;
    int32_t vol_path_version;
    int32_t max_null_collisions;
} lj_options_desc;

typedef struct lj_scene_desc {
    lj_camera_desc camera;
    lj_options_desc options;
    int32_t num_images, num_materials, num_shapes, num_lights, num_media;
    int32_t envmap_light_id; /* -1 if none */
    const lj_image_desc *images;
    const lj_material_desc *materials;
    const lj_shape_desc *shapes;
    const lj_light_desc *lights;
    const lj_medium_desc *media;
} lj_scene_desc;

/* ---- render seam S1 ------------------------------------------------------------------------ */
typedef struct lj_scene lj_scene; /* opaque, device-resident */

typedef struct lj_render_opts {
    int32_t spp;           /* <=0: use the scene's samples_per_pixel */
    int32_t sample_begin;  /* this call renders samples [sample_begin, sample_end) of every pixel; */
    int32_t sample_end;    /* both 0 => [0, spp).  PCG stream of a path = hash(pixel*spp + sample)    */
    int32_t normalize;     /* 1: divide by (sample_end-sample_begin) like render.cpp:94; 0: raw sums */
    int32_t pool_paths;    /* path slots resident in HBM (rounded up to a multiple of 256, at least 1024);
                              0: sized to the work, at most 1<<23 */
    uint64_t seed;         /* 0 => pcg.h:33 default seed */
    float *variance_out;   /* optional host w*h*3: per-pixel sample variance of the mean (NULL to skip) */
    /* Image-space share of this call (the reference's decomposition is image tiles, render.cpp:75-100): only the
     * 8x4-pixel tiles t with t % tile_stride == tile_offset are rendered, the other pixels stay 0.
     * tile_stride <= 1: the whole image. */
    int32_t tile_stride, tile_offset;
    /* lj_render only: GPUs to use (0: every device given to lj_init, 1: the scene's primary device) and how the work
     * is split among them; the per-GPU films are summed on the primary device. */
    int32_t num_gpus;
    int32_t split;         /* LJ_SPLIT_* */
    int32_t reduce;        /* LJ_REDUCE_* */
    int32_t _pad;
} lj_render_opts;
enum { LJ_SPLIT_AUTO = 0, LJ_SPLIT_SPP = 1, LJ_SPLIT_TILES = 2 };
enum { LJ_REDUCE_AUTO = 0, LJ_REDUCE_P2P = 1, LJ_REDUCE_NCCL = 2 };

typedef struct lj_stats {
    double render_ms;          /* CUDA-event time of the wavefront loop (inputs resident) */
    double extend_ms, shadow_ms, shade_ms, regen_ms; /* per-stage CUDA-event time */
    uint64_t samples;          /* camera paths finished */
    uint64_t closest_rays;     /* intersect()-equivalents */
    uint64_t shadow_rays;      /* occluded()-equivalents */
    uint64_t bounces;          /* iterations of path_tracing.h:66 executed */
    uint64_t kernel_launches;  /* kernels launched by this call */
    uint64_t waves;
    uint64_t extend_launches, shadow_launches, shade_launches, regen_launches;
    uint64_t node_steps;       /* wide-node tests executed by both traversal kernels */
    uint64_t prim_tests;       /* ray/primitive tests executed by both traversal kernels */
    uint64_t node_passes;      /* warp passes that ran a node step / a primitive step: steps / (32 * passes) is the */
    uint64_t prim_passes;      /* SIMD efficiency of the traversal kernels */
    uint64_t pool_paths;       /* path slots the call used */
    int32_t gpus_used, _pad;
    double reduce_ms;          /* multi-GPU: film reduction on the primary device */
} lj_stats;

/* Names the CUDA devices the library may use.  device_ids[0] is the PRIMARY device: it becomes
 * current on the calling thread, scenes are created on it and multi-GPU renders are reduced on it; lj_scene_create
 * replicates every scene on the other devices.  device_ids == NULL: devices 0 .. num_devices-1; num_devices <= 0: one
 * device.  LJ_ERR_NO_DEVICE if there is no CUDA device (the library has no CPU path). */
int lj_init(const int *device_ids, int num_devices);
const char *lj_last_error(void);

/* Uploads the description, builds the BVH (replaces rtcNewScene..rtcCommitScene, scene.cpp:20-27),
 * the mip chains (mipmap.h:24-48), the light / triangle / envmap tables (scene.cpp:36-52). */
int lj_scene_create(const lj_scene_desc *desc, lj_scene **out);
void lj_scene_destroy(lj_scene *scene);

/* Replaces Image3 render(const Scene&) (render.cpp:155).  out_rgb: HOST w*h*3 fp32, row-major,
 * top row first.  stats may be NULL. */
int lj_render(lj_scene *scene, const lj_render_opts *opts, float *out_rgb, lj_stats *stats);
/* Same, result left in DEVICE memory (w*h*3 fp32) on `stream` (a cudaStream_t, may be NULL):
 * the multi-GPU driver reduces these buffers with NCCL. */
int lj_render_device(lj_scene *scene, const lj_render_opts *opts, float *d_out_rgb, void *stream,
                     lj_stats *stats);

/* ---- query seam S2: batch forms of the reference's per-ray functions ----------------------- */
typedef struct lj_ray { float org[3]; float tnear; float dir[3]; float tfar; } lj_ray;   /* ray.h:9-12 */
typedef struct lj_hit { float t, u, v; int32_t shape_id, primitive_id; } lj_hit;       /* shape_id -1 = miss */

/* intersect() / occluded() (intersection.cpp:7-85), host buffers. */
int lj_trace_closest(lj_scene *scene, const lj_ray *rays, int64_t n, lj_hit *hits, double *kernel_ms);
int lj_trace_any(lj_scene *scene, const lj_ray *rays, int64_t n, uint8_t *occluded, double *kernel_ms);

/* The same two queries with a choice of traversal code (parity tests of the kernels that render):
 *   LJ_TRACE_PLAIN            one thread per ray, plain loop (what lj_trace_closest / lj_trace_any run)
 *   LJ_TRACE_WAVEFRONT        the rays are loaded into the path pool and traced by the persistent queue kernels
 *                             exactly as lj_render launches them for the path integrator (closest hit / NEE shadow)
 *   LJ_TRACE_WAVEFRONT_LANE   same, through the one-ray-per-lane persistent kernels (the volpath integrator's)
 * For the wavefront any-hit kernels rays[].tnear must equal the scene's shadow epsilon (lj_scene_info). */
enum { LJ_TRACE_PLAIN = 0, LJ_TRACE_WAVEFRONT = 1, LJ_TRACE_WAVEFRONT_LANE = 2,
       LJ_TRACE_WALK_WHOLE = 3, LJ_TRACE_WALK_STEP = 4, LJ_TRACE_WALK_STAGED = 5 /* lj_nee_walk_batch: the walk kernels, below */ };
typedef struct lj_trace_opts {
    int32_t kernel;      /* LJ_TRACE_* */
    int32_t pool_paths;  /* path-pool slots (0: sized to the batch); batches larger than the pool run in rounds */
    int32_t slot_stride; /* the rays occupy every slot_stride-th slot, the other slots hold no path (<= 1: dense) */
    int32_t walk_rounds; /* LJ_TRACE_WALK_STAGED: rounds of [segment traversal -> tracking] before the remaining walks finish
                            as whole loops (0: what lj_render uses); tests lower it to exercise that last kernel */
} lj_trace_opts;
int lj_trace_closest_ex(lj_scene *scene, const lj_ray *rays, int64_t n, const lj_trace_opts *opts, lj_hit *hits, double *kernel_ms);
int lj_trace_any_ex(lj_scene *scene, const lj_ray *rays, int64_t n, const lj_trace_opts *opts, uint8_t *occluded, double *kernel_ms);

/* The volpath integrator's next-event-estimation WALK (homework2.tex:459-510 + 771-810): from `origin` towards
 * `light_point` through index-matched surfaces, ratio tracking over every segment inside a medium, blocked by the
 * first surface with a material or by the bounce budget.  contribution_rgb[3*i..] = c * T * MIS weight / pdf of walk i
 * (0 if blocked).  opts->kernel: LJ_TRACE_PLAIN (one thread per walk, whole loops), LJ_TRACE_WALK_WHOLE / _STEP / _STAGED (the
 * persistent kernels lj_render launches for homogeneous / grid media: walks loaded into the path pool).  The random
 * numbers of walk i come from a stream that depends on (i, seed) only, so all three give bit-identical answers. */
typedef struct lj_walk_query {
    float origin[3]; int32_t medium_id;   /* medium the walk starts in, -1: none */
    float light_point[3]; uint32_t seed;
    float c[3]; float pdf_nee;            /* throughput * f * Le * G / pdf_nee of the NEE sample; its pdf */
    float pdf_dir; int32_t budget;        /* pdf of the scattering direction * G; index-matched surfaces the walk may still cross (-1: no limit) */
    int32_t _pad[2];
} lj_walk_query;
int lj_nee_walk_batch(lj_scene *scene, const lj_walk_query *q, int64_t n, const lj_trace_opts *opts, float *contribution_rgb, double *kernel_ms);

/* PathVertex as intersect() assembles it (intersection.h:15-35, intersection.cpp:37-62). */
typedef struct lj_vertex {
    float position[3], geometric_normal[3];
    float frame_x[3], frame_y[3], frame_n[3];
    float st[2], uv[2];
    float uv_screen_size, mean_curvature, ray_radius;
    int32_t shape_id, primitive_id, material_id, interior_medium_id, exterior_medium_id;
} lj_vertex;
/* intersect() with ray differentials (radius, spread per ray; may be NULL = {0,0}). */
int lj_intersect(lj_scene *scene, const lj_ray *rays, const float *ray_diff_radius_spread, int64_t n,
                 lj_vertex *vertices);

/* eval / pdf_sample_bsdf / sample_bsdf (material.h:119-163) on caller-supplied vertices.
 * dirs are world space, pointing away from the surface.  transport: 0 TO_LIGHT, 1 TO_VIEW. */
typedef struct lj_bsdf_query {
    lj_vertex vertex;
    float dir_in[3], dir_out[3];
    float rnd_uv[2], rnd_w;
    int32_t transport;
} lj_bsdf_query;
typedef struct lj_bsdf_result {
    float f[3];        /* eval(): BSDF * |n.dir_out| */
    float pdf;         /* pdf_sample_bsdf(dir_in, dir_out) */
    int32_t sampled;   /* sample_bsdf() succeeded */
    float s_dir_out[3], s_eta, s_roughness; /* BSDFSampleRecord */
} lj_bsdf_result;
int lj_bsdf_batch(lj_scene *scene, const lj_bsdf_query *q, int64_t n, lj_bsdf_result *out);

/* sample_light + sample_point_on_light + pdf_point_on_light + light_pmf + emission
 * (scene.cpp:61-67, light.cpp, lights/ *.inl). */
typedef struct lj_light_query { float ref_point[3]; float rnd_uv[2], rnd_w, light_w; } lj_light_query;
typedef struct lj_light_result {
    int32_t light_id;
    float position[3], normal[3];
    float pmf, pdf;
    float emission[3]; /* emission(light, -dir_light, 0, point) as path_tracing.h:174 */
} lj_light_result;
int lj_light_batch(lj_scene *scene, const lj_light_query *q, int64_t n, lj_light_result *out);

/* Medium and phase-function queries (medium.h:25-31, phase_function.h:18-29) for medium `medium_id`:
 *   majorant       = get_majorant(medium, Ray{org, dir, 0, tfar})          medium.cpp:27-29
 *   sigma_a/_s     = get_sigma_a / get_sigma_s(medium, org + t * dir)      medium.cpp:31-37, volume.h:45-81
 *   phase_dir      = sample_phase_function(phase, -dir, rnd)               phase_functions/ *.inl
 *   phase_eval/pdf = eval / pdf_sample_phase(phase, -dir, phase_dir) */
typedef struct lj_medium_query { float org[3], tfar, dir[3], t, rnd[2]; int32_t medium_id, _pad; } lj_medium_query;
typedef struct lj_medium_result { float majorant[3], sigma_a[3], sigma_s[3], phase_dir[3], phase_eval, phase_pdf; } lj_medium_result;
int lj_medium_batch(lj_scene *scene, const lj_medium_query *q, int64_t n, lj_medium_result *out);

/* The bound the tracking loops actually use at the point org + t * dir of medium `medium_id` (lj_media.h): a grid medium is
 * bounded block by block by its majorant grid where the reference uses the global maximum of get_majorant (medium.cpp:27-29).
 *   majorant = the block's bound (common to the three channels), t_exit = the ray parameter at which it stops holding,
 *   sigma_t  = sigma_a + sigma_s at the point (what the bound must dominate), local = 0 if the medium has no majorant grid
 * (then majorant is the global one and t_exit is infinite). */
typedef struct lj_medium_bound { float majorant[3], t_exit, sigma_t[3]; int32_t local; } lj_medium_bound;
int lj_medium_bound_batch(lj_scene *scene, const lj_medium_query *q, int64_t n, lj_medium_bound *out);

/* sample_primary (camera.cpp:23-47): screen_pos (x,y in [0,1]^2) -> ray. */
int lj_camera_rays(lj_scene *scene, const float *screen_pos_xy, int64_t n, lj_ray *rays);

/* eval(Texture) (texture.h:117-163) for material `material_id`, slot `slot`. */
int lj_texture_batch(lj_scene *scene, int32_t material_id, int32_t slot, const float *uv_footprint /*3 per query*/,
                     int64_t n, float *out_rgb);

/* PCG32 (pcg.h:22-68): n_streams x n_draws raw uint32 outputs of init_pcg32(first_stream+i, seed);
 * and the fp32 uniforms the device path derives from them. */
int lj_pcg32_batch(uint64_t first_stream, uint64_t seed, int32_t n_streams, int32_t n_draws,
                   uint32_t *out_u32, float *out_f32);

/* ---- introspection -------------------------------------------------------------------------- */
typedef struct lj_scene_info {
    int32_t num_prims, num_triangles, num_spheres, num_bvh_nodes, bvh_width, bvh_depth;
    float bounds_lo[3], bounds_hi[3];
    float bsphere_radius, bsphere_center[3]; /* scene.cpp:29-34 */
    float shadow_epsilon;                    /* scene.h:99-105 */
    double bvh_build_ms, upload_ms, prep_ms;
    double sah_cost;
    int64_t device_bytes;
    int32_t num_prim_refs; /* BVH leaf entries: > num_prims when large triangles were split spatially */
    int32_t _pad;
} lj_scene_info;
int lj_scene_get_info(lj_scene *scene, lj_scene_info *info);
/* Light table as built on the device (scene.cpp:48-52): pmf[num_lights], cdf[num_lights+1]. */
int lj_scene_get_light_table(lj_scene *scene, float *pmf, float *cdf);
/* One mip level of image `image_id` with `channels` in {1,3} (mipmap.h): returns dims; data may be NULL. */
int lj_scene_get_mip_level(lj_scene *scene, int32_t channels, int32_t image_id, int32_t level,
                           int32_t *width, int32_t *height, float *data);

/* Roofline denominators measured on the spot: GB/s of a read-only stream of 16-byte loads, all SMs, over a working
 * set of `bytes` swept `iters` times (after a warm-up sweep).  A set that fits the 126 MB L2 gives the L2 read rate
 * the BVH / primitive / texture fetches are bounded by; a much larger one gives HBM's. */
int lj_measure_read_bandwidth(int64_t bytes, int32_t iters, double *gb_per_s);

#ifdef __cplusplus
}
#endif
#endif /* LAJOLLA_B200_H */
