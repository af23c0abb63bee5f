// scene.cuh -- the opaque lj_scene: device allocations + the DevScene handed to every kernel.
#pragma once
#include "lj_cuda.h"

#include <string>
#include <vector>

#include "../../include/lajolla_b200.h"
#include "bvh_build.cuh"
#include "lj_path.h"

namespace lj {
struct LaunchGeom { int trace_blocks = 0, walk_blocks = 1, step_blocks = 1, flight_blocks = 1, q_blocks = 1, wtrack_blocks = 1; };
}  // namespace lj

struct lj_scene {
    lj::DevScene dev;                 // by-value kernel parameter
    std::vector<void *> allocations;  // everything cudaMalloc'ed for this scene
    lj_scene_info info;
    int device = 0;
    bool only_lambertian = false; // every material is Lambertian: k_shade without the material dispatchers
    bool has_disney = false;      // some material is a Disney lobe: k_shade calls the BSDF dispatchers out of line
    bool has_grid_media = false;  // some medium is heterogeneous: tracking loops run ~100 collisions per segment
    // host copies needed by introspection entry points
    std::vector<float> h_light_pmf, h_light_cdf;
    std::vector<lj::DevImage> h_images1, h_images3;
    // persistent path pool (allocated on first render, reused)
    lj::PathPool pool;
    void *pool_block = nullptr;
    size_t pool_bytes = 0;
    int pool_capacity = 0;
    float *d_film = nullptr;    // w*h*4 fp32: sum rgb, sample count
    float *d_film_sq = nullptr; // w*h*4 fp32: sum of squares rgb
    size_t film_bytes = 0, h_counters_bytes = 0, qstack_bytes = 0;
    // reused across renders: event creation / pinned allocation cost ~0.4 s per 1024-spp render otherwise
    std::vector<cudaEvent_t> event_pool;
    unsigned long long *d_counters = nullptr, *h_counters = nullptr;
    unsigned int *d_cursors = nullptr;
    lj::LaunchGeom geom;              // persistent grid sizes for this scene's device
    void *d_qstack = nullptr;         // scratch group stacks of k_trace_q
    int qdepth = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    std::vector<lj_scene *> replicas;  // the same scene on the other devices of lj_init (owned; empty on a replica)
    bool peer_checked = false, peer_ok = false;  // the primary device can read the replicas' films directly
};

namespace lj {
// internal Hit (BVH leaf order) -> the ABI's lj_hit (shape id, primitive id, Embree-style (u, v))
LJ_HD void hit_to_abi(const DevScene &sc, V3 org, V3 dir, const Hit &h, lj_hit &out) {
    if (h.prim == kNoHit) { out.t = h.t; out.u = 0; out.v = 0; out.shape_id = -1; out.primitive_id = -1; return; }
    V4 pc = ld4(&sc.prims[h.prim].c);
    out.t = h.t;
    out.shape_id = prim_shape_id(pc);
    out.primitive_id = prim_primitive_id(pc);
    if (prim_is_sphere(pc)) {
        V4 pa = ld4(&sc.prims[h.prim].a);
        V2 st = sphere_st((org + dir * h.t) - xyz(pa), pa.w);
        out.u = st.x; out.v = st.y;
    } else {
        out.u = h.u; out.v = h.v;
    }
}

// Makes the scene's device current on the calling thread for the lifetime of the guard (every entry point that
// takes an lj_scene* holds one: streams, events and allocations of a scene belong to its device).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// Path-pool blocks (0.6-0.9 GB) are recycled across scenes of one process: a cudaMalloc / cudaFree pair of that size
// stalls the host for up to 100 ms now and then, which a render-per-scene caller pays on every call.
void *pool_block_take(int device, size_t bytes, size_t *got_bytes);
void pool_block_give(int device, void *block, size_t bytes);
// recycled cudaMalloc (films) / cudaMallocHost (counters) buffers, exact size match; nullptr if the allocation fails
enum { kSpareDevice = 0, kSpareHost = 1 };
void *spare_take(int kind, int device, size_t bytes);
void spare_give(int kind, int device, void *ptr, size_t bytes);
void set_error(const std::string &msg);
std::vector<int> init_devices();  // the device list of the last lj_init
int cuda_fail(cudaError_t e, const char *what);
}  // namespace lj

#define LJ_CUDA(x)                                                        \
    do {                                                                  \
        cudaError_t e__ = (x);                                            \
        if (e__ != cudaSuccess) return lj::cuda_fail(e__, #x);            \
    } while (0)
