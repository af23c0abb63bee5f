// scene.cu -- lj_scene_create / destroy / introspection: flat description -> device tables.
// Does what the reference's Scene::Scene does at scene.cpp:4-53 (commit geometry, scene bounds,
// shape + light sampling tables, light power table) and make_mipmap (mipmap.h:24-48), for HBM.
#include "scene.cuh"

#include <mutex>
#include <thread>

#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>

namespace lj {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int cuda_fail(cudaError_t e, const char *what) {
    g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    cudaGetLastError();
    return LJ_ERR_CUDA;
}

namespace {

// 2x2 box filter, mipmap.h:32-44.  One thread per destination texel.
__global__ void k_mip_down3(const V4 *src, int sw, int sh, V4 *dst, int dw, int dh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dw * dh) return;
    int x = i % dw, y = i / dw;
    // A 1-texel-wide/high level makes upstream read past the image (mipmap.h:38-42 with
    // prev.width == 1); the edge texel is replicated here instead.
    int x1 = min(2 * x + 1, sw - 1), y1 = min(2 * y + 1, sh - 1);
    V4 a = src[(2 * y) * sw + 2 * x], b = src[(2 * y) * sw + x1];
    V4 c = src[y1 * sw + 2 * x], d = src[y1 * sw + x1];
    dst[i] = mk4((a.x + b.x + c.x + d.x) / 4, (a.y + b.y + c.y + d.y) / 4, (a.z + b.z + c.z + d.z) / 4, 0.f);
}
__global__ void k_mip_down1(const float *src, int sw, int sh, float *dst, int dw, int dh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dw * dh) return;
    int x = i % dw, y = i / dw;
    int x1 = min(2 * x + 1, sw - 1), y1 = min(2 * y + 1, sh - 1);
    dst[i] = (src[(2 * y) * sw + 2 * x] + src[(2 * y) * sw + x1] + src[y1 * sw + 2 * x] + src[y1 * sw + x1]) / 4;
}

struct Uploader {
    lj_scene *s;
    cudaError_t err = cudaSuccess;
    template <typename T>
    T *alloc(size_t n) {
        void *p = nullptr;
        size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
        cudaError_t e = lj_dev_alloc(&p, bytes);
        if (e != cudaSuccess) { if (err == cudaSuccess) err = e; return nullptr; }
        s->allocations.push_back(p);
        s->info.device_bytes += (int64_t)bytes;
        return (T *)p;
    }
    template <typename T>
    T *upload(const std::vector<T> &v) {
        T *p = alloc<T>(v.size());
        if (p && !v.empty()) {
            cudaError_t e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
            if (e != cudaSuccess && err == cudaSuccess) err = e;
        }
        return p;
    }
};

DevTexture conv_texture(const lj_texture_desc &t) {
    DevTexture d;
    memset(&d, 0, sizeof(d));
    d.kind = t.kind;
    d.image_id = t.image_id;
    for (int c = 0; c < 3; c++) { d.v0[c] = t.value[c]; d.v1[c] = t.color1[c]; }
    d.uscale = t.uscale; d.vscale = t.vscale; d.uoffset = t.uoffset; d.voffset = t.voffset;
    return d;
}

double lum(const float *c) { return c[0] * 0.212671 + c[1] * 0.715160 + c[2] * 0.072169; }

// Block size of the majorant grid of heterogeneous media in voxel cells (LJ_MAJ_BLOCK; 0: no grid, the tracking loops
// use the reference's one global majorant, medium.cpp:27-29).
int majorant_block() {  // (read at every lj_scene_create: the parity tests build one scene of each kind)
    const char *e = getenv("LJ_MAJ_BLOCK");
    return e ? std::min(64, std::max(0, atoi(e))) : 8;
}

DevVolume conv_volume(const lj_volume_desc &v, Uploader &up, bool want_majorant = false) {
    DevVolume d;
    memset(&d, 0, sizeof(d));
    d.is_grid = v.is_grid;
    for (int c = 0; c < 3; c++) {
        d.res[c] = v.res[c]; d.value[c] = v.value[c]; d.p_min[c] = v.p_min[c]; d.p_max[c] = v.p_max[c];
        d.max_data[c] = v.value[c];
    }
    d.scale = v.scale;
    if (v.is_grid) {
        size_t n = (size_t)v.res[0] * v.res[1] * v.res[2];
        std::vector<V4> tex(n);
        float mx[3] = {0, 0, 0};  // volume.h get_max_value starts from the first voxel; data are >= 0 in practice
        for (size_t i = 0; i < n; i++) {
            tex[i] = mk4(v.data[3 * i], v.data[3 * i + 1], v.data[3 * i + 2], 0.f);
            for (int c = 0; c < 3; c++) mx[c] = i == 0 ? v.data[c] : std::max(mx[c], v.data[3 * i + c]);
        }
        for (int c = 0; c < 3; c++) d.max_data[c] = mx[c];
        d.data = up.upload(tex);
        const int B = majorant_block();
        if (want_majorant && B > 0) {
            // block (bx, by, bz) covers the voxel cells [b*B, (b+1)*B) per axis, i.e. the nodes b*B .. (b+1)*B of the
            // trilinear lookup; one more node on either side keeps the bound valid for a point that sits a rounding
            // error across the block face, and the factor covers the rounding of the interpolation weights
            int mr[3];
            for (int c = 0; c < 3; c++) mr[c] = std::max(1, (std::max(v.res[c] - 1, 1) + B - 1) / B);
            std::vector<V4> maj((size_t)mr[0] * mr[1] * mr[2]);
            for (int bz = 0; bz < mr[2]; bz++)
                for (int by = 0; by < mr[1]; by++)
                    for (int bx = 0; bx < mr[0]; bx++) {
                        float m3[3] = {0, 0, 0};
                        const int x0 = std::max(bx * B - 1, 0), x1 = std::min((bx + 1) * B + 1, v.res[0] - 1);
                        const int y0 = std::max(by * B - 1, 0), y1 = std::min((by + 1) * B + 1, v.res[1] - 1);
                        const int z0 = std::max(bz * B - 1, 0), z1 = std::min((bz + 1) * B + 1, v.res[2] - 1);
                        for (int z = z0; z <= z1; z++)
                            for (int y = y0; y <= y1; y++)
                                for (int x = x0; x <= x1; x++) {
                                    const float *px = v.data + 3 * (((size_t)z * v.res[1] + y) * v.res[0] + x);
                                    for (int c = 0; c < 3; c++) m3[c] = std::max(m3[c], px[c]);
                                }
                        maj[((size_t)bz * mr[1] + by) * mr[0] + bx] = mk4(m3[0] * 1.000002f, m3[1] * 1.000002f, m3[2] * 1.000002f, 0.f);
                    }
            d.maj = up.upload(maj);
            for (int c = 0; c < 3; c++) d.maj_res[c] = mr[c];
            d.maj_block = B;
        }
    }
    return d;
}


// ---- spatial reference splitting (host, before the GPU build) ------------------------------------------------------
// Large triangles (sponza's floors, walls, arches) make every BVH over whole-triangle boxes overlap badly.  Like the
// pre-splitting of Ernst & Greiner 2007 / Karras & Aila 2013, a triangle whose box is large is cut at the midpoint of
// its box's longest axis, recursively; each piece becomes a separate REFERENCE to the same triangle with the tight box
// of the clipped polygon.  The size threshold is found by bisection so that the reference count stays within a budget
// (LJ_SPLIT_BUDGET, percent of extra references).  Traversal is unchanged: a reference is tested as the whole triangle,
// and a triangle found through two references is the same (shape, primitive) either way -- renders are bit-identical
// with and without.  OFF by default: measured on sponza (profiles/r02_bvh_quality.txt) a +30 % budget cuts primitive
// tests per ray from 7.2 to 4.8 but adds wide-node steps (11.9 -> 14.0), and a node step costs more than a primitive
// test in the L1-bound traversal kernels, so the trade is a wash; it is kept as a build option for scenes dominated
// by large diagonal triangles.
struct SplitPoly { int n; double v[10][3]; };

double box_half_area_d(const double *lo, const double *hi) {
    double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}
void poly_box(const SplitPoly &p, double *lo, double *hi) {
    for (int c = 0; c < 3; c++) { lo[c] = 1e300; hi[c] = -1e300; }
    for (int i = 0; i < p.n; i++)
        for (int c = 0; c < 3; c++) { lo[c] = std::min(lo[c], p.v[i][c]); hi[c] = std::max(hi[c], p.v[i][c]); }
}
// the part of a convex polygon with coordinate `axis` <= pos (side 0) or >= pos (side 1)
SplitPoly poly_clip(const SplitPoly &p, int axis, double pos, int side) {
    SplitPoly out;
    out.n = 0;
    for (int i = 0; i < p.n; i++) {
        const double *a = p.v[i], *b = p.v[(i + 1) % p.n];
        double da = side ? pos - a[axis] : a[axis] - pos, db = side ? pos - b[axis] : b[axis] - pos;  // <= 0: inside
        if (da <= 0 && out.n < 10) { for (int c = 0; c < 3; c++) out.v[out.n][c] = a[c]; out.n++; }
        if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
            double t = da / (da - db);
            if (out.n < 10) {
                for (int c = 0; c < 3; c++) out.v[out.n][c] = a[c] + t * (b[c] - a[c]);
                out.v[out.n][axis] = pos;
                out.n++;
            }
        }
    }
    return out;
}
// recursion: emits (or only counts) the references of one triangle for size threshold T.  A piece is cut only where
// that pays: its box is larger than T AND the two halves' boxes together are clearly smaller than the piece's own (a
// diagonal triangle loses the empty part of its box; an axis-aligned floor quad would only gain interior nodes).
// `limit`: counting stops once more than this many references have been produced.
constexpr double kSplitGain = 0.85;
size_t split_recurse(const SplitPoly &poly, const double *lo, const double *hi, double T, int depth, size_t limit, std::vector<float> *boxes_out) {
    double ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    int axis = ext[0] >= ext[1] ? (ext[0] >= ext[2] ? 0 : 2) : (ext[1] >= ext[2] ? 1 : 2);
    const double area = box_half_area_d(lo, hi);
    if (depth < 12 && limit > 1 && area > T && ext[axis] > 0) {
        double pos = 0.5 * (lo[axis] + hi[axis]);
        SplitPoly a = poly_clip(poly, axis, pos, 0), b = poly_clip(poly, axis, pos, 1);
        if (a.n >= 3 && b.n >= 3) {
            double alo[3], ahi[3], blo[3], bhi[3];
            poly_box(a, alo, ahi);
            poly_box(b, blo, bhi);
            for (int c = 0; c < 3; c++) {  // pieces stay inside the parent's box
                alo[c] = std::max(alo[c], lo[c]); ahi[c] = std::min(ahi[c], hi[c]);
                blo[c] = std::max(blo[c], lo[c]); bhi[c] = std::min(bhi[c], hi[c]);
            }
            if (box_half_area_d(alo, ahi) + box_half_area_d(blo, bhi) <= kSplitGain * area) {
                size_t na = split_recurse(a, alo, ahi, T, depth + 1, limit, boxes_out);
                if (na >= limit && !boxes_out) return na;
                return na + split_recurse(b, blo, bhi, T, depth + 1, boxes_out ? limit : limit - na, boxes_out);
            }
        }
    }
    if (boxes_out) {
        for (int c = 0; c < 3; c++) boxes_out->push_back(nextafterf((float)lo[c], -INFINITY));
        for (int c = 0; c < 3; c++) boxes_out->push_back(nextafterf((float)hi[c], INFINITY));
    }
    return 1;
}

// Rewrites prim_shape / prim_local into reference lists and fills ref_box (6 floats per reference).  Returns false
// (lists untouched, ref_box empty) when splitting is off or would not split anything.
bool split_references(const std::vector<DevShape> &shapes, const std::vector<float> &positions, const std::vector<int> &indices,
                      std::vector<int> &prim_shape, std::vector<int> &prim_local, std::vector<float> &ref_box) {
    int budget_pct = 0;
    if (const char *e = getenv("LJ_SPLIT_BUDGET")) budget_pct = atoi(e);
    const size_t n = prim_shape.size();
    if (budget_pct <= 0 || n < 64) return false;
    // polygons + boxes of the triangles; scene box
    std::vector<SplitPoly> polys(n);
    std::vector<double> lo(3 * n), hi(3 * n);
    double slo[3] = {1e300, 1e300, 1e300}, shi[3] = {-1e300, -1e300, -1e300};
    size_t n_tris = 0;
    for (size_t i = 0; i < n; i++) {
        const DevShape &sh = shapes[prim_shape[i]];
        SplitPoly &p = polys[i];
        if (sh.type == 0) {
            p.n = 0;
            const float c[3] = {sh.cx, sh.cy, sh.cz};
            for (int k = 0; k < 3; k++) { lo[3 * i + k] = c[k] - sh.radius; hi[3 * i + k] = c[k] + sh.radius; }
        } else {
            p.n = 3;
            const int *idx = &indices[3 * ((size_t)sh.tri_offset + prim_local[i])];
            for (int v = 0; v < 3; v++)
                for (int k = 0; k < 3; k++) p.v[v][k] = positions[3 * (size_t)idx[v] + k];
            poly_box(p, &lo[3 * i], &hi[3 * i]);
            n_tris++;
        }
        for (int k = 0; k < 3; k++) { slo[k] = std::min(slo[k], lo[3 * i + k]); shi[k] = std::max(shi[k], hi[3 * i + k]); }
    }
    const double scene_area = box_half_area_d(slo, shi);
    if (!(scene_area > 0) || n_tris == 0) return false;
    const size_t max_refs = n + n_tris * (size_t)budget_pct / 100;
    auto count = [&](double T) {  // (stops counting once the budget is exceeded)
        size_t total = 0;
        for (size_t i = 0; i < n && total <= max_refs; i++)
            total += polys[i].n ? split_recurse(polys[i], &lo[3 * i], &hi[3 * i], T, 0, max_refs + 1 - total, nullptr) : 1;
        return total;
    };
    double t_hi = scene_area, t_lo = scene_area * 1e-9;  // count(t_hi) == n; bisect (log scale) for the smallest T within budget
    if (count(t_lo) <= max_refs) t_hi = t_lo;
    else
        for (int it = 0; it < 24; it++) {
            double mid = sqrt(t_lo * t_hi);
            if (count(mid) <= max_refs) t_hi = mid; else t_lo = mid;
        }
    const double T = t_hi;
    if (count(T) <= n) return false;
    std::vector<int> shape2, local2;
    ref_box.clear();
    for (size_t i = 0; i < n; i++) {
        size_t before = ref_box.size() / 6;
        if (polys[i].n) split_recurse(polys[i], &lo[3 * i], &hi[3 * i], T, 0, (size_t)1 << 20, &ref_box);
        else {
            for (int k = 0; k < 3; k++) ref_box.push_back((float)lo[3 * i + k]);
            for (int k = 0; k < 3; k++) ref_box.push_back((float)hi[3 * i + k]);
        }
        size_t made = ref_box.size() / 6 - before;
        if (polys[i].n) {
            // never outside the triangle's own fp32 box: the union of the reference boxes is then exactly the box Embree
            // would report for the primitive (scene bounds, epsilons)
            for (size_t r = before; r < before + made; r++)
                for (int k = 0; k < 3; k++) {
                    ref_box[6 * r + k] = std::max(ref_box[6 * r + k], (float)lo[3 * i + k]);
                    ref_box[6 * r + 3 + k] = std::min(ref_box[6 * r + 3 + k], (float)hi[3 * i + k]);
                }
        }
        for (size_t r = 0; r < made; r++) { shape2.push_back(prim_shape[i]); local2.push_back(prim_local[i]); }
    }
    prim_shape.swap(shape2);
    prim_local.swap(local2);
    return true;
}

}  // namespace
}  // namespace lj

using namespace lj;

namespace lj {
namespace {
std::mutex g_pool_mutex;
struct CachedBlock { int device; void *block; size_t bytes; };
CachedBlock g_cached = {-1, nullptr, 0};  // at most one spare block per process
}  // namespace
void *pool_block_take(int device, size_t bytes, size_t *got_bytes) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (g_cached.block && g_cached.device == device && g_cached.bytes >= bytes && g_cached.bytes <= 2 * bytes) {
            void *b = g_cached.block;
            *got_bytes = g_cached.bytes;
            g_cached.block = nullptr;
            return b;
        }
    }
    void *b = nullptr;
    DeviceGuard guard(device);
    if (cudaMalloc(&b, bytes) != cudaSuccess) return nullptr;
    *got_bytes = bytes;
    return b;
}
void pool_block_give(int device, void *block, size_t bytes) {
    void *old = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        old = g_cached.block;
        g_cached = {device, block, bytes};
    }
    if (old) cudaFree(old);  // (cudaFree works on a pointer of any device)
}

// The small per-scene render buffers that must stay plain cudaMalloc / cudaMallocHost memory (the films are read by peer
// GPUs and by NCCL, the counters are pinned) are recycled the same way: on a shared node a cudaFree / cudaFreeHost
// sometimes blocks for 0.1 - 1 s (measured: lj_scene_destroy 2 .. 950 ms, profiles/r02y_diag_e2e.txt), which a
// scene-per-render caller paid on every call.  A handful of spares per process; exact size match.
namespace {
struct SpareBuffer { int kind, device; void *ptr; size_t bytes; };
std::vector<SpareBuffer> g_spares;
constexpr size_t kMaxSpares = 8;
void spare_release(const SpareBuffer &b) { if (b.kind == kSpareHost) cudaFreeHost(b.ptr); else cudaFree(b.ptr); }
}  // namespace
void *spare_take(int kind, int device, size_t bytes) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t i = 0; i < g_spares.size(); i++)
            if (g_spares[i].kind == kind && g_spares[i].device == device && g_spares[i].bytes == bytes) {
                void *p = g_spares[i].ptr;
                g_spares.erase(g_spares.begin() + (long)i);
                return p;
            }
    }
    void *p = nullptr;
    DeviceGuard guard(device);
    cudaError_t e = kind == kSpareHost ? cudaMallocHost(&p, bytes) : cudaMalloc(&p, bytes);
    return e == cudaSuccess ? p : nullptr;
}
void spare_give(int kind, int device, void *ptr, size_t bytes) {
    if (!ptr) return;
    SpareBuffer old = {0, 0, nullptr, 0};
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (g_spares.size() >= kMaxSpares) { old = g_spares.front(); g_spares.erase(g_spares.begin()); }
        g_spares.push_back({kind, device, ptr, bytes});
    }
    if (old.ptr) spare_release(old);
}
}  // namespace lj

extern "C" const char *lj_last_error(void) { return g_error.c_str(); }

namespace lj {
// devices handed to lj_init, in order; [0] is the primary device (scenes are created there, multi-GPU renders are
// reduced there).  Process-wide, set once per lj_init call.
static std::mutex g_devices_mutex;
static std::vector<int> g_devices;
std::vector<int> init_devices() {
    std::lock_guard<std::mutex> lock(g_devices_mutex);
    return g_devices;
}
}  // namespace lj

extern "C" int lj_init(const int *device_ids, int num_devices) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: libljb200 has no CPU path");
        return LJ_ERR_NO_DEVICE;
    }
    std::vector<int> ids;
    if (num_devices <= 0) ids.push_back(device_ids ? device_ids[0] : 0);
    else for (int i = 0; i < num_devices; i++) ids.push_back(device_ids ? device_ids[i] : i);
    for (size_t i = 0; i < ids.size(); i++) {
        if (ids[i] < 0 || ids[i] >= count) { set_error("device index out of range"); return LJ_ERR_INVALID; }
        for (size_t j = 0; j < i; j++) if (ids[j] == ids[i]) { set_error("duplicate device index"); return LJ_ERR_INVALID; }
    }
    for (size_t i = ids.size(); i-- > 0;) {  // (the primary device last, so that it stays current)
        LJ_CUDA(cudaSetDevice(ids[i]));
        LJ_CUDA(cudaFree(0));
#if !defined(LJ_HOSTSIM)
        {   // keep freed blocks in the stream-ordered pool (see lj_dev_alloc)
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, ids[i]) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
        }
#endif
    }
    {
        std::lock_guard<std::mutex> lock(g_devices_mutex);
        g_devices = ids;
    }
    return LJ_OK;
}

extern "C" void lj_scene_destroy(lj_scene *s) {
    if (!s) return;
    for (lj_scene *r : s->replicas) lj_scene_destroy(r);
    s->replicas.clear();
    DeviceGuard guard(s->device);
    cudaDeviceSynchronize();  // the caller's streams may still read the tables
    for (void *p : s->allocations) lj_dev_free(p);
    if (s->pool_block) pool_block_give(s->device, s->pool_block, s->pool_bytes);
    spare_give(kSpareDevice, s->device, s->d_film, s->film_bytes);
    spare_give(kSpareDevice, s->device, s->d_film_sq, s->film_bytes);
    for (cudaEvent_t e : s->event_pool) cudaEventDestroy(e);
    lj_dev_free(s->d_counters);
    lj_dev_free(s->d_cursors);
    spare_give(kSpareDevice, s->device, s->d_qstack, s->qstack_bytes);
    spare_give(kSpareHost, s->device, s->h_counters, s->h_counters_bytes);
    for (auto &e : s->ev) if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// builds the scene on the CURRENT device of the calling thread
static int scene_create_here(const lj_scene_desc *desc, lj_scene **out) {
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: libljb200 has no CPU path");
        return LJ_ERR_NO_DEVICE;
    }
    if (desc->num_shapes <= 0) { set_error("scene has no shapes"); return LJ_ERR_INVALID; }
    auto t_start = std::chrono::steady_clock::now();
    lj_scene *s = new lj_scene();
    memset(&s->info, 0, sizeof(s->info));
    memset(&s->dev, 0, sizeof(s->dev));
    memset(&s->pool, 0, sizeof(s->pool));
    cudaGetDevice(&s->device);
    Uploader up{s};
    DevScene &sc = s->dev;
    auto fail = [&](int code, const std::string &msg) { set_error(msg); lj_scene_destroy(s); return code; };

    // ---- camera / options
    memcpy(sc.camera.cam_to_world.m, desc->camera.cam_to_world, 64);
    memcpy(sc.camera.sample_to_cam.m, desc->camera.sample_to_cam, 64);
    sc.camera.width = desc->camera.width;
    sc.camera.height = desc->camera.height;
    sc.camera.filter_type = desc->camera.filter_type;
    sc.camera.filter_param = desc->camera.filter_param;
    sc.camera.medium_id = desc->camera.medium_id;
    sc.options.integrator = desc->options.integrator;
    sc.options.spp = desc->options.samples_per_pixel;
    sc.options.max_depth = desc->options.max_depth;
    sc.options.rr_depth = desc->options.rr_depth;
    sc.options.vol_path_version = desc->options.vol_path_version;
    sc.options.max_null_collisions = desc->options.max_null_collisions;
    if (sc.camera.width <= 0 || sc.camera.height <= 0) return fail(LJ_ERR_INVALID, "bad film size");

    // ---- geometry pools
    std::vector<float> positions, normals, uvs, tri_cdf;
    std::vector<int> indices, prim_shape, prim_local;
    std::vector<DevShape> shapes(desc->num_shapes);
    int n_tris = 0, n_spheres = 0;
    for (int i = 0; i < desc->num_shapes; i++) {
        const lj_shape_desc &sd = desc->shapes[i];
        DevShape &sh = shapes[i];
        memset(&sh, 0, sizeof(sh));
        sh.type = sd.type;
        sh.material_id = sd.material_id;
        sh.area_light_id = sd.area_light_id;
        sh.interior_medium_id = sd.interior_medium_id;
        sh.exterior_medium_id = sd.exterior_medium_id;
        if (sd.material_id >= desc->num_materials) return fail(LJ_ERR_INVALID, "shape material_id out of range");
        if (sd.type == LJ_SHAPE_SPHERE) {
            sh.cx = sd.center[0]; sh.cy = sd.center[1]; sh.cz = sd.center[2];
            sh.radius = sd.radius;
            prim_shape.push_back(i);
            prim_local.push_back(0);
            n_spheres++;
            continue;
        }
        if (sd.type != LJ_SHAPE_MESH || !sd.positions || !sd.indices || sd.num_vertices <= 0 || sd.num_triangles < 0)
            return fail(LJ_ERR_INVALID, "malformed mesh shape");
        int vbase = (int)(positions.size() / 3);
        sh.vertex_offset = vbase;
        sh.tri_offset = (int)(indices.size() / 3);
        sh.num_tris = sd.num_triangles;
        sh.has_normals = sd.normals != nullptr;
        sh.has_uvs = sd.uvs != nullptr;
        positions.insert(positions.end(), sd.positions, sd.positions + 3 * (size_t)sd.num_vertices);
        if (sd.normals) normals.insert(normals.end(), sd.normals, sd.normals + 3 * (size_t)sd.num_vertices);
        else normals.resize(normals.size() + 3 * (size_t)sd.num_vertices, 0.f);
        if (sd.uvs) uvs.insert(uvs.end(), sd.uvs, sd.uvs + 2 * (size_t)sd.num_vertices);
        else uvs.resize(uvs.size() + 2 * (size_t)sd.num_vertices, 0.f);
        // triangle areas -> TableDist1D, triangle_mesh.inl:60-75 + table_dist.cpp:3-25 (double, then fp32)
        std::vector<double> cdf(sd.num_triangles + 1, 0.0);
        for (int t = 0; t < sd.num_triangles; t++) {
            int i0 = sd.indices[3 * t], i1 = sd.indices[3 * t + 1], i2 = sd.indices[3 * t + 2];
            if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= sd.num_vertices || i1 >= sd.num_vertices || i2 >= sd.num_vertices)
                return fail(LJ_ERR_INVALID, "mesh index out of range");
            indices.push_back(vbase + i0); indices.push_back(vbase + i1); indices.push_back(vbase + i2);
            const float *p0 = sd.positions + 3 * i0, *p1 = sd.positions + 3 * i1, *p2 = sd.positions + 3 * i2;
            double e1[3], e2[3];
            for (int c = 0; c < 3; c++) { e1[c] = (double)p1[c] - p0[c]; e2[c] = (double)p2[c] - p0[c]; }
            double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
            cdf[t + 1] = cdf[t] + sqrt(cx * cx + cy * cy + cz * cz) / 2;
            prim_shape.push_back(i);
            prim_local.push_back(t);
        }
        double total = cdf[sd.num_triangles];
        sh.total_area = (float)total;
        sh.cdf_offset = (int)tri_cdf.size();
        if (sd.area_light_id >= 0) {  // the table is only ever sampled for emitters
            for (int t = 0; t <= sd.num_triangles; t++) {
                double v = total > 0 ? cdf[t] / total : (double)t / std::max(sd.num_triangles, 1);
                tri_cdf.push_back((float)v);
            }
            tri_cdf.back() = 1.f;  // upstream leaves cdf[n] un-normalised (TRAP a18); same samples either way
        }
        n_tris += sd.num_triangles;
    }
    if (prim_shape.empty()) return fail(LJ_ERR_INVALID, "scene has no primitives");
    std::vector<float> ref_box;
    const int n_primitives = (int)prim_shape.size();
    const bool split = split_references(shapes, positions, indices, prim_shape, prim_local, ref_box);
    int n_prims = (int)prim_shape.size();  // primitive REFERENCES from here on (== primitives when nothing was split)
    sc.positions = up.upload(positions);
    sc.normals = up.upload(normals);
    sc.uvs = up.upload(uvs);
    sc.indices = up.upload(indices);
    sc.tri_cdf = up.upload(tri_cdf);
    sc.shapes = up.upload(shapes);
    sc.num_shapes = desc->num_shapes;
    int *d_prim_shape = up.upload(prim_shape);
    int *d_prim_local = up.upload(prim_local);
    float *d_ref_box = split ? up.upload(ref_box) : nullptr;
    if (up.err != cudaSuccess) { int r = cuda_fail(up.err, "geometry upload"); lj_scene_destroy(s); return r; }

    cudaStreamCreate(&s->stream);
    for (auto &e : s->ev) cudaEventCreate(&e);

    // ---- images + mip chains (mipmap.h:24-48) built on the device
    std::vector<DevImage> images(std::max(desc->num_images, 0));
    size_t total1 = 0, total3 = 0;
    for (int i = 0; i < desc->num_images; i++) {
        const lj_image_desc &im = desc->images[i];
        if (im.width <= 0 || im.height <= 0 || (im.channels != 1 && im.channels != 3) || !im.data)
            return fail(LJ_ERR_INVALID, "malformed image");
        DevImage &d = images[i];
        memset(&d, 0, sizeof(d));
        d.channels = im.channels;
        int size = std::max(im.width, im.height);
        d.levels = std::min((int)ceil(log2((double)size) + 1), kMaxMipLevels);
        int w = im.width, h = im.height;
        size_t &total = im.channels == 3 ? total3 : total1;
        for (int l = 0; l < d.levels; l++) {
            d.w[l] = w; d.h[l] = h; d.offset[l] = (int)total;
            total += (size_t)w * h;
            w = std::max(w / 2, 1); h = std::max(h / 2, 1);
        }
    }
    V4 *texels3 = up.alloc<V4>(total3);
    float *texels1 = up.alloc<float>(total1);
    if (up.err != cudaSuccess) { int r = cuda_fail(up.err, "texture alloc"); lj_scene_destroy(s); return r; }
    auto t_prep0 = std::chrono::steady_clock::now();
    for (int i = 0; i < desc->num_images; i++) {
        const lj_image_desc &im = desc->images[i];
        const DevImage &d = images[i];
        size_t n0 = (size_t)im.width * im.height;
        if (im.channels == 3) {
            std::vector<V4> l0(n0);
            for (size_t k = 0; k < n0; k++) l0[k] = mk4(im.data[3 * k], im.data[3 * k + 1], im.data[3 * k + 2], 0.f);
            cudaMemcpyAsync(texels3 + d.offset[0], l0.data(), n0 * sizeof(V4), cudaMemcpyHostToDevice, s->stream);
            cudaStreamSynchronize(s->stream);
            for (int l = 1; l < d.levels; l++) {
                int n = d.w[l] * d.h[l];
                LJ_LAUNCH(k_mip_down3, (n + 255) / 256, 256, s->stream, texels3 + d.offset[l - 1], d.w[l - 1], d.h[l - 1], texels3 + d.offset[l], d.w[l], d.h[l]);
            }
        } else {
            cudaMemcpyAsync(texels1 + d.offset[0], im.data, n0 * sizeof(float), cudaMemcpyHostToDevice, s->stream);
            for (int l = 1; l < d.levels; l++) {
                int n = d.w[l] * d.h[l];
                LJ_LAUNCH(k_mip_down1, (n + 255) / 256, 256, s->stream, texels1 + d.offset[l - 1], d.w[l - 1], d.h[l - 1], texels1 + d.offset[l], d.w[l], d.h[l]);
            }
        }
    }
    {
        cudaError_t e = cudaStreamSynchronize(s->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { int r = cuda_fail(e, "mip build"); lj_scene_destroy(s); return r; }
    }
    s->h_images3 = images;
    sc.images1 = sc.images3 = up.upload(images);
    sc.texels1 = texels1;
    sc.texels3 = texels3;

    // ---- materials
    std::vector<DevMaterial> mats(std::max(desc->num_materials, 0));
    for (int i = 0; i < desc->num_materials; i++) {
        const lj_material_desc &md = desc->materials[i];
        DevMaterial &m = mats[i];
        memset(&m, 0, sizeof(m));
        m.type = md.type;
        m.eta = md.eta;
        if (md.type < 0 || md.type > LJ_MAT_DISNEY_BSDF) return fail(LJ_ERR_INVALID, "unknown material type");
        for (int k = 0; k < LJ_NUM_TEX_SLOTS; k++) {
            m.tex[k] = conv_texture(md.tex[k]);
            if (md.tex[k].kind == LJ_TEX_IMAGE && (md.tex[k].image_id < 0 || md.tex[k].image_id >= desc->num_images))
                return fail(LJ_ERR_INVALID, "texture image_id out of range");
        }
    }
    sc.materials = up.upload(mats);
    sc.num_materials = desc->num_materials;

    // ---- BVH (K0)
    BvhResult bvh;
    auto t_bvh0 = std::chrono::steady_clock::now();
    {
        cudaError_t e = build_bvh8(sc, d_prim_shape, d_prim_local, d_ref_box, n_prims, s->stream, &bvh);
        if (e != cudaSuccess) { int r = cuda_fail(e, "BVH build"); lj_scene_destroy(s); return r; }
    }
    auto t_bvh1 = std::chrono::steady_clock::now();
    s->allocations.push_back(bvh.nodes);
    s->allocations.push_back(bvh.prims);
    s->info.device_bytes += (int64_t)bvh.num_nodes * sizeof(DevNode8) + (int64_t)n_prims * sizeof(DevPrim);
    if (bvh.num_prims_placed != n_prims) return fail(LJ_ERR_CUDA, "BVH collapse lost primitives");
    if (bvh.depth > kBvh8StackLimit) return fail(LJ_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
    sc.nodes8 = bvh.nodes;
    sc.prims = bvh.prims;
    sc.num_prims = n_prims;

    // ---- scene bounds -> bounding sphere + epsilons (scene.cpp:29-34, scene.h:99-105)
    {
        double d2 = 0;
        for (int c = 0; c < 3; c++) {
            double lo = (&bvh.bounds.lo.x)[c], hi = (&bvh.bounds.hi.x)[c];
            d2 += (hi - lo) * (hi - lo);
            (&sc.bsphere_center.x)[c] = (float)((lo + hi) / 2);
            s->info.bounds_lo[c] = (float)lo;
            s->info.bounds_hi[c] = (float)hi;
            s->info.bsphere_center[c] = (float)((lo + hi) / 2);
        }
        double radius = sqrt(d2) / 2;
        sc.bsphere_radius = (float)radius;
        sc.shadow_eps = sc.isect_eps = (float)std::min(radius * 1e-5, 0.01);
        s->info.bsphere_radius = sc.bsphere_radius;
        s->info.shadow_epsilon = sc.shadow_eps;
    }

    // ---- lights: envmap table (envmap.inl:75-98, table_dist.cpp:40-115), power table (scene.cpp:48-52)
    std::vector<DevLight> lights(std::max(desc->num_lights, 0));
    std::vector<double> power(lights.size(), 0.0);
    sc.envmap_light_id = desc->envmap_light_id;
    for (int i = 0; i < desc->num_lights; i++) {
        const lj_light_desc &ld = desc->lights[i];
        DevLight &l = lights[i];
        memset(&l, 0, sizeof(l));
        l.type = ld.type;
        l.shape_id = ld.shape_id;
        for (int c = 0; c < 3; c++) l.intensity[c] = ld.intensity[c];
        if (ld.type == LJ_LIGHT_AREA) {
            if (ld.shape_id < 0 || ld.shape_id >= desc->num_shapes) return fail(LJ_ERR_INVALID, "light shape_id out of range");
            const DevShape &sh = shapes[ld.shape_id];
            double area = sh.type == 0 ? 4 * M_PI * (double)sh.radius * sh.radius : (double)sh.total_area;
            power[i] = lum(ld.intensity) * area * M_PI;  // diffuse_area_light.inl:1-3
            continue;
        }
        l.values = conv_texture(ld.values);
        memcpy(l.to_world.m, ld.to_world, 64);
        memcpy(l.to_local.m, ld.to_local, 64);
        l.scale = ld.scale;
        if (ld.values.kind != LJ_TEX_IMAGE) { power[i] = 0; continue; }  // envmap.inl:76: only image envmaps get a table
        if (ld.values.image_id < 0 || ld.values.image_id >= desc->num_images || desc->images[ld.values.image_id].channels != 3)
            return fail(LJ_ERR_INVALID, "envmap image invalid");
        const lj_image_desc &im = desc->images[ld.values.image_id];
        int w = im.width, h = im.height;
        std::vector<double> f((size_t)w * h);
        for (int y = 0; y < h; y++) {
            double sin_el = sin(M_PI * (y + 0.5) / h);
            for (int x = 0; x < w; x++) f[(size_t)y * w + x] = lum(im.data + 3 * ((size_t)y * w + x)) * sin_el;
        }
        std::vector<float> cdf_rows((size_t)h * (w + 1)), pdf_rows((size_t)h * w), cdf_m(h + 1), pdf_m(h);
        std::vector<double> row_int(h);
        for (int y = 0; y < h; y++) {
            std::vector<double> c(w + 1, 0.0);
            for (int x = 0; x < w; x++) c[x + 1] = c[x] + f[(size_t)y * w + x];
            double integral = c[w];
            row_int[y] = integral;
            for (int x = 0; x < w; x++) {
                cdf_rows[(size_t)y * (w + 1) + x] = (float)(integral > 0 ? c[x] / integral : (double)x / w);
                pdf_rows[(size_t)y * w + x] = (float)(integral > 0 ? f[(size_t)y * w + x] / integral : 1.0 / w);
            }
            cdf_rows[(size_t)y * (w + 1) + w] = 1.f;
        }
        std::vector<double> cm(h + 1, 0.0);
        for (int y = 0; y < h; y++) cm[y + 1] = cm[y] + row_int[y];
        double total_values = cm[h];
        for (int y = 0; y < h; y++) {
            cdf_m[y] = (float)(total_values > 0 ? cm[y] / total_values : (double)y / h);
            pdf_m[y] = (float)(total_values > 0 ? row_int[y] / total_values : 1.0 / h);
        }
        cdf_m[h] = 1.f;
        sc.envmap_dist.width = w;
        sc.envmap_dist.height = h;
        sc.envmap_dist.total_values = (float)total_values;
        sc.envmap_dist.cdf_rows = up.upload(cdf_rows);
        sc.envmap_dist.pdf_rows = up.upload(pdf_rows);
        sc.envmap_dist.cdf_marginals = up.upload(cdf_m);
        sc.envmap_dist.pdf_marginals = up.upload(pdf_m);
        double R = sc.bsphere_radius;
        power[i] = M_PI * R * R * total_values / ((double)w * h);  // envmap.inl:1-5
    }
    {
        int n = (int)lights.size();
        s->h_light_pmf.assign(n, 0.f);
        s->h_light_cdf.assign(n + 1, 0.f);
        std::vector<double> cdf(n + 1, 0.0);
        for (int i = 0; i < n; i++) cdf[i + 1] = cdf[i] + power[i];
        double total = n ? cdf[n] : 0;
        for (int i = 0; i < n; i++) {
            s->h_light_pmf[i] = (float)(total > 0 ? power[i] / total : 1.0 / n);
            s->h_light_cdf[i] = (float)(total > 0 ? cdf[i] / total : (double)i / n);
        }
        s->h_light_cdf[n] = 1.f;
    }
    sc.lights = up.upload(lights);
    sc.num_lights = desc->num_lights;
    sc.light_pmf = up.upload(s->h_light_pmf);
    sc.light_cdf = up.upload(s->h_light_cdf);

    // ---- media
    std::vector<DevMedium> media(std::max(desc->num_media, 0));
    for (int i = 0; i < desc->num_media; i++) {
        const lj_medium_desc &md = desc->media[i];
        DevMedium &m = media[i];
        memset(&m, 0, sizeof(m));
        m.type = md.type;
        m.phase_type = md.phase_type;
        m.phase_g = md.phase_g;
        for (int c = 0; c < 3; c++) { m.sigma_a[c] = md.sigma_a[c]; m.sigma_s[c] = md.sigma_s[c]; }
        if (md.type == LJ_MEDIUM_HETEROGENEOUS) {
            s->has_grid_media = true;
            m.albedo = conv_volume(md.albedo, up);
            m.density = conv_volume(md.density, up, true);
        }
    }
    for (int i = 0; i < desc->num_materials; i++) if (desc->materials[i].type >= LJ_MAT_DISNEY_DIFFUSE) s->has_disney = true;
    s->only_lambertian = desc->num_materials > 0;
    for (int i = 0; i < desc->num_materials; i++) if (desc->materials[i].type != LJ_MAT_LAMBERTIAN) s->only_lambertian = false;
    sc.media = up.upload(media);
    sc.num_media = desc->num_media;
    if (up.err != cudaSuccess) { int r = cuda_fail(up.err, "table upload"); lj_scene_destroy(s); return r; }
    cudaDeviceSynchronize();

    auto t_end = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    s->info.num_prims = n_primitives;
    s->info.num_prim_refs = n_prims;
    s->info.num_triangles = n_tris;
    s->info.num_spheres = n_spheres;
    s->info.num_bvh_nodes = bvh.num_nodes;
    s->info.bvh_width = 8;
    s->info.bvh_depth = bvh.depth;
    s->info.bvh_build_ms = ms(t_bvh0, t_bvh1);
    s->info.prep_ms = ms(t_prep0, t_bvh0);
    s->info.upload_ms = ms(t_start, t_end) - s->info.bvh_build_ms - s->info.prep_ms;
    s->info.sah_cost = bvh.sah_cost;
    *out = s;
    return LJ_OK;
}

extern "C" int lj_scene_get_info(lj_scene *s, lj_scene_info *info) {
    if (!s || !info) { set_error("null argument"); return LJ_ERR_INVALID; }
    *info = s->info;
    return LJ_OK;
}

extern "C" int lj_scene_get_light_table(lj_scene *s, float *pmf, float *cdf) {
    if (!s) { set_error("null argument"); return LJ_ERR_INVALID; }
    if (pmf) memcpy(pmf, s->h_light_pmf.data(), s->h_light_pmf.size() * sizeof(float));
    if (cdf) memcpy(cdf, s->h_light_cdf.data(), s->h_light_cdf.size() * sizeof(float));
    return LJ_OK;
}

extern "C" int lj_scene_get_mip_level(lj_scene *s, int32_t channels, int32_t image_id, int32_t level,
                                      int32_t *width, int32_t *height, float *data) {
    if (!s || image_id < 0 || image_id >= (int)s->h_images3.size()) { set_error("bad image id"); return LJ_ERR_INVALID; }
    DeviceGuard guard(s->device);
    const DevImage &d = s->h_images3[image_id];
    if (d.channels != channels || level < 0 || level >= d.levels) { set_error("bad level/channels"); return LJ_ERR_INVALID; }
    if (width) *width = d.w[level];
    if (height) *height = d.h[level];
    if (!data) return LJ_OK;
    size_t n = (size_t)d.w[level] * d.h[level];
    if (channels == 1) {
        LJ_CUDA(cudaMemcpy(data, s->dev.texels1 + d.offset[level], n * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        std::vector<V4> tmp(n);
        LJ_CUDA(cudaMemcpy(tmp.data(), s->dev.texels3 + d.offset[level], n * sizeof(V4), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < n; k++) { data[3 * k] = tmp[k].x; data[3 * k + 1] = tmp[k].y; data[3 * k + 2] = tmp[k].z; }
    }
    return LJ_OK;
}

extern "C" int lj_scene_create(const lj_scene_desc *desc, lj_scene **out) {
    if (!desc || !out) { set_error("null argument"); return LJ_ERR_INVALID; }
    int r = scene_create_here(desc, out);
    if (r != LJ_OK) return r;
#if !defined(LJ_HOSTSIM)
    // one replica per further device of lj_init (the scene is replicated on every GPU), built
    // concurrently, one host thread per device.  A single-device lj_init (or a scene created on a device that is not
    // the primary one) has none.
    std::vector<int> devs = init_devices();
    lj_scene *primary = *out;
    if (devs.size() > 1 && devs[0] == primary->device) {
        const size_t n = devs.size() - 1;
        std::vector<lj_scene *> reps(n, nullptr);
        std::vector<int> codes(n, LJ_OK);
        std::vector<std::string> errors(n);
        std::vector<std::thread> threads;
        for (size_t k = 0; k < n; k++)
            threads.emplace_back([&, k] {
                if (cudaSetDevice(devs[k + 1]) != cudaSuccess) { codes[k] = LJ_ERR_CUDA; errors[k] = "cudaSetDevice failed"; return; }
                codes[k] = scene_create_here(desc, &reps[k]);
                if (codes[k] != LJ_OK) errors[k] = lj_last_error();
            });
        for (auto &t : threads) t.join();
        for (size_t k = 0; k < n; k++)
            if (codes[k] != LJ_OK) {
                for (lj_scene *rp : reps) if (rp) lj_scene_destroy(rp);
                lj_scene_destroy(primary);
                *out = nullptr;
                set_error("replica on device " + std::to_string(devs[k + 1]) + ": " + errors[k]);
                return codes[k];
            }
        primary->replicas = reps;
    }
#endif
    return LJ_OK;
}
