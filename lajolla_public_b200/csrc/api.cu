// api.cu -- query seam S2 of the C ABI: batch forms of the reference's per-ray functions, used by
// the parity tests (ray parity vs the Embree-API oracle, BSDF/light/camera/texture/PCG parity vs
// the reference's own object code).  Every entry point runs the same LJ_HD device functions the
// wavefront kernels call.
#include "scene.cuh"
#include "lj_media.h"

#include <algorithm>

namespace lj {
namespace {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    cudaError_t err = cudaSuccess;
    explicit DevBuf(size_t n) { err = cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)); }
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t up(const T *h, size_t n) { return cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice); }
    cudaError_t down(T *h, size_t n) { return cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost); }
};

__global__ void __launch_bounds__(256) k_trace_closest(const LJ_GRID_CONSTANT DevScene sc, const lj_ray *rays, long long n, lj_hit *hits) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lj_ray r = rays[i];
    V3 o = mk3(r.org[0], r.org[1], r.org[2]), d = mk3(r.dir[0], r.dir[1], r.dir[2]);
    Hit h;
    trace8<false>(sc.nodes8, sc.prims, o, d, r.tnear, r.tfar, h);
    lj_hit out;
    hit_to_abi(sc, o, d, h, out);
    hits[i] = out;
}

__global__ void __launch_bounds__(256) k_trace_any(const LJ_GRID_CONSTANT DevScene sc, const lj_ray *rays, long long n, uint8_t *occ) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lj_ray r = rays[i];
    Hit h;
    occ[i] = trace8<true>(sc.nodes8, sc.prims, mk3(r.org[0], r.org[1], r.org[2]), mk3(r.dir[0], r.dir[1], r.dir[2]), r.tnear, r.tfar, h) ? 1 : 0;
}

LJ_HD void v3_out(float *o, V3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
LJ_HD V3 v3_in(const float *p) { return mk3(p[0], p[1], p[2]); }

LJ_HD void vertex_to_abi(const Vertex &v, lj_vertex &o) {
    v3_out(o.position, v.position); v3_out(o.geometric_normal, v.geometric_normal);
    v3_out(o.frame_x, v.shading_frame.x); v3_out(o.frame_y, v.shading_frame.y); v3_out(o.frame_n, v.shading_frame.n);
    o.st[0] = v.st.x; o.st[1] = v.st.y; o.uv[0] = v.uv.x; o.uv[1] = v.uv.y;
    o.uv_screen_size = v.uv_screen_size; o.mean_curvature = v.mean_curvature; o.ray_radius = v.ray_radius;
    o.shape_id = v.shape_id; o.primitive_id = v.primitive_id; o.material_id = v.material_id;
    o.interior_medium_id = v.interior_medium_id; o.exterior_medium_id = v.exterior_medium_id;
}
LJ_HD Vertex vertex_from_abi(const lj_vertex &o) {
    Vertex v;
    v.position = v3_in(o.position); v.geometric_normal = v3_in(o.geometric_normal);
    v.shading_frame = make_frame(v3_in(o.frame_x), v3_in(o.frame_y), v3_in(o.frame_n));
    v.st = mk2(o.st[0], o.st[1]); v.uv = mk2(o.uv[0], o.uv[1]);
    v.uv_screen_size = o.uv_screen_size; v.mean_curvature = o.mean_curvature; v.ray_radius = o.ray_radius;
    v.shape_id = o.shape_id; v.primitive_id = o.primitive_id; v.material_id = o.material_id;
    v.interior_medium_id = o.interior_medium_id; v.exterior_medium_id = o.exterior_medium_id;
    return v;
}

__global__ void __launch_bounds__(128) k_intersect(const LJ_GRID_CONSTANT DevScene sc, const lj_ray *rays, const float *rd, long long n, lj_vertex *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lj_ray r = rays[i];
    V3 o = v3_in(r.org), d = v3_in(r.dir);
    Hit h;
    lj_vertex ov;
    memset(&ov, 0, sizeof(ov));
    ov.shape_id = ov.primitive_id = ov.material_id = ov.interior_medium_id = ov.exterior_medium_id = -1;
    if (trace8<false>(sc.nodes8, sc.prims, o, d, r.tnear, r.tfar, h)) {
        Vertex v = make_vertex(sc, o, d, h, rd ? rd[2 * i] : 0.f, rd ? rd[2 * i + 1] : 0.f);
        vertex_to_abi(v, ov);
    }
    out[i] = ov;
}

__global__ void __launch_bounds__(128) k_bsdf(const LJ_GRID_CONSTANT DevScene sc, const lj_bsdf_query *q, long long n, lj_bsdf_result *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Vertex v = vertex_from_abi(q[i].vertex);
    lj_bsdf_result r;
    memset(&r, 0, sizeof(r));
    if (v.material_id >= 0 && v.material_id < sc.num_materials) {
        const DevMaterial &m = sc.materials[v.material_id];
        V3 wi = v3_in(q[i].dir_in), wo = v3_in(q[i].dir_out);
        v3_out(r.f, bsdf_eval(sc, m, wi, wo, v, q[i].transport));
        r.pdf = bsdf_pdf(sc, m, wi, wo, v);
        BsdfSample s;
        if (bsdf_sample(sc, m, wi, v, mk2(q[i].rnd_uv[0], q[i].rnd_uv[1]), q[i].rnd_w, s)) {
            r.sampled = 1;
            v3_out(r.s_dir_out, s.dir_out);
            r.s_eta = s.eta;
            r.s_roughness = s.roughness;
        }
    }
    out[i] = r;
}

__global__ void __launch_bounds__(128) k_medium(const LJ_GRID_CONSTANT DevScene sc, const lj_medium_query *q, long long n, lj_medium_result *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lj_medium_query in = q[i];
    const DevMedium &m = sc.media[in.medium_id];
    V3 o = v3_in(in.org), d = v3_in(in.dir);
    lj_medium_result r;
    v3_out(r.majorant, medium_majorant(m, o, d, in.tfar));
    V3 sa, ss;
    medium_sigmas(m, o + d * in.t, sa, ss);
    v3_out(r.sigma_a, sa); v3_out(r.sigma_s, ss);
    V3 pd = phase_sample(m, -d, mk2(in.rnd[0], in.rnd[1]));
    v3_out(r.phase_dir, pd);
    r.phase_eval = phase_eval(m, -d, pd);
    r.phase_pdf = phase_pdf(m, -d, pd);
    out[i] = r;
}

__global__ void __launch_bounds__(128) k_medium_bound(const LJ_GRID_CONSTANT DevScene sc, const lj_medium_query *q, long long n, lj_medium_bound *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lj_medium_query in = q[i];
    const DevMedium &m = sc.media[in.medium_id];
    V3 o = v3_in(in.org), d = v3_in(in.dir);
    lj_medium_bound r;
    r.local = medium_has_local_majorant(m) ? 1 : 0;
    if (r.local) {
        v3_out(r.majorant, volume_block_majorant(m.density, o, d, in.t, r.t_exit));
    } else {
        v3_out(r.majorant, medium_majorant(m, o, d, in.tfar));
        r.t_exit = LJ_INF;
    }
    v3_out(r.sigma_t, medium_sigma_t(m, o + d * in.t));
    out[i] = r;
}

__global__ void __launch_bounds__(128) k_light(const LJ_GRID_CONSTANT DevScene sc, const lj_light_query *q, long long n, lj_light_result *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lj_light_result r;
    memset(&r, 0, sizeof(r));
    V3 ref = v3_in(q[i].ref_point);
    int id = sample_light(sc, q[i].light_w);
    const DevLight &l = sc.lights[id];
    PointAndNormal pn = sample_point_on_light(sc, l, ref, mk2(q[i].rnd_uv[0], q[i].rnd_uv[1]), q[i].rnd_w);
    r.light_id = id;
    v3_out(r.position, pn.position);
    v3_out(r.normal, pn.normal);
    r.pmf = light_pmf(sc, id);
    r.pdf = pdf_point_on_light(sc, l, pn, ref);
    V3 dir_light = light_is_envmap(l) ? -pn.normal : normalize(pn.position - ref);
    v3_out(r.emission, light_emission(sc, l, -dir_light, 0.f, pn));
    out[i] = r;
}

__global__ void k_camera(const LJ_GRID_CONSTANT DevScene sc, const float *xy, long long n, lj_ray *rays) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 o, d;
    sample_primary(sc.camera, mk2(xy[2 * i], xy[2 * i + 1]), o, d);
    lj_ray r;
    v3_out(r.org, o); v3_out(r.dir, d);
    r.tnear = 0; r.tfar = LJ_INF;
    rays[i] = r;
}

__global__ void k_texture(const LJ_GRID_CONSTANT DevScene sc, int material_id, int slot, const float *q, long long n, float *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevTexture &t = sc.materials[material_id].tex[slot];
    V3 v;
    bool one_channel = t.kind == 1 && sc.images3[t.image_id].channels == 1;
    if (one_channel) v = eval_texture<1>(sc, t, mk2(q[3 * i], q[3 * i + 1]), q[3 * i + 2]);
    else v = eval_texture<3>(sc, t, mk2(q[3 * i], q[3 * i + 1]), q[3 * i + 2]);
    v3_out(out + 3 * i, v);
}

__global__ void k_pcg(uint64_t first_stream, uint64_t seed, int n_streams, int n_draws, uint32_t *out_u32, float *out_f32) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_streams) return;
    Pcg a = pcg_init(first_stream + i, seed), b = a;
    for (int k = 0; k < n_draws; k++) {
        if (out_u32) out_u32[(size_t)i * n_draws + k] = pcg_next(a);
        if (out_f32) out_f32[(size_t)i * n_draws + k] = pcg_uniform(b);
    }
}

int grid_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }

#if !defined(LJ_HOSTSIM)
// read-only stream of 16-byte loads over a working set (roofline denominators measured on the spot)
__global__ void __launch_bounds__(256) k_read_stream(const float4 *src, long long n16, int iters, float *sink) {
    float acc = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; it++) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += 4 * stride) {
            // four independent loads in flight per thread
            float4 a = src[i];
            float4 b = i + stride < n16 ? src[i + stride] : make_float4(0, 0, 0, 0);
            float4 c = i + 2 * stride < n16 ? src[i + 2 * stride] : make_float4(0, 0, 0, 0);
            float4 d = i + 3 * stride < n16 ? src[i + 3 * stride] : make_float4(0, 0, 0, 0);
            acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w);
        }
    }
    if (acc == 12345.678f) *sink = acc;  // never true for the zero-filled buffer: keeps the loads alive
}
#endif

}  // namespace
}  // namespace lj

using namespace lj;

#define LJ_CHECK_ARGS(cond) do { if (!(cond)) { set_error("invalid argument"); return LJ_ERR_INVALID; } } while (0)

static int timed_done(cudaEvent_t e0, cudaEvent_t e1, double *kernel_ms) {
    LJ_CUDA(cudaEventSynchronize(e1));
    LJ_CUDA(cudaGetLastError());
    if (kernel_ms) { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); *kernel_ms = ms; }
    return LJ_OK;
}

extern "C" int lj_trace_closest(lj_scene *s, const lj_ray *rays, int64_t n, lj_hit *hits, double *kernel_ms) {
    LJ_CHECK_ARGS(s && rays && hits && n >= 0);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    DevBuf<lj_ray> dr(n); DevBuf<lj_hit> dh(n);
    LJ_CUDA(dr.err); LJ_CUDA(dh.err);
    LJ_CUDA(dr.up(rays, n));
    LJ_CUDA(cudaEventRecord(s->ev[0], s->stream));
    LJ_LAUNCH(k_trace_closest, grid_for(n, 256), 256, s->stream, s->dev, dr.p, n, dh.p);
    LJ_CUDA(cudaEventRecord(s->ev[1], s->stream));
    int r = timed_done(s->ev[0], s->ev[1], kernel_ms);
    if (r != LJ_OK) return r;
    LJ_CUDA(dh.down(hits, n));
    return LJ_OK;
}

extern "C" int lj_trace_any(lj_scene *s, const lj_ray *rays, int64_t n, uint8_t *occluded, double *kernel_ms) {
    LJ_CHECK_ARGS(s && rays && occluded && n >= 0);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    DevBuf<lj_ray> dr(n); DevBuf<uint8_t> dq(n);
    LJ_CUDA(dr.err); LJ_CUDA(dq.err);
    LJ_CUDA(dr.up(rays, n));
    LJ_CUDA(cudaEventRecord(s->ev[0], s->stream));
    LJ_LAUNCH(k_trace_any, grid_for(n, 256), 256, s->stream, s->dev, dr.p, n, dq.p);
    LJ_CUDA(cudaEventRecord(s->ev[1], s->stream));
    int r = timed_done(s->ev[0], s->ev[1], kernel_ms);
    if (r != LJ_OK) return r;
    LJ_CUDA(dq.down(occluded, n));
    return LJ_OK;
}

extern "C" int lj_intersect(lj_scene *s, const lj_ray *rays, const float *rd, int64_t n, lj_vertex *vertices) {
    LJ_CHECK_ARGS(s && rays && vertices && n >= 0);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    DevBuf<lj_ray> dr(n); DevBuf<lj_vertex> dv(n); DevBuf<float> dd(rd ? 2 * n : 1);
    LJ_CUDA(dr.err); LJ_CUDA(dv.err); LJ_CUDA(dd.err);
    LJ_CUDA(dr.up(rays, n));
    if (rd) LJ_CUDA(dd.up(rd, 2 * n));
    LJ_LAUNCH(k_intersect, grid_for(n, 128), 128, s->stream, s->dev, dr.p, rd ? dd.p : nullptr, n, dv.p);
    LJ_CUDA(cudaStreamSynchronize(s->stream));
    LJ_CUDA(cudaGetLastError());
    LJ_CUDA(dv.down(vertices, n));
    return LJ_OK;
}

extern "C" int lj_bsdf_batch(lj_scene *s, const lj_bsdf_query *q, int64_t n, lj_bsdf_result *out) {
    LJ_CHECK_ARGS(s && q && out && n >= 0);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    DevBuf<lj_bsdf_query> dq(n); DevBuf<lj_bsdf_result> dr(n);
    LJ_CUDA(dq.err); LJ_CUDA(dr.err);
    LJ_CUDA(dq.up(q, n));
    LJ_LAUNCH(k_bsdf, grid_for(n, 128), 128, s->stream, s->dev, dq.p, n, dr.p);
    LJ_CUDA(cudaStreamSynchronize(s->stream));
    LJ_CUDA(cudaGetLastError());
    LJ_CUDA(dr.down(out, n));
    return LJ_OK;
}

extern "C" int lj_medium_batch(lj_scene *s, const lj_medium_query *q, int64_t n, lj_medium_result *out) {
    LJ_CHECK_ARGS(s && q && out && n >= 0);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    for (int64_t i = 0; i < n; i++)
        if (q[i].medium_id < 0 || q[i].medium_id >= s->dev.num_media) { set_error("medium id out of range"); return LJ_ERR_INVALID; }
    DevBuf<lj_medium_query> dq(n); DevBuf<lj_medium_result> dr(n);
    LJ_CUDA(dq.err); LJ_CUDA(dr.err);
    LJ_CUDA(dq.up(q, n));
    LJ_LAUNCH(k_medium, grid_for(n, 128), 128, s->stream, s->dev, dq.p, n, dr.p);
    LJ_CUDA(cudaStreamSynchronize(s->stream));
    LJ_CUDA(cudaGetLastError());
    LJ_CUDA(dr.down(out, n));
    return LJ_OK;
}

extern "C" int lj_medium_bound_batch(lj_scene *s, const lj_medium_query *q, int64_t n, lj_medium_bound *out) {
    LJ_CHECK_ARGS(s && q && out && n >= 0);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    for (int64_t i = 0; i < n; i++)
        if (q[i].medium_id < 0 || q[i].medium_id >= s->dev.num_media) { set_error("medium id out of range"); return LJ_ERR_INVALID; }
    DevBuf<lj_medium_query> dq(n); DevBuf<lj_medium_bound> dr(n);
    LJ_CUDA(dq.err); LJ_CUDA(dr.err);
    LJ_CUDA(dq.up(q, n));
    LJ_LAUNCH(k_medium_bound, grid_for(n, 128), 128, s->stream, s->dev, dq.p, n, dr.p);
    LJ_CUDA(cudaStreamSynchronize(s->stream));
    LJ_CUDA(cudaGetLastError());
    LJ_CUDA(dr.down(out, n));
    return LJ_OK;
}

extern "C" int lj_light_batch(lj_scene *s, const lj_light_query *q, int64_t n, lj_light_result *out) {
    LJ_CHECK_ARGS(s && q && out && n >= 0);
    DeviceGuard guard(s->device);
    if (s->dev.num_lights <= 0) { set_error("scene has no lights"); return LJ_ERR_INVALID; }
    if (n == 0) return LJ_OK;
    DevBuf<lj_light_query> dq(n); DevBuf<lj_light_result> dr(n);
    LJ_CUDA(dq.err); LJ_CUDA(dr.err);
    LJ_CUDA(dq.up(q, n));
    LJ_LAUNCH(k_light, grid_for(n, 128), 128, s->stream, s->dev, dq.p, n, dr.p);
    LJ_CUDA(cudaStreamSynchronize(s->stream));
    LJ_CUDA(cudaGetLastError());
    LJ_CUDA(dr.down(out, n));
    return LJ_OK;
}

extern "C" int lj_camera_rays(lj_scene *s, const float *xy, int64_t n, lj_ray *rays) {
    LJ_CHECK_ARGS(s && xy && rays && n >= 0);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    DevBuf<float> dq(2 * n); DevBuf<lj_ray> dr(n);
    LJ_CUDA(dq.err); LJ_CUDA(dr.err);
    LJ_CUDA(dq.up(xy, 2 * n));
    LJ_LAUNCH(k_camera, grid_for(n, 256), 256, s->stream, s->dev, dq.p, n, dr.p);
    LJ_CUDA(cudaStreamSynchronize(s->stream));
    LJ_CUDA(cudaGetLastError());
    LJ_CUDA(dr.down(rays, n));
    return LJ_OK;
}

extern "C" int lj_texture_batch(lj_scene *s, int32_t material_id, int32_t slot, const float *q, int64_t n, float *out_rgb) {
    LJ_CHECK_ARGS(s && q && out_rgb && n >= 0 && material_id >= 0 && material_id < s->dev.num_materials && slot >= 0 && slot < LJ_NUM_TEX_SLOTS);
    DeviceGuard guard(s->device);
    if (n == 0) return LJ_OK;
    DevBuf<float> dq(3 * n), dr(3 * n);
    LJ_CUDA(dq.err); LJ_CUDA(dr.err);
    LJ_CUDA(dq.up(q, 3 * n));
    LJ_LAUNCH(k_texture, grid_for(n, 256), 256, s->stream, s->dev, material_id, slot, dq.p, n, dr.p);
    LJ_CUDA(cudaStreamSynchronize(s->stream));
    LJ_CUDA(cudaGetLastError());
    LJ_CUDA(dr.down(out_rgb, 3 * n));
    return LJ_OK;
}

extern "C" int lj_pcg32_batch(uint64_t first_stream, uint64_t seed, int32_t n_streams, int32_t n_draws, uint32_t *out_u32, float *out_f32) {
    LJ_CHECK_ARGS(n_streams >= 0 && n_draws >= 0 && (out_u32 || out_f32));
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: libljb200 has no CPU path");
        return LJ_ERR_NO_DEVICE;
    }
    size_t n = (size_t)n_streams * n_draws;
    if (n == 0) return LJ_OK;
    DevBuf<uint32_t> du(out_u32 ? n : 1); DevBuf<float> df(out_f32 ? n : 1);
    LJ_CUDA(du.err); LJ_CUDA(df.err);
    LJ_LAUNCH(k_pcg, grid_for(n_streams, 128), 128, (cudaStream_t)0, first_stream, seed ? seed : kPcgDefaultSeed, n_streams, n_draws, out_u32 ? du.p : nullptr, out_f32 ? df.p : nullptr);
    LJ_CUDA(cudaDeviceSynchronize());
    LJ_CUDA(cudaGetLastError());
    if (out_u32) LJ_CUDA(du.down(out_u32, n));
    if (out_f32) LJ_CUDA(df.down(out_f32, n));
    return LJ_OK;
}

extern "C" int lj_measure_read_bandwidth(int64_t bytes, int32_t iters, double *gb_per_s) {
    LJ_CHECK_ARGS(bytes >= 16 && iters > 0 && gb_per_s);
#if defined(LJ_HOSTSIM)
    set_error("no device in the host simulation");
    return LJ_ERR_UNSUPPORTED;
#else
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_error("no CUDA device"); return LJ_ERR_NO_DEVICE; }
    LJ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long n16 = bytes / 16;
    float4 *buf = nullptr;
    float *sink = nullptr;
    LJ_CUDA(cudaMalloc(&buf, (size_t)n16 * 16));
    cudaError_t e = cudaMalloc(&sink, sizeof(float));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e == cudaSuccess) e = cudaMemset(buf, 0, (size_t)n16 * 16);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    float ms = 0;
    if (e == cudaSuccess) {
        const int grid = std::max(1, sms) * 8;
        LJ_LAUNCH(k_read_stream, grid, 256, (cudaStream_t)0, buf, n16, 2, sink);  // warm-up: brings the set into L2 if it fits
        cudaEventRecord(e0, 0);
        LJ_LAUNCH(k_read_stream, grid, 256, (cudaStream_t)0, buf, n16, iters, sink);
        cudaEventRecord(e1, 0);
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(buf);
    if (sink) cudaFree(sink);
    if (e != cudaSuccess) return cuda_fail(e, "read-bandwidth probe");
    *gb_per_s = (double)n16 * 16.0 * iters / (ms * 1e-3) / 1e9;
    return LJ_OK;
#endif
}
