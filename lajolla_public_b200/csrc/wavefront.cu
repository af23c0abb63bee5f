// wavefront.cu -- the wavefront path tracer: queue-driven regenerate / extend / shade / shadow kernels
// over a pool of path records resident in HBM, and the render entry points of the C ABI.
// Replaces the reference's tile thread pool (render.cpp:71-152, parallel.cpp) and the per-sample
// control flow of path_tracing.h.  Stage kernels:
//   k_regen   K1+K7  over the free queue: flush finished paths to the film, start new camera paths
//   k_trace<> K2/K3  persistent threads: each warp pulls rays from the extension / shadow queue with a
//                    warp-aggregated cursor and refills idle lanes while the others keep traversing
//   k_shade   K4     over the extension queue: emission+MIS, Russian roulette, NEE sample, BSDF
//                    sample (lj_path.h); warp-ballot compaction into next wave's queues (K6)
// Queues are arrays of slot indices in HBM with device-side counters; one wave = regen, extend,
// shade, shadow, counter reset.  The host reads one counter per wave to detect the end.
#include "scene.cuh"

#include <algorithm>
#include <vector>

namespace lj {

// device counters (unsigned int each)
enum {
    Q_EXT0 = 0, Q_EXT1, Q_FREE0, Q_FREE1, Q_SHADOW,  // queue lengths
    CUR_EXT, CUR_SHADOW,                               // fetch cursors of the persistent kernels
    Q_COUNT
};
// statistics (unsigned long long each)
enum { C_SAMPLES = 0, C_CLOSEST, C_SHADOW, C_BOUNCES, C_NEXT, C_COUNT };

struct WaveArgs {
    PathPool pool;
    RenderParams rp;
    unsigned long long *stats;     // C_COUNT
    unsigned int *qn;              // Q_COUNT
    int *q_ext[2], *q_free[2], *q_shadow;
    float *film;                   // w*h*4: sum rgb, n
    float *film_sq;                // w*h*4: sum of squares rgb (may be null)
    unsigned long long total_items;  // padded pixels * samples in this call
    int tiles_x, tiles_y;
    int cur;                       // parity of the wave: q_ext[cur] is traced and shaded now
};

__device__ __forceinline__ void warp_add(unsigned long long *ctr, unsigned v) {
    unsigned s = __reduce_add_sync(0xffffffffu, v);
    if (LJ_LANE() == 0 && s) atomicAdd(ctr, (unsigned long long)s);
}

// K6: warp-ballot stream compaction -- one atomic per warp, lanes get consecutive queue positions.
__device__ __forceinline__ void warp_push(int *queue, unsigned int *count, bool pred, int value) {
    unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (!mask) return;
    int lane = LJ_LANE();
    int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + __popc(mask & ((1u << lane) - 1))] = value;
}

// work item k -> (sample, pixel): 8x4 pixel tiles so one warp starts 32 neighbouring pixels
__device__ __forceinline__ bool item_to_pixel(const WaveArgs &a, unsigned long long k, uint32_t &pixel, uint32_t &sample) {
    unsigned long long per_sample = (unsigned long long)a.tiles_x * a.tiles_y * 32ull;
    sample = a.rp.sample_begin + (uint32_t)(k / per_sample);
    uint32_t p = (uint32_t)(k % per_sample);
    uint32_t tile = p >> 5, l = p & 31;
    uint32_t x = (tile % a.tiles_x) * 8 + (l & 7);
    uint32_t y = (tile / a.tiles_x) * 4 + (l >> 3);
    pixel = y * a.rp.width + x;
    return x < (uint32_t)a.rp.width && y < (uint32_t)a.rp.height;
}

// K1 + K7 over the free queue.
__global__ void __launch_bounds__(256) k_regen(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool need = i < a.qn[Q_FREE0 + a.cur];
    int slot = -1;
    unsigned finished = 0;
    if (need) {
        slot = a.q_free[a.cur][i];
        V4 meta = a.pool.meta[slot];
        if (f2u(meta.y) & kOccupied) {
            // K7: accumulate the finished sample (render.cpp:92) -- non-finite samples are dropped
            V4 r = a.pool.rad[slot];
            uint32_t pixel = f2u(meta.x);
            if (is_finite(r.x) && is_finite(r.y) && is_finite(r.z)) {
                float *px = a.film + 4 * (size_t)pixel;
                atomicAdd(px + 0, r.x); atomicAdd(px + 1, r.y); atomicAdd(px + 2, r.z); atomicAdd(px + 3, 1.f);
                if (a.film_sq) {
                    float *sq = a.film_sq + 4 * (size_t)pixel;
                    atomicAdd(sq + 0, r.x * r.x); atomicAdd(sq + 1, r.y * r.y); atomicAdd(sq + 2, r.z * r.z);
                }
            }
            finished = 1;
        }
    }
    // warp-aggregated grab of the next work items
    unsigned mask = __ballot_sync(0xffffffffu, need);
    unsigned long long base = 0;
    int lane = LJ_LANE();
    if (mask) {
        int leader = __ffs(mask) - 1;
        if (lane == leader) base = atomicAdd(&a.stats[C_NEXT], (unsigned long long)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
    }
    bool started = false;
    if (need) {
        unsigned long long k = base + __popc(mask & ((1u << lane) - 1));
        uint32_t pixel = 0, sample = 0;
        bool ok = false;
        while (k < a.total_items) {
            ok = item_to_pixel(a, k, pixel, sample);
            if (ok) break;
            k = atomicAdd(&a.stats[C_NEXT], 1ull);  // padded tile pixel outside the film: take another item
        }
        if (ok) {
            PathState s;
            generate_path(sc, a.rp, pixel, sample, s);
            store_state(a.pool, slot, s, true);
            started = true;
        } else {
            a.pool.meta[slot] = mk4(0, 0, 0, 0);  // out of samples: the slot stays empty
        }
    }
    warp_push(a.q_ext[a.cur], &a.qn[Q_EXT0 + a.cur], started, slot);
    warp_add(&a.stats[C_SAMPLES], finished);
}

// K2 / K3: persistent-thread traversal.  SHADOW = false: closest hit of pool.ray -> pool.hit.
// SHADOW = true: any hit of the NEE segment; an unoccluded segment adds its contribution to pool.rad.
constexpr int kRefillThreshold = 20;  // refill idle lanes once fewer than this many lanes still traverse

template <bool SHADOW>
__global__ void __launch_bounds__(128) k_trace(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    const int *queue = SHADOW ? a.q_shadow : a.q_ext[a.cur];
    const unsigned n = a.qn[SHADOW ? Q_SHADOW : Q_EXT0 + a.cur];
    unsigned int *cursor = &a.qn[SHADOW ? CUR_SHADOW : CUR_EXT];
    const int lane = LJ_LANE();
    Trav tr;
    tr.node = kSentinel;
    int slot = -1;
    bool has_ray = false;
    bool drained = false;  // warp-uniform: the queue has no entries left for this warp
    unsigned traced = 0;
    for (;;) {
        // ---- fetch: lanes without a ray take the next queue entries (one atomic per warp)
        unsigned want = drained ? 0u : __ballot_sync(0xffffffffu, !has_ray);
        if (want) {
            int leader = __ffs(want) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(cursor, (unsigned)__popc(want));
            base = __shfl_sync(0xffffffffu, base, leader);
            drained = base + (unsigned)__popc(want) >= n;
            if (!has_ray) {
                unsigned idx = base + __popc(want & ((1u << lane) - 1));
                if (idx < n) {
                    slot = queue[idx];
                    V4 o = a.pool.ray_o[slot];
                    if (SHADOW) {
                        V4 sd = a.pool.sh_d[slot];
                        // The segment starts at the shaded vertex: pool.ray_o if the path continued (the
                        // extension ray starts there too); if it ended, ray_o still holds the previous
                        // origin and the vertex is rebuilt from the old ray and its hit distance.
                        V3 org = xyz(o);
                        if (!(f2u(a.pool.meta[slot].y) & kAlive)) org = org + xyz(a.pool.ray_d[slot]) * a.pool.hit[slot].x;
                        trav_init(tr, org, xyz(sd), sc.shadow_eps, sd.w);
                    } else {
                        V4 d = a.pool.ray_d[slot];
                        trav_init(tr, xyz(o), xyz(d), o.w, d.w);
                    }
                    has_ray = true;
                    traced++;
                }
            }
        }
        if (!__ballot_sync(0xffffffffu, has_ray)) break;  // queue drained and every lane finished
        // ---- traverse (while-while): inner nodes until a leaf, then the leaf
        if (has_ray) {
            while (tr.node != kSentinel) {
                while (tr.node >= 0 && tr.node != kSentinel) trav_inner(sc.nodes2, tr);
                if (tr.node < 0) trav_leaf<SHADOW>(sc.prims, tr);
                if (!drained && __popc(__activemask()) < kRefillThreshold) break;
            }
            if (tr.node == kSentinel) {
                if (SHADOW) {
                    if (tr.hit.prim == kNoHit) {
                        V4 r = a.pool.rad[slot], c = a.pool.sh_c[slot];
                        a.pool.rad[slot] = mk4(r.x + c.x, r.y + c.y, r.z + c.z, r.w);
                    }
                } else {
                    trav_finish_closest(sc.prims, tr);
                    a.pool.hit[slot] = mk4(tr.hit.t, tr.hit.u, tr.hit.v, u2f((uint32_t)tr.hit.prim));
                }
                has_ray = false;
            }
        }
    }
    warp_add(&a.stats[SHADOW ? C_SHADOW : C_CLOSEST], traced);
}

// K4 over the extension queue; compacts survivors / finished paths / shadow rays into the next queues.
__global__ void __launch_bounds__(128) k_shade(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    ShadeCounters cnt = {0, 0, 0, 0};
    bool active = i < a.qn[Q_EXT0 + a.cur];
    int slot = -1;
    bool alive = false, shadow = false;
    if (active) {
        slot = a.q_ext[a.cur][i];
        PathState s;
        load_state(a.pool, slot, s);
        shade_path(sc, a.rp, s, cnt);
        alive = (s.flags & kAlive) != 0;
        shadow = s.sh_tfar >= 0;
        store_state(a.pool, slot, s, alive);
    }
    const int nxt = a.cur ^ 1;
    warp_push(a.q_ext[nxt], &a.qn[Q_EXT0 + nxt], alive, slot);
    warp_push(a.q_free[nxt], &a.qn[Q_FREE0 + nxt], active && !alive, slot);
    warp_push(a.q_shadow, &a.qn[Q_SHADOW], shadow, slot);
    warp_add(&a.stats[C_BOUNCES], cnt.bounces);
}

__global__ void k_init_pool(PathPool pool, int *q_free0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pool.capacity) return;
    pool.meta[i] = mk4(0, 0, 0, 0);
    pool.sh_d[i] = mk4(0, 0, 0, -1.f);
    pool.rad[i] = mk4(0, 0, 0, 1.f);
    q_free0[i] = i;
}

// end of wave: the consumed queues and the cursors go back to zero
__global__ void k_end_wave(unsigned int *qn, int cur) {
    qn[Q_EXT0 + cur] = 0;
    qn[Q_FREE0 + cur] = 0;
    qn[Q_SHADOW] = 0;
    qn[CUR_EXT] = 0;
    qn[CUR_SHADOW] = 0;
}

// film -> caller's w*h*3 buffer (render.cpp:94 divides by spp) and optional variance of the mean
__global__ void k_resolve(const float *film, const float *film_sq, int npix, float inv_n, int normalize, float *out, float *var_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 f = reinterpret_cast<const float4 *>(film)[i];
    float s = normalize ? inv_n : 1.f;
    out[3 * i] = f.x * s; out[3 * i + 1] = f.y * s; out[3 * i + 2] = f.z * s;
    if (var_out && film_sq) {
        float4 q = reinterpret_cast<const float4 *>(film_sq)[i];
        float n = f.w;
        float v[3] = {0, 0, 0};
        if (n > 1) {
            v[0] = fmaxf(q.x - f.x * f.x / n, 0.f) / (n - 1) / n;
            v[1] = fmaxf(q.y - f.y * f.y / n, 0.f) / (n - 1) / n;
            v[2] = fmaxf(q.z - f.z * f.z / n, 0.f) / (n - 1) / n;
        }
        var_out[3 * i] = v[0]; var_out[3 * i + 1] = v[1]; var_out[3 * i + 2] = v[2];
    }
}

static int ensure_pool(lj_scene *s, int capacity) {
    if (s->pool_capacity == capacity && s->pool_block) return LJ_OK;
    if (s->pool_block) { cudaFree(s->pool_block); s->pool_block = nullptr; }
    const int kFields = 9;
    // 9 record arrays + 5 index queues
    LJ_CUDA(cudaMalloc(&s->pool_block, (size_t)capacity * (sizeof(V4) * kFields + sizeof(int) * 5)));
    V4 *base = (V4 *)s->pool_block;
    PathPool &p = s->pool;
    p.ray_o = base + (size_t)capacity * 0; p.ray_d = base + (size_t)capacity * 1; p.hit = base + (size_t)capacity * 2;
    p.thr = base + (size_t)capacity * 3; p.rad = base + (size_t)capacity * 4; p.sh_d = base + (size_t)capacity * 5;
    p.sh_c = base + (size_t)capacity * 6; p.meta = base + (size_t)capacity * 7; p.aux = base + (size_t)capacity * 8;
    p.capacity = capacity;
    s->pool_capacity = capacity;
    return LJ_OK;
}

struct EventPool {
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    cudaEvent_t next() {
        if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
        return ev[used++];
    }
    ~EventPool() { for (auto e : ev) cudaEventDestroy(e); }
};

static int g_num_sms = 0;

static int render_impl(lj_scene *s, const lj_render_opts *opts_in, float *d_out, float *d_var, cudaStream_t stream, lj_stats *stats) {
    lj_render_opts opts;
    memset(&opts, 0, sizeof(opts));
    if (opts_in) opts = *opts_in;
    const DevScene &sc = s->dev;
    if (sc.options.integrator != LJ_INT_PATH) {
        set_error("integrator not supported by this build of the device path");
        return LJ_ERR_UNSUPPORTED;
    }
    int spp = opts.spp > 0 ? opts.spp : sc.options.spp;
    int sb = opts.sample_begin, se = opts.sample_end;
    if (sb == 0 && se == 0) se = spp;
    if (sb < 0 || se > spp || sb >= se) { set_error("bad sample range"); return LJ_ERR_INVALID; }
    int w = sc.camera.width, h = sc.camera.height, npix = w * h;
    int capacity = opts.pool_paths > 0 ? opts.pool_paths : (1 << 22);
    {
        // no point in holding more slots than there are samples
        unsigned long long want = (unsigned long long)npix * (unsigned)(se - sb);
        if (want < (unsigned long long)capacity) capacity = (int)((want + 255) / 256 * 256);
    }
    int r = ensure_pool(s, capacity);
    if (r != LJ_OK) return r;
    if (!s->d_film) LJ_CUDA(cudaMalloc(&s->d_film, (size_t)npix * 16));
    if (d_var && !s->d_film_sq) LJ_CUDA(cudaMalloc(&s->d_film_sq, (size_t)npix * 16));
    unsigned long long *d_stats = nullptr;
    unsigned int *d_qn = nullptr;
    LJ_CUDA(cudaMalloc(&d_stats, sizeof(unsigned long long) * C_COUNT));
    LJ_CUDA(cudaMalloc(&d_qn, sizeof(unsigned int) * Q_COUNT));
    unsigned int *h_qn = nullptr;
    unsigned long long *h_stats = nullptr;
    LJ_CUDA(cudaMallocHost(&h_qn, sizeof(unsigned int) * Q_COUNT));
    LJ_CUDA(cudaMallocHost(&h_stats, sizeof(unsigned long long) * C_COUNT));
    if (g_num_sms == 0) {
#if defined(LJ_HOSTSIM)
        g_num_sms = 1;
#else
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, s->device);
        if (g_num_sms <= 0) g_num_sms = 148;
#endif
    }

    WaveArgs a;
    a.pool = s->pool;
    a.rp.spp_total = (uint32_t)spp;
    a.rp.sample_begin = (uint32_t)sb;
    a.rp.sample_end = (uint32_t)se;
    a.rp.seed = opts.seed ? opts.seed : kPcgDefaultSeed;
    a.rp.width = w;
    a.rp.height = h;
    a.stats = d_stats;
    a.qn = d_qn;
    int *qbase = (int *)((V4 *)s->pool_block + (size_t)capacity * 9);
    a.q_ext[0] = qbase; a.q_ext[1] = qbase + (size_t)capacity; a.q_free[0] = qbase + (size_t)capacity * 2;
    a.q_free[1] = qbase + (size_t)capacity * 3; a.q_shadow = qbase + (size_t)capacity * 4;
    a.film = s->d_film;
    a.film_sq = d_var ? s->d_film_sq : nullptr;
    a.tiles_x = (w + 7) / 8;
    a.tiles_y = (h + 3) / 4;
    a.total_items = (unsigned long long)a.tiles_x * a.tiles_y * 32ull * (unsigned)(se - sb);
    a.cur = 0;

    const int nb256 = (capacity + 255) / 256, nb128 = (capacity + 127) / 128;
    // persistent grid: exactly one wave of resident CTAs (148 SMs x the occupancy of k_trace)
    static int trace_ctas_per_sm = 0;
    if (trace_ctas_per_sm == 0) {
#if defined(LJ_HOSTSIM)
        trace_ctas_per_sm = 1;
#else
        int a0 = 0, a1 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a0, k_trace<false>, 128, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a1, k_trace<true>, 128, 0);
        trace_ctas_per_sm = std::max(1, std::min(a0, a1));
#endif
    }
    const int trace_blocks = g_num_sms * trace_ctas_per_sm;
    EventPool evp;
    std::vector<cudaEvent_t> marks;  // 5 per wave: before regen, extend, shade, shadow, after shadow
    uint64_t launches = 0, waves = 0;

    cudaEvent_t ev_begin = evp.next(), ev_end = evp.next();
    LJ_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(unsigned long long) * C_COUNT, stream));
    LJ_CUDA(cudaMemsetAsync(d_qn, 0, sizeof(unsigned int) * Q_COUNT, stream));
    LJ_CUDA(cudaMemsetAsync(s->d_film, 0, (size_t)npix * 16, stream));
    if (a.film_sq) LJ_CUDA(cudaMemsetAsync(s->d_film_sq, 0, (size_t)npix * 16, stream));
    LJ_CUDA(cudaEventRecord(ev_begin, stream));
    LJ_LAUNCH(k_init_pool, nb256, 256, stream, s->pool, a.q_free[0]);
    {
        unsigned cap = (unsigned)capacity;
        LJ_CUDA(cudaMemcpyAsync(&d_qn[Q_FREE0], &cap, sizeof(unsigned), cudaMemcpyHostToDevice, stream));
    }
    launches++;
    unsigned n_free = (unsigned)capacity, n_ext_prev = 0;
    for (;;) {
        cudaEvent_t e0 = evp.next(), e1 = evp.next(), e2 = evp.next(), e3 = evp.next(), e4 = evp.next();
        LJ_CUDA(cudaEventRecord(e0, stream));
        // grid sizes come from the previous wave's counters (an upper bound: queues only shrink until regen refills)
        int regen_blocks = (int)((n_free + 255) / 256);
        if (regen_blocks > 0) { LJ_LAUNCH(k_regen, regen_blocks, 256, stream, sc, a); launches++; }
        LJ_CUDA(cudaEventRecord(e1, stream));
        LJ_CUDA(cudaMemcpyAsync(h_qn, d_qn, sizeof(unsigned int) * Q_COUNT, cudaMemcpyDeviceToHost, stream));
        LJ_CUDA(cudaStreamSynchronize(stream));
        unsigned n_ext = h_qn[Q_EXT0 + a.cur];
        (void)n_ext_prev;
        if (n_ext == 0) { marks.push_back(e0); marks.push_back(e1); marks.push_back(nullptr); break; }
        int tb = (int)std::min<unsigned>((unsigned)trace_blocks, (n_ext + 127) / 128);
        LJ_LAUNCH(k_trace<false>, tb, 128, stream, sc, a);
        LJ_CUDA(cudaEventRecord(e2, stream));
        LJ_LAUNCH(k_shade, (int)((n_ext + 127) / 128), 128, stream, sc, a);
        LJ_CUDA(cudaEventRecord(e3, stream));
        LJ_LAUNCH(k_trace<true>, tb, 128, stream, sc, a);
        LJ_CUDA(cudaEventRecord(e4, stream));
        LJ_LAUNCH(k_end_wave, 1, 1, stream, d_qn, a.cur);
        launches += 4;
        waves++;
        marks.push_back(e0); marks.push_back(e1); marks.push_back(e2); marks.push_back(e3); marks.push_back(e4);
        // every path of this wave either continues (<= n_ext) or frees its slot (<= n_ext)
        n_free = n_ext;
        n_ext_prev = n_ext;
        a.cur ^= 1;
        if (waves > 1000000) { set_error("wavefront loop did not terminate"); return LJ_ERR_CUDA; }
    }
    (void)nb128;
    LJ_CUDA(cudaEventRecord(ev_end, stream));
    LJ_LAUNCH(k_resolve, (npix + 255) / 256, 256, stream, s->d_film, a.film_sq, npix, 1.f / (float)(se - sb), opts.normalize, d_out, d_var);
    launches++;
    LJ_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(unsigned long long) * C_COUNT, cudaMemcpyDeviceToHost, stream));
    LJ_CUDA(cudaStreamSynchronize(stream));
    LJ_CUDA(cudaGetLastError());
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        stats->render_ms = ms;
        size_t k = 0;
        while (k < marks.size()) {
            if (k + 2 < marks.size() && marks[k + 2] == nullptr) {
                cudaEventElapsedTime(&ms, marks[k], marks[k + 1]); stats->regen_ms += ms;
                break;
            }
            cudaEventElapsedTime(&ms, marks[k], marks[k + 1]); stats->regen_ms += ms;
            cudaEventElapsedTime(&ms, marks[k + 1], marks[k + 2]); stats->extend_ms += ms;
            cudaEventElapsedTime(&ms, marks[k + 2], marks[k + 3]); stats->shade_ms += ms;
            cudaEventElapsedTime(&ms, marks[k + 3], marks[k + 4]); stats->shadow_ms += ms;
            k += 5;
        }
        stats->samples = h_stats[C_SAMPLES];
        stats->closest_rays = h_stats[C_CLOSEST];
        stats->shadow_rays = h_stats[C_SHADOW];
        stats->bounces = h_stats[C_BOUNCES];
        stats->kernel_launches = launches;
        stats->waves = waves;
        stats->extend_launches = stats->shade_launches = stats->shadow_launches = waves;
        stats->regen_launches = waves + 1;
    }
    cudaFree(d_stats);
    cudaFree(d_qn);
    cudaFreeHost(h_qn);
    cudaFreeHost(h_stats);
    return LJ_OK;
}

}  // namespace lj

using namespace lj;

extern "C" int lj_render_device(lj_scene *s, const lj_render_opts *opts, float *d_out_rgb, void *stream, lj_stats *stats) {
    if (!s || !d_out_rgb) { set_error("null argument"); return LJ_ERR_INVALID; }
    if (opts && opts->variance_out) { set_error("variance_out needs lj_render (host buffers)"); return LJ_ERR_INVALID; }
    return render_impl(s, opts, d_out_rgb, nullptr, stream ? (cudaStream_t)stream : s->stream, stats);
}

extern "C" int lj_render(lj_scene *s, const lj_render_opts *opts, float *out_rgb, lj_stats *stats) {
    if (!s || !out_rgb) { set_error("null argument"); return LJ_ERR_INVALID; }
    int npix = s->dev.camera.width * s->dev.camera.height;
    float *d_out = nullptr, *d_var = nullptr;
    LJ_CUDA(cudaMalloc(&d_out, (size_t)npix * 3 * sizeof(float)));
    bool want_var = opts && opts->variance_out;
    if (want_var) LJ_CUDA(cudaMalloc(&d_var, (size_t)npix * 3 * sizeof(float)));
    int r = render_impl(s, opts, d_out, d_var, s->stream, stats);
    if (r == LJ_OK) {
        cudaError_t e = cudaMemcpy(out_rgb, d_out, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && want_var) e = cudaMemcpy(opts->variance_out, d_var, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) r = cuda_fail(e, "framebuffer download");
    }
    cudaFree(d_out);
    if (d_var) cudaFree(d_var);
    return r;
}
