// wavefront.cu -- the wavefront path tracer: regenerate / extend / shade / shadow kernels over a
// pool of path records resident in HBM, and the render entry points of the C ABI.
// Replaces the reference's tile thread pool (render.cpp:71-152, parallel.cpp) and the per-sample
// control flow of path_tracing.h.  Stage kernels:
//   k_regen      K1+K7  flush finished paths to the film, start new camera paths in free slots
//   k_trace_q<0> K2     persistent, queue form: closest-hit traversal for every live path (volpath: k_trace<0>)
//   k_shade      K4     emission+MIS, Russian roulette, NEE sample, BSDF sample (lj_path.h); k_flight + k_shade_vol for volpath
//   k_trace_q<1> K3     persistent: any-hit traversal of the NEE shadow rays (volpath: the NEE walk, k_trace<2> or staged)
// Path records are addressed by slot = thread index in regen and in the first shade pass (fully coalesced 16-byte
// records); the pool is kept dense by regenerating finished paths in place.  A general index queue in front of k_shade
// made it 2x slower through uncoalesced record access (profiles/r01_b_*), so queues are used only where the code behind
// them is expensive and sparse: the Disney pass of k_shade and the surface pass of k_shade_vol gather their records
// through pool.class_queue (profiles/r02l_* -> r02m_*).
#include "scene.cuh"
#include "lj_volpath.h"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if !defined(LJ_HOSTSIM)
#include <dlfcn.h>
#endif

namespace lj {

enum { C_SAMPLES = 0, C_CLOSEST, C_SHADOW, C_BOUNCES, C_ACTIVE, C_NEXT, C_NODE_STEPS, C_PRIM_TESTS, C_NODE_PASSES, C_PRIM_PASSES, C_COUNT };
// Statistics counters every warp of a full-pool kernel adds to are striped over kStripes addresses (folded on the
// host): 131k same-address reductions per launch serialise in one L2 slice otherwise.
constexpr int kStripes = 64;
constexpr int C_BOUNCES_STRIPED = C_COUNT;               // kStripes entries
constexpr int C_TOTAL = C_COUNT + kStripes;

struct WaveArgs {
    PathPool pool;
    RenderParams rp;
    unsigned long long *counters;  // C_COUNT
    float *film;                   // w*h*4: sum rgb, n
    float *film_sq;                // w*h*4: sum of squares rgb (may be null)
    unsigned int *cursors;         // fetch cursors of the persistent kernels: [0] extend, [1] shadow / walk, [2] flight
    unsigned long long total_items;  // padded pixels * samples in this call
    int tiles_x, tiles_y;
    int prim_min_lanes, refill_threshold;  // scheduling policy of k_trace
    int chunk, shadow_chunk;               // slots a warp takes from the cursor at a time
    int track_refill;                      // k_flight / walk: refill once fewer lanes than this are tracking
    int trav_min;                          // k_trace<3>: lanes that must wait to traverse before a traversal pass pre-empts tracking
    // k_trace_q: per-warp scratch stacks (qdepth entries x kQRays rays x 8 B per warp) and its refill threshold
    U2 *qstack;
    int qdepth, q_refill, q_chunk;
    uint32_t one_bits;  // 0x3f800000, see byte_m (lj_bvh.h)
    // image-space split (lj_render_opts.split == LJ_SPLIT_TILES): this call renders the 8x4 pixel tiles t with
    // t % tile_stride == tile_offset; a sample split leaves tile_stride = 1
    int tile_stride, tile_offset, tiles_local;
    int walk_whole_groups;  // k_trace<2|3>: test a lane's whole primitive group in one pass (trees of a few nodes)
    int closest_counter;    // k_trace_q<0>: the counter its rays are added to (C_CLOSEST; C_SHADOW for walk segments)
};

__device__ __forceinline__ void warp_add(unsigned long long *ctr, unsigned v) {
    unsigned s = __reduce_add_sync(0xffffffffu, v);
    if (LJ_LANE() == 0 && s) atomicAdd(ctr, (unsigned long long)s);
}

// One 16-byte reduction per finished sample (red.global.add.v4.f32, sm_90+) instead of four scalar atomics.
__device__ __forceinline__ void film_add(float *px, float r, float g, float b, float n) {
#if defined(__CUDA_ARCH__)
    atomicAdd(reinterpret_cast<float4 *>(px), make_float4(r, g, b, n));
#else
    px[0] += r; px[1] += g; px[2] += b; px[3] += n;
#endif
}

// K1 + K7.  Work item k -> (sample, 8x4 pixel tile, lane) so one warp starts 32 neighbouring pixels.
__global__ void __launch_bounds__(256) k_regen(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool in_range = i < a.pool.capacity;
    uint32_t flags = 0, pixel = 0;
    V4 meta = mk4(0, 0, 0, 0);
    if (in_range) {
        meta = a.pool.meta[i];
        flags = f2u(meta.y) & 0xffff0000u;
        pixel = f2u(meta.x);
    }
    bool alive = in_range && (flags & kAlive);
    bool need = in_range && !alive;
    unsigned finished = 0;
    if (need && (flags & kOccupied)) {
        // K7: accumulate the finished sample (render.cpp:92) -- non-finite samples are dropped
        V4 r = a.pool.rad[i];
        if (is_finite(r.x) && is_finite(r.y) && is_finite(r.z)) {
            film_add(a.film + 4 * (size_t)pixel, r.x, r.y, r.z, 1.f);
            if (a.film_sq) film_add(a.film_sq + 4 * (size_t)pixel, r.x * r.x, r.y * r.y, r.z * r.z, 0.f);
        }
        finished = 1;
    }
    unsigned mask = __ballot_sync(0xffffffffu, need);
    const int lane = LJ_LANE();
#if defined(LJ_HOSTSIM)  // serial stand-in: "warps" of one thread, no block to cooperate with
    unsigned long long base = need ? atomicAdd(&a.counters[C_NEXT], 1ull) : 0ull;
#else
    // block-aggregated grab of the next work items: one returning atomic per 256 slots (a per-warp atomic on the
    // one cursor costs ~0.4 ms per wave in same-address serialisation)
    __shared__ unsigned s_warp_need[8];
    __shared__ unsigned long long s_base;
    __shared__ unsigned s_cnt[2];
    const int warp = threadIdx.x >> 5;
    if (lane == 0) s_warp_need[warp] = __popc(mask);
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int k = 0; k < 8; k++) { unsigned c = s_warp_need[k]; s_warp_need[k] = tot; tot += c; }
        s_base = tot ? atomicAdd(&a.counters[C_NEXT], (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    unsigned long long base = s_base + s_warp_need[warp];
#endif
    unsigned started = 0;
    if (need) {
        unsigned long long k = base + __popc(mask & ((1u << lane) - 1));
        bool ok = k < a.total_items;
        uint32_t x = 0, y = 0, sample = 0;
        if (ok) {
            unsigned long long per_sample = (unsigned long long)a.tiles_local * 32ull;
            sample = a.rp.sample_begin + (uint32_t)(k / per_sample);
            uint32_t p = (uint32_t)(k % per_sample);
            uint32_t tile = (p >> 5) * (uint32_t)a.tile_stride + (uint32_t)a.tile_offset, l = p & 31;
            x = (tile % a.tiles_x) * 8 + (l & 7);
            y = (tile / a.tiles_x) * 4 + (l >> 3);
            ok = x < (uint32_t)a.rp.width && y < (uint32_t)a.rp.height;
        }
        if (ok) {
            PathState s;
            generate_path(sc, a.rp, y * a.rp.width + x, sample, s);
            if (a.pool.vol0) store_state_vol(a.pool, i, s, true); else store_state(a.pool, i, s, true);
            a.pool.hit[i] = mk4(0, 0, 0, u2f((uint32_t)kNoHit));
            started = 1;
        } else if (flags) {
            a.pool.meta[i] = mk4(meta.x, u2f(0u), meta.z, meta.w);  // slot is now empty
            a.pool.sh_d[i] = mk4(0, 0, 0, -1.f);
        }
    }
#if defined(LJ_HOSTSIM)
    warp_add(&a.counters[C_SAMPLES], finished);
    warp_add(&a.counters[C_ACTIVE], started + (alive ? 1u : 0u));
#else
    // statistics: block totals through shared memory, one reduction per block and counter
    unsigned f = __reduce_add_sync(0xffffffffu, finished), l = __reduce_add_sync(0xffffffffu, started + (alive ? 1u : 0u));
    if (lane == 0) { if (f) atomicAdd(&s_cnt[0], f); if (l) atomicAdd(&s_cnt[1], l); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&a.counters[C_SAMPLES], (unsigned long long)s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&a.counters[C_ACTIVE], (unsigned long long)s_cnt[1]);
    }
#endif
}

// K2 / K3: persistent-thread traversal over the path pool.  Each warp takes 32 consecutive slots at a
// time from a global cursor (one atomic per warp); lanes whose slot carries no ray of this kind, or
// whose ray is finished, are refilled from the cursor once fewer than kRefillThreshold lanes of the
// warp are still traversing.  The work loop is warp-synchronous (full-mask ballots at the top of every
// iteration), so the lanes reconverge once per step: one iteration = one wide-node step for the lanes
// that need one, then the primitive tests the step produced.  When only a few lanes hold primitives
// they are postponed (pushed as a group) so the tests run with more of the warp active.
// SHADOW = false: closest hit of pool.ray -> pool.hit.  SHADOW = true: any hit of the NEE segment; an
// unoccluded segment adds its contribution to pool.rad.
#ifndef LJ_SHADOW_MIN_BLOCKS
#define LJ_SHADOW_MIN_BLOCKS 8
#endif
#ifndef LJ_WALK_MIN_BLOCKS
#define LJ_WALK_MIN_BLOCKS 5
#endif
constexpr int kRefillThreshold = 24;  // defaults; LJ_REFILL / LJ_PRIM_MIN_LANES override them for tuning runs
constexpr int kPrimMinLanes = 10;

// MODE 2 (volpath NEE walk, homework2.tex:459-510): the lane keeps its slot across the segments of one walk --
// each segment is a closest-hit traversal, followed by ratio tracking over it and the index-matched / opaque test
// (nee_walk_step); the walk's transmittance products stay in registers until the segment chain ends.
template <int MODE>
__global__ void __launch_bounds__(128, MODE >= 2 ? LJ_WALK_MIN_BLOCKS : (MODE == 1 ? LJ_SHADOW_MIN_BLOCKS : 8)) k_trace(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    constexpr bool SHADOW = MODE == 1;
    constexpr bool WALK = MODE >= 2;   // 2: whole ratio-tracking loop when a segment ends (homogeneous media: 1-2 collisions)
    constexpr bool STEP = MODE == 3;   // 3: ratio tracking one collision per pass (heterogeneous media: ~100 per segment)
    NeeWalk wk;
    TrackState ts;
    float seg_tfar = 0, next_t = 0;
    bool tracking = false;  // WALK: the lane is ratio tracking over the segment it just traversed
    const unsigned n = (unsigned)a.pool.capacity;
    unsigned int *cursor = &a.cursors[MODE == 0 ? 0 : 1];
    const int lane = LJ_LANE();
    TravL tr;
    trav_terminate(tr);
    int slot = -1;
    bool has_ray = false;
    bool drained = false;  // warp-uniform: the cursor ran past the pool and the warp's own run is used up
    bool global_out = false;
    unsigned chunk_next = 0, chunk_end = 0;
    unsigned cur_mask = 0, cur_word = 0;  // MODE >= 1: unread bits of the warp's current sh_mask word
    const unsigned chunk = (unsigned)(MODE == 1 ? a.shadow_chunk : a.chunk);  // walks: small runs, their cost per slot varies a lot
    unsigned traced = 0, node_steps = 0, prim_tests = 0, node_passes = 0, prim_passes = 0;
    for (;;) {
        // ---- fetch: lanes without a ray take the next slots.  The warp owns a private run of slots [chunk_next,
        // chunk_end) and goes to the global cursor only when that runs out: one same-address atomic per a.chunk
        // slots instead of one per refill.
        if (MODE == 0) {
            unsigned want = drained ? 0u : __ballot_sync(0xffffffffu, !has_ray);
            if (want) {
                const unsigned cnt = (unsigned)__popc(want);
                const unsigned lim = chunk_end < n ? chunk_end : n;
                const unsigned left = chunk_next < lim ? lim - chunk_next : 0u;
                unsigned nb = 0;
                bool fresh = false;
                if (cnt > left && !global_out) {
                    if (lane == 0) nb = atomicAdd(cursor, chunk);
                    nb = __shfl_sync(0xffffffffu, nb, 0);
                    if (nb >= n) global_out = true; else fresh = true;
                }
                const unsigned rank = (unsigned)__popc(want & ((1u << lane) - 1));
                const unsigned idx = rank < left ? chunk_next + rank : (fresh ? nb + (rank - left) : 0xffffffffu);
                if (fresh) { chunk_next = nb + (cnt - left); chunk_end = nb + chunk; }
                else chunk_next += cnt < left ? cnt : left;
                drained = global_out && chunk_next >= (chunk_end < n ? chunk_end : n);
                if (!has_ray && idx < n && (f2u(a.pool.meta[idx].y) & kAlive)) {
                    slot = (int)idx;
                    V4 o = a.pool.ray_o[slot], d = a.pool.ray_d[slot];
                    trav_init(tr, xyz(o), xyz(d), o.w, d.w);
                    has_ray = true;
                    traced++;
                }
            }
        } else {
            // Shadow rays / walks exist for a fraction of the slots only: the shade kernel left one bit per slot in
            // pool.sh_mask, and lanes are handed the set bits of the warp's current word (chunk_next counts WORDS here),
            // so no lane ever loads the record of a slot without a ray.
            while (!drained) {
                const unsigned want = __ballot_sync(0xffffffffu, !has_ray);
                if (!want) break;
                if (cur_mask == 0) {
                    if (chunk_next >= chunk_end) {
                        unsigned nb = 0;
                        if (lane == 0) nb = atomicAdd(cursor, chunk);
                        nb = __shfl_sync(0xffffffffu, nb, 0);
                        if (nb >= n) { drained = true; break; }
                        chunk_next = nb / LJ_WARP_WIDTH;
                        chunk_end = ((nb + chunk < n ? nb + chunk : n) + LJ_WARP_WIDTH - 1) / LJ_WARP_WIDTH;
                    }
                    cur_word = chunk_next++;
                    cur_mask = a.pool.sh_mask[cur_word];
                    continue;
                }
                const unsigned cnt = (unsigned)__popc(want), avail = (unsigned)__popc(cur_mask);
                const unsigned take = cnt < avail ? cnt : avail;
                const unsigned rank = (unsigned)__popc(want & ((1u << lane) - 1));
                if (!has_ray && rank < take) {
                    slot = (int)(cur_word * LJ_WARP_WIDTH + __fns(cur_mask, 0, (int)rank + 1));
                    V4 sd = a.pool.sh_d[slot];
                    if (WALK) {
                        V4 so = a.pool.sh_o[slot], pl = a.pool.sh_pl[slot], cc = a.pool.sh_c[slot];
                        uint32_t mb = f2u(so.w);
                        wk.pc = xyz(so); wk.dir = xyz(sd); wk.pl = xyz(pl);
                        wk.T_light = mk3(1); wk.p_nee = mk3(1); wk.p_dir = mk3(1);
                        wk.c = xyz(cc); wk.pdf_nee = cc.w; wk.pdf_dir = sd.w;
                        wk.medium = (int)(mb & 0xffffu) - 1;
                        wk.budget = mb >> 16;
                        wk.shadow_bounces = 0;
                        // the walk's own stream: a function of the PATH (pixel, sample) and of one draw of the path's
                        // stream (pl.w), never of the slot the path happens to sit in -- renders are reproducible
                        // whatever the pool size or the split across GPUs
                        {
                            const uint64_t path_id = (uint64_t)f2u(a.pool.meta[slot].x) * a.rp.spp_total + f2u(a.pool.aux[slot].z);
                            wk.rng = walk_rng(path_id, f2u(pl.w), a.rp.seed);
                        }
                        float tn;
                        nee_walk_segment(sc, wk, tn, seg_tfar);
                        trav_init(tr, wk.pc, wk.dir, tn, seg_tfar);
                        tracking = false;
                    } else {
                        // The segment starts at the shaded vertex: pool.ray_o if the path continued (the
                        // extension ray starts there too); if it ended, ray_o still holds the previous
                        // origin and the vertex is rebuilt from the old ray and its hit distance.
                        V4 o = a.pool.ray_o[slot];
                        V3 org = xyz(o);
                        if (!(f2u(a.pool.meta[slot].y) & kAlive)) org = org + xyz(a.pool.ray_d[slot]) * a.pool.hit[slot].x;
                        trav_init(tr, org, xyz(sd), sc.shadow_eps, sd.w);
                    }
                    has_ray = true;
                    traced++;
                }
                const unsigned last = __fns(cur_mask, 0, (int)take);  // position of the last bit handed out
                cur_mask &= ~((2u << last) - 1u);
                if (take == cnt) break;
            }
        }
        if (!__ballot_sync(0xffffffffu, has_ray)) {
            if (drained) break;  // pool exhausted and every lane finished
            continue;            // this batch of slots held no rays: fetch again
        }
        // ---- work loop.  Every pass the warp votes for one of two steps: a primitive step (each lane holding
        // primitives tests ONE of them) when at least prim_min_lanes lanes hold some or nobody can descend, else a
        // node step, in which lanes holding primitives push them for later and descend too if they can.
        // (a lane without a ray keeps both groups empty, so "has work" is just "not done")
        for (bool first = true;; first = false) {
            const bool has_p = tr.Gt.y != 0;
            const bool work = !trav_done(tr);
            const unsigned wm = __ballot_sync(0xffffffffu, work);
            // WALK: a lane is also busy while it tracks, or holds a traversed segment it has not looked at yet.  Each
            // pass runs ONE of the two phases, the one more lanes are waiting for: traversal steps, or walk steps
            // (segment end -> tracking step -> surface test -> next segment).
            const unsigned km = STEP ? __ballot_sync(0xffffffffu, tracking || (has_ray && !work)) : 0u;
            const unsigned busy = wm | km;
            // (measured on hetvol: letting as few as 4 waiting lanes pre-empt the tracking phase is best -- segments are
            //  short to traverse, and lanes that get through traversal quickly join the ~100-collision tracking runs)
            const bool trav_phase = !STEP || km == 0 || __popc(wm) >= a.trav_min;
            if (busy == 0) break;
            // refill once too few lanes are still busy (never before the pass after a fetch made progress)
            if (!first && !drained && __popc(busy) < (STEP ? a.track_refill : a.refill_threshold)) break;
            if (wm && trav_phase) {
                const unsigned pm = __ballot_sync(0xffffffffu, has_p);
                if (__popc(pm) >= a.prim_min_lanes || pm == wm) {
                    if (lane == 0) prim_passes++;
                    if (has_p) {
                        prim_tests++;
                        if (trav_prim<SHADOW>(sc.prims, tr)) trav_terminate(tr);  // any-hit ends at the first hit
                        if (WALK && a.walk_whole_groups) {
                            // walk kernels, scenes whose primitives sit in a handful of groups (hetvol: 15 primitives under
                            // the root): the whole group in this pass.  A pass is shared with lanes that wait to track,
                            // and one pass per primitive kept the warp at 6 lanes (hetvol 154 -> 166 Msamples/s); with a
                            // real hierarchy the groups are short and waiting for the longest one costs more than it saves
                            // (vol_cbox_teapot: walk stage +18 %), so render_impl sets this for tiny trees only.
                            while (tr.Gt.y != 0) { prim_tests++; trav_prim<SHADOW>(sc.prims, tr); }
                        }
                    }
                } else if ((lane == 0 ? (void)node_passes++ : (void)0), work) {
                    bool descend = !has_p;
                    if (has_p && tr.G.y != 0 && tr.sp < kTriPostponeMax) {
                        tr.stack.put(tr.sp++, tr.Gt);  // postpone these primitives, keep descending
                        tr.Gt.y = 0;
                        descend = true;
                    }
                    if (descend) {
                        node_steps++;
                        trav_node(sc.nodes8, tr, tr.stack, a.one_bits);
                    }
                }
                trav_next_group(tr, tr.stack);
            }
            if (STEP && !trav_phase) {
                // A lane that finished traversing its segment starts tracking over it (or goes straight to the surface
                // test) and takes its first collision step in the same pass; a tracking lane takes ONE collision step
                // per pass (homework2.tex:771-810).  In homogeneous media a segment is over after one or two steps.
                bool decide = false;
                if (has_ray && !tracking && trav_done(tr)) {
                    trav_finish_closest(sc.prims, tr);
                    next_t = nee_walk_next_t(wk, tr.hit);
                    if (wk.medium >= 0 && track_begin(sc.media[wk.medium], wk.pc, wk.dir, seg_tfar, wk.rng, ts)) tracking = true;
                    else decide = true;
                }
                if (tracking) {
                    if (ratio_step(sc.media[wk.medium], wk.pc, wk.dir, next_t, sc.options.max_null_collisions, wk.rng, ts,
                                   wk.T_light, wk.p_nee, wk.p_dir) != kTrackContinue) { tracking = false; decide = true; }
                }
                if (decide) {
                    V3 contrib;
                    if (nee_walk_decide(sc, wk, tr.hit, contrib)) {
                        if (max3(contrib) > 0 || min3(contrib) < 0 || contrib.x != contrib.x || contrib.y != contrib.y || contrib.z != contrib.z) {
                            V4 r = a.pool.rad[slot];
                            a.pool.rad[slot] = mk4(r.x + contrib.x, r.y + contrib.y, r.z + contrib.z, r.w);
                        }
                        has_ray = false;
                        trav_terminate(tr);
                    } else {  // next segment of the same walk
                        float tn;
                        nee_walk_segment(sc, wk, tn, seg_tfar);
                        trav_init(tr, wk.pc, wk.dir, tn, seg_tfar);
                        traced++;
                    }
                }
            }
        }
        if (!STEP && has_ray && trav_done(tr)) {  // finished (possibly in an earlier pass of this loop)
            if (WALK) {
                trav_finish_closest(sc.prims, tr);
                next_t = nee_walk_next_t(wk, tr.hit);
                if (wk.medium >= 0) {
                    ratio_track(sc.media[wk.medium], wk.pc, wk.dir, seg_tfar, next_t, sc.options.max_null_collisions, wk.rng,
                                wk.T_light, wk.p_nee, wk.p_dir);
                }
                V3 contrib;
                if (nee_walk_decide(sc, wk, tr.hit, contrib)) {
                    if (max3(contrib) > 0 || min3(contrib) < 0 || contrib.x != contrib.x || contrib.y != contrib.y || contrib.z != contrib.z) {
                        V4 r = a.pool.rad[slot];
                        a.pool.rad[slot] = mk4(r.x + contrib.x, r.y + contrib.y, r.z + contrib.z, r.w);
                    }
                    has_ray = false;
                } else {  // next segment of the same walk
                    float tn;
                    nee_walk_segment(sc, wk, tn, seg_tfar);
                    trav_init(tr, wk.pc, wk.dir, tn, seg_tfar);
                    traced++;
                }
            } else if (SHADOW) {
                if (tr.hit.prim == kNoHit) {
                    V4 r = a.pool.rad[slot], c = a.pool.sh_c[slot];
                    a.pool.rad[slot] = mk4(r.x + c.x, r.y + c.y, r.z + c.z, r.w);
                }
            } else {
                trav_finish_closest(sc.prims, tr);
                a.pool.hit[slot] = mk4(tr.hit.t, tr.hit.u, tr.hit.v, u2f((uint32_t)tr.hit.prim));
            }
            if (!WALK) has_ray = false;
        }
    }
    warp_add(&a.counters[MODE == 0 ? C_CLOSEST : C_SHADOW], traced);
    warp_add(&a.counters[C_NODE_STEPS], node_steps);
    warp_add(&a.counters[C_PRIM_TESTS], prim_tests);
    warp_add(&a.counters[C_NODE_PASSES], node_passes);
    warp_add(&a.counters[C_PRIM_PASSES], prim_passes);
}

// K2 / K3, queue form (the kernels lj_render launches for the path integrator).  k_trace above runs one ray per
// lane and votes per pass for a node step or a primitive step; measured on sponza only 18 of 32 lanes do the voted
// step (profiles/r01w_sponza_ncu.txt) -- the others hold a ray that wants the other kind of step, or none.  Here a
// warp owns kQRays = 64 ray slots whose traversal state lives in SHARED memory (structure of arrays, 4.2 KB per warp),
// and every pass compacts the slots that want the chosen kind of step onto the lanes with two ballots and a rank
// (the shared-memory / warp-ballot compaction north_star asks for): a pass runs with 32 active lanes whenever 32 of
// the 64 rays want the same step.  Three kinds of step: NODE (one wide-node test), PRIM (one primitive test), FIN
// (write the result of a finished ray: barycentrics + fp64 t for closest hits, the NEE contribution for shadow
// rays).  Group stacks sit in a per-warp global scratch area addressed [entry][ray] (same cache path as local
// memory, but reachable from whichever lane processes the ray).  Same results as trace8 / k_trace: the step
// functions are the same (lj_bvh.h) and the equal-t tie policy makes the winner independent of visiting order.
#ifndef LJ_Q_PRIMS_PER_PASS
#define LJ_Q_PRIMS_PER_PASS 2
#endif
constexpr int kQPrimsPerPass = LJ_Q_PRIMS_PER_PASS;
#ifndef LJ_Q_NODES_PER_PASS
#define LJ_Q_NODES_PER_PASS 1
#endif
constexpr int kQNodesPerPass = LJ_Q_NODES_PER_PASS;
constexpr int kQWarps = 4;
constexpr int kQRays = 2 * LJ_WARP_WIDTH;
enum { Q_EMPTY = 0, Q_NODE = 1, Q_PRIM = 2, Q_FIN = 3 };

struct QRays {
    float ox[kQRays], oy[kQRays], oz[kQRays], dx[kQRays], dy[kQRays], dz[kQRays], tnear[kQRays], tfar[kQRays];
    int prim[kQRays], slot[kQRays];
    uint32_t Gx[kQRays], Gy[kQRays], Tx[kQRays], Ty[kQRays], sp[kQRays], status[kQRays];
    uint32_t list[LJ_WARP_WIDTH];
};

LJ_HD uint32_t q_status_of(const Trav &tr) { return tr.Gt.y ? (uint32_t)Q_PRIM : (tr.G.y ? (uint32_t)Q_NODE : (uint32_t)Q_FIN); }

// a new ray into slot r of the warp's table
LJ_HD void q_put_ray(QRays &q, int r, int slot, V3 o, V3 d, float tnear, float tfar) {
    q.ox[r] = o.x; q.oy[r] = o.y; q.oz[r] = o.z; q.dx[r] = d.x; q.dy[r] = d.y; q.dz[r] = d.z;
    q.tnear[r] = tnear; q.tfar[r] = tfar;
    q.prim[r] = kNoHit; q.slot[r] = slot;
    q.Gx[r] = 0; q.Gy[r] = (tnear <= tfar) ? 0x80000000u : 0u;  // root group, as trav_init
    q.Tx[r] = 0; q.Ty[r] = 0; q.sp[r] = 0;
    q.status[r] = (tnear <= tfar) ? (uint32_t)Q_NODE : (uint32_t)Q_FIN;
}

template <int MODE>
__global__ void __launch_bounds__(kQWarps * LJ_WARP_WIDTH, 8) k_trace_q(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    constexpr bool SHADOW = MODE == 1;
    constexpr int W = LJ_WARP_WIDTH;
    __shared__ QRays s_rays[kQWarps];
    const int lane = LJ_LANE();
#if defined(LJ_HOSTSIM)
    const int warp = 0;
    const size_t warp_global = 0;
#else
    const int warp = (int)(threadIdx.x / W);
    const size_t warp_global = (size_t)blockIdx.x * kQWarps + warp;
#endif
    QRays &q = s_rays[warp];
    U2 *const stack_base = a.qstack + warp_global * (size_t)a.qdepth * kQRays;
    const unsigned n = (unsigned)a.pool.capacity;
    unsigned int *cursor = &a.cursors[MODE];
    const unsigned chunk = (unsigned)a.q_chunk;
    const unsigned lt = (1u << lane) - 1u;
    unsigned chunk_next = 0, chunk_end = 0;   // MODE 0: slots; MODE 1: sh_mask words
    unsigned cur_mask = 0, cur_word = 0;
    bool drained = false, global_out = false;
    unsigned traced = 0, node_steps = 0, prim_tests = 0, node_passes = 0, prim_passes = 0;
    q.status[lane] = Q_EMPTY;
    q.status[lane + W] = Q_EMPTY;
    __syncwarp();
    int cF = 0;  // warp-uniform: slots waiting for their result to be written
    for (unsigned pass = 0;; pass++) {
        uint32_t s0 = q.status[lane], s1 = q.status[lane + W];
        const unsigned mN0 = __ballot_sync(0xffffffffu, s0 == Q_NODE), mN1 = __ballot_sync(0xffffffffu, s1 == Q_NODE);
        const unsigned mP0 = __ballot_sync(0xffffffffu, s0 == Q_PRIM), mP1 = __ballot_sync(0xffffffffu, s1 == Q_PRIM);
        const int cN = __popc(mN0) + __popc(mN1), cP = __popc(mP0) + __popc(mP1);
        const int live = cN + cP;
        const bool want_refill = !drained && live < a.q_refill;
        int kind;
        if (cF >= W || (cF > 0 && (want_refill || live == 0))) {
            kind = Q_FIN;
        } else if (want_refill) {
            // ---- fetch new rays into the empty slots (both halves of the table), from the warp's private run of the
            // pool and, when that runs out, the global cursor: one same-address atomic per q_chunk slots
            if (MODE == 0) {
                const unsigned e0 = __ballot_sync(0xffffffffu, s0 == Q_EMPTY), e1 = __ballot_sync(0xffffffffu, s1 == Q_EMPTY);
                const unsigned c0 = (unsigned)__popc(e0), cnt = c0 + (unsigned)__popc(e1);
                const unsigned lim = chunk_end < n ? chunk_end : n;
                const unsigned left = chunk_next < lim ? lim - chunk_next : 0u;
                unsigned nb = 0;
                bool fresh = false;
                if (cnt > left && !global_out) {
                    if (lane == 0) nb = atomicAdd(cursor, chunk);
                    nb = __shfl_sync(0xffffffffu, nb, 0);
                    if (nb >= n) global_out = true; else fresh = true;
                }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int half = 0; half < 2; half++) {
                    const bool mine = (half ? s1 : s0) == Q_EMPTY;
                    const unsigned rank = half ? c0 + (unsigned)__popc(e1 & lt) : (unsigned)__popc(e0 & lt);
                    const unsigned idx = rank < left ? chunk_next + rank : (fresh ? nb + (rank - left) : 0xffffffffu);
                    if (mine && idx < n && (f2u(a.pool.meta[idx].y) & kAlive)) {
                        V4 o = a.pool.ray_o[idx], d = a.pool.ray_d[idx];
                        q_put_ray(q, lane + half * W, (int)idx, xyz(o), xyz(d), o.w, d.w);
                        traced++;
                    }
                }
                if (fresh) { chunk_next = nb + (cnt - left); chunk_end = nb + chunk; }
                else chunk_next += cnt < left ? cnt : left;
                drained = global_out && chunk_next >= (chunk_end < n ? chunk_end : n);
            } else {
                // shadow rays exist for a fraction of the slots: hand out the set bits of pool.sh_mask (one word per
                // warp of the shade kernel), so no record of a slot without a ray is ever loaded
                while (!drained) {
                    const unsigned e0 = __ballot_sync(0xffffffffu, s0 == Q_EMPTY), e1 = __ballot_sync(0xffffffffu, s1 == Q_EMPTY);
                    const unsigned c0 = (unsigned)__popc(e0), cnt = c0 + (unsigned)__popc(e1);
                    if (!cnt) break;
                    if (cur_mask == 0) {
                        if (chunk_next >= chunk_end) {
                            unsigned nb = 0;
                            if (lane == 0) nb = atomicAdd(cursor, chunk);
                            nb = __shfl_sync(0xffffffffu, nb, 0);
                            if (nb >= n) { drained = true; break; }
                            chunk_next = nb / W;
                            chunk_end = ((nb + chunk < n ? nb + chunk : n) + W - 1) / W;
                        }
                        cur_word = chunk_next++;
                        cur_mask = a.pool.sh_mask[cur_word];
                        continue;
                    }
                    const unsigned avail = (unsigned)__popc(cur_mask), take = cnt < avail ? cnt : avail;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                    for (int half = 0; half < 2; half++) {
                        const bool mine = (half ? s1 : s0) == Q_EMPTY;
                        const unsigned rank = half ? c0 + (unsigned)__popc(e1 & lt) : (unsigned)__popc(e0 & lt);
                        if (mine && rank < take) {
                            const int slot = (int)(cur_word * W + __fns(cur_mask, 0, (int)rank + 1));
                            V4 sd = a.pool.sh_d[slot];
                            // the segment starts at the shaded vertex: pool.ray_o if the path continued; if it ended,
                            // ray_o still holds the previous origin and the vertex is rebuilt from the old ray
                            V3 org = xyz(a.pool.ray_o[slot]);
                            if (!(f2u(a.pool.meta[slot].y) & kAlive)) org = org + xyz(a.pool.ray_d[slot]) * a.pool.hit[slot].x;
                            q_put_ray(q, lane + half * W, slot, org, xyz(sd), sc.shadow_eps, sd.w);
                            if (half) s1 = Q_NODE; else s0 = Q_NODE;  // (any non-empty value: only emptiness is re-read here)
                            traced++;
                        }
                    }
                    const unsigned last = __fns(cur_mask, 0, (int)take);  // position of the last bit handed out
                    cur_mask &= ~((2u << last) - 1u);
                    if (take == cnt) break;
                }
            }
            __syncwarp();
            cF = __popc(__ballot_sync(0xffffffffu, q.status[lane] == Q_FIN)) + __popc(__ballot_sync(0xffffffffu, q.status[lane + W] == Q_FIN));
            continue;
        } else if (live == 0) {
            break;  // pool exhausted, nothing in flight, nothing left to write
        } else {
            // the kind that fills more lanes; primitives first on a tie (they shorten the ray)
            const int fN = cN < W ? cN : W, fP = cP < W ? cP : W;
            kind = fP >= fN ? Q_PRIM : Q_NODE;
        }
        // ---- compaction: the first W slots that want this kind of step, the two halves taking turns to go first.  Slots
        // r and r + W share a shared-memory bank, so of the second half the lanes whose first-half slot is NOT selected
        // go first: the 32 selected slots then sit in 32 different banks whenever that is possible.
        {
            unsigned m0, m1;
            if (kind == Q_FIN) { m0 = __ballot_sync(0xffffffffu, s0 == Q_FIN); m1 = __ballot_sync(0xffffffffu, s1 == Q_FIN); }
            else { m0 = kind == Q_NODE ? mN0 : mP0; m1 = kind == Q_NODE ? mN1 : mP1; }
            const bool flip = (pass & 1u) != 0;
            const unsigned first = flip ? m1 : m0, second = flip ? m0 : m1;
            const unsigned pref = second & ~first, rest = second & first;
            const int cf = __popc(first), cp = __popc(pref);
            if ((first >> lane) & 1u) q.list[__popc(first & lt)] = (uint32_t)(lane + (flip ? W : 0));
            if ((second >> lane) & 1u) {
                const int r = cf + (((pref >> lane) & 1u) ? __popc(pref & lt) : cp + __popc(rest & lt));
                if (r < W) q.list[r] = (uint32_t)(lane + (flip ? 0 : W));
            }
            __syncwarp();
            const int nsel = cf + __popc(second) < W ? cf + __popc(second) : W;
            node_passes += kind == Q_NODE;
            prim_passes += kind == Q_PRIM;
            bool finished = false, more = false;
            if (lane < nsel) {
                // (each kind of step reads and writes only the fields it needs: the kernel is bound by the shared-memory /
                //  L1 data pipe, profiles/r02b_sponza_ncu.txt)
                const int r = (int)q.list[lane];
                Trav tr;
                tr.o = mk3(q.ox[r], q.oy[r], q.oz[r]);
                tr.d = mk3(q.dx[r], q.dy[r], q.dz[r]);
                tr.hit.t = q.tfar[r];
                tr.hit.u = tr.hit.v = 0;
                if (kind == Q_FIN) {
                    const int slot = q.slot[r];
                    tr.hit.prim = q.prim[r];
                    if (SHADOW) {
                        if (tr.hit.prim == kNoHit) {
                            V4 rd = a.pool.rad[slot], c = a.pool.sh_c[slot];
                            a.pool.rad[slot] = mk4(rd.x + c.x, rd.y + c.y, rd.z + c.z, rd.w);
                        }
                    } else {
                        trav_finish_closest(sc.prims, tr);
                        a.pool.hit[slot] = mk4(tr.hit.t, tr.hit.u, tr.hit.v, u2f((uint32_t)tr.hit.prim));
                    }
                    q.status[r] = Q_EMPTY;
                } else {
                    tr.tnear = q.tnear[r];
                    const int sp0 = (int)q.sp[r];
                    tr.sp = sp0;
                    StridedStack stk;
                    stk.base = stack_base + r;
                    stk.stride = kQRays;
                    if (kind == Q_NODE) {
                        tr.G.x = q.Gx[r]; tr.G.y = q.Gy[r];
                        tr.hit.prim = kNoHit;  // (not read by a node step)
                        tr.idir = trav_idir_fast(tr.d);
                        const uint32_t oct = (tr.d.x < 0 ? 1u : 0u) | (tr.d.y < 0 ? 2u : 0u) | (tr.d.z < 0 ? 4u : 0u);
                        tr.octinv4 = (7u ^ oct) * 0x01010101u;
                        node_steps++;
                        trav_node(sc.nodes8, tr, stk, a.one_bits);
                        trav_next_group(tr, stk);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                        for (int extra = 1; extra < kQNodesPerPass; extra++) {  // (a ray that still wants a node step takes it now)
                            if (tr.Gt.y == 0 && tr.G.y != 0) {
                                node_steps++; more = true;
                                trav_node(sc.nodes8, tr, stk, a.one_bits);
                                trav_next_group(tr, stk);
                            }
                        }
                        q.Gx[r] = tr.G.x; q.Gy[r] = tr.G.y; q.Tx[r] = tr.Gt.x; q.Ty[r] = tr.Gt.y;
                        if (tr.sp != sp0) q.sp[r] = (uint32_t)tr.sp;
                        const uint32_t st = q_status_of(tr);
                        if (st != (uint32_t)Q_NODE) q.status[r] = st;
                        finished = st == Q_FIN;
                    } else {
                        const float t0 = tr.hit.t;
                        const int p0 = q.prim[r];
                        tr.hit.prim = p0;
                        tr.G.x = 0; tr.G.y = q.Gy[r];  // (G.x is only needed once the group is entered: it stays in the table)
                        tr.Gt.x = q.Tx[r]; tr.Gt.y = q.Ty[r];
                        prim_tests++;
                        bool ended = trav_prim<SHADOW>(sc.prims, tr);  // any-hit ends at the first hit
                        // up to kQPrimsPerPass primitives of the ray's group in one pass, in the group's own order (the
                        // answers cannot change): selecting a ray and moving its state costs about three times what one
                        // primitive test does, and most groups hold two or three primitives
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                        for (int extra = 1; extra < kQPrimsPerPass; extra++) {
                            if (!ended && tr.Gt.y != 0) { prim_tests++; more = true; ended = trav_prim<SHADOW>(sc.prims, tr); }
                        }
                        if (ended) trav_terminate(tr);
                        if (tr.hit.t != t0 || tr.hit.prim != p0) { q.tfar[r] = tr.hit.t; q.prim[r] = tr.hit.prim; }
                        const bool pop = !ended && (tr.G.y | tr.Gt.y) == 0 && tr.sp > 0;
                        if (pop) {
                            tr.G.x = q.Gx[r];  // (a popped primitive group leaves G as it was: empty)
                            trav_next_group(tr, stk);
                            q.Gx[r] = tr.G.x; q.Gy[r] = tr.G.y; q.Tx[r] = tr.Gt.x; q.Ty[r] = tr.Gt.y;
                            q.sp[r] = (uint32_t)tr.sp;
                        } else {
                            q.Ty[r] = tr.Gt.y;
                            if (ended) { q.Gy[r] = 0; q.sp[r] = 0; }
                        }
                        const uint32_t st = q_status_of(tr);
                        if (st != (uint32_t)Q_PRIM) q.status[r] = st;
                        finished = st == Q_FIN;
                    }
                }
            }
            if (kind == Q_FIN) cF -= nsel;
            else cF += __popc(__ballot_sync(0xffffffffu, finished));  // (the ballot also orders the shared-memory writes)
            if ((kQPrimsPerPass > 1 || kQNodesPerPass > 1) && kind != Q_FIN && __ballot_sync(0xffffffffu, more)) { if (kind == Q_PRIM) prim_passes++; else node_passes++; }  // (statistics: a pass with second tests counts twice)
            __syncwarp();
        }
    }
    warp_add(&a.counters[MODE == 0 ? a.closest_counter : C_SHADOW], traced);
    warp_add(&a.counters[C_NODE_STEPS], node_steps);
    warp_add(&a.counters[C_PRIM_TESTS], prim_tests);
    warp_add(&a.counters[C_NODE_PASSES], lane == 0 ? node_passes : 0u);  // (warp-uniform counters)
    warp_add(&a.counters[C_PRIM_PASSES], lane == 0 ? prim_passes : 0u);
}


// K4.  CLASS = kMatSmall: the kernel every path-integrator scene runs.  In a scene WITH Disney materials
// (split_classes) it shades the paths whose vertex carries one of the three small materials (or missed) and APPENDS
// the slots of the others to pool.class_queue (one cursor atomic per block); the second pass (CLASS = kMatDisney)
// runs one thread per QUEUE ENTRY, so its warps are full whatever fraction of the pool hit a Disney surface -- the
// material-sorted shading of north_star at class granularity.  (Before the queue the second pass ran one thread per
// slot: 9 of 32 lanes held a Disney vertex on disney_bsdf, profiles/r02l_disney_*.)  The records stay where they are;
// the second pass gathers them through the queue.  The first pass writes the sh_mask words, the second ORs its bits in.
// Scenes whose every material is Lambertian (cbox, sponza) shade with k_shade<.., kMatLambert>: the dispatchers of the other
// materials are compiled out (0.54 MB -> 0.14 MB of SASS, 96 registers at 5 resident CTAs per SM without spills): sponza
// shade stage 131 -> 111 ms per 256 spp, cbox 45 -> 38 (profiles/r02za_lambert.txt; 6 CTAs spill and gain nothing more).
#ifndef LJ_LAMBERT_MIN_BLOCKS
#define LJ_LAMBERT_MIN_BLOCKS 5
#endif
constexpr int kCursorClass = 3;  // WaveArgs::cursors[3]: entries in pool.class_queue (path integrator only)
template <int MIN_BLOCKS, int CLASS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_shade(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a, int split_classes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    ShadeCounters cnt = {0, 0, 0, 0};
    if (CLASS == kMatDisney) {
        if ((unsigned)t < a.cursors[kCursorClass]) {
            const int i = (int)a.pool.class_queue[t];
            PathState s;
            load_state(a.pool, i, s);
            shade_path<CLASS>(sc, a.rp, s, cnt);
            store_state(a.pool, i, s, (s.flags & kAlive) != 0);
            if (s.sh_tfar >= 0) atomicOr(&a.pool.sh_mask[i / LJ_WARP_WIDTH], 1u << (i % LJ_WARP_WIDTH));
        }
    } else {
        const int i = t;
        bool has_shadow = false, queued = false;
        if (i < a.pool.capacity) {
            uint32_t flags = f2u(a.pool.meta[i].y);
            if (flags & kAlive) {
                if (!split_classes || path_material_class(sc, (int)f2u(a.pool.hit[i].w)) == CLASS) {
                    PathState s;
                    load_state(a.pool, i, s);
                    shade_path<CLASS>(sc, a.rp, s, cnt);
                    store_state(a.pool, i, s, (s.flags & kAlive) != 0);
                    has_shadow = s.sh_tfar >= 0;
                } else {
                    queued = true;
                }
            }
        }
        {   // capacity is a multiple of 256 (render_impl): whole warps are in range
            unsigned m = __ballot_sync(0xffffffffu, has_shadow);
            if (LJ_LANE() == 0 && i < a.pool.capacity) a.pool.sh_mask[i / LJ_WARP_WIDTH] = m;
        }
        if (split_classes) {
            const unsigned qm = __ballot_sync(0xffffffffu, queued);
#if defined(LJ_HOSTSIM)
            const unsigned base = queued ? atomicAdd(&a.cursors[kCursorClass], 1u) : 0u;
#else
            __shared__ unsigned s_q[4], s_qbase;
            const int warp = threadIdx.x >> 5;
            if (LJ_LANE() == 0) s_q[warp] = (unsigned)__popc(qm);
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned tot = 0;
                for (int k = 0; k < 4; k++) { unsigned c = s_q[k]; s_q[k] = tot; tot += c; }
                s_qbase = tot ? atomicAdd(&a.cursors[kCursorClass], tot) : 0u;
            }
            __syncthreads();
            const unsigned base = s_qbase + s_q[warp];
#endif
            if (queued) a.pool.class_queue[base + (unsigned)__popc(qm & ((1u << LJ_LANE()) - 1u))] = (uint32_t)i;
        }
    }
    warp_add(&a.counters[C_BOUNCES_STRIPED + ((blockIdx.x * 4 + (threadIdx.x >> 5)) & (kStripes - 1))], cnt.bounces);
}

// K5: free flight of every live volpath ray by chromatic delta tracking (homework2.tex:713-758), persistent threads.
// One warp iteration = ONE collision step for each lane that holds a path; lanes whose flight ended (real collision,
// surface reached, or null-collision limit) write the result and are refilled from the cursor once fewer than
// track_refill lanes are busy.  Result per path: thr *= transmittance / avg(trans_dir_pdf) (:814-816), the MIS caches
// vol0 / vol1 *= the two pdf products (:521-558), rng state, and hit = (distance, kScatter) on a real collision.
template <bool GRID>
__global__ void __launch_bounds__(128) k_flight(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    const unsigned n = (unsigned)a.pool.capacity;
    unsigned int *cursor = &a.cursors[2];
    const int lane = LJ_LANE();
    const unsigned chunk = (unsigned)a.chunk;
    unsigned chunk_next = 0, chunk_end = 0;
    bool drained = false, global_out = false;
    bool busy = false;
    int slot = -1, medium = -1;
    V3 o = mk3(0), d = mk3(0), tr = mk3(1), pd = mk3(1), pn = mk3(1);
    float t_hit = 0;
    Pcg rng;
    rng.state = 0; rng.inc = 1;
    TrackState ts;
    for (;;) {
        unsigned want = drained ? 0u : __ballot_sync(0xffffffffu, !busy);
        if (want) {
            const unsigned cnt = (unsigned)__popc(want);
            const unsigned lim = chunk_end < n ? chunk_end : n;
            const unsigned left = chunk_next < lim ? lim - chunk_next : 0u;
            unsigned nb = 0;
            bool fresh = false;
            if (cnt > left && !global_out) {
                if (lane == 0) nb = atomicAdd(cursor, chunk);
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if (nb >= n) global_out = true; else fresh = true;
            }
            const unsigned rank = (unsigned)__popc(want & ((1u << lane) - 1));
            const unsigned idx = rank < left ? chunk_next + rank : (fresh ? nb + (rank - left) : 0xffffffffu);
            if (fresh) { chunk_next = nb + (cnt - left); chunk_end = nb + chunk; }
            else chunk_next += cnt < left ? cnt : left;
            drained = global_out && chunk_next >= (chunk_end < n ? chunk_end : n);
            if (!busy && idx < n) {
                V4 m = a.pool.meta[idx];
                if (f2u(m.y) & kAlive) {
                    V4 x = a.pool.aux[idx];
                    medium = (int)f2u(x.w);
                    if (medium >= 0) {
                        slot = (int)idx;
                        V4 ro = a.pool.ray_o[idx], rd = a.pool.ray_d[idx], h = a.pool.hit[idx];
                        o = xyz(ro); d = xyz(rd);
                        t_hit = (int)f2u(h.w) == kNoHit ? LJ_INF : h.x;
                        rng.state = (uint64_t)f2u(m.z) | ((uint64_t)f2u(m.w) << 32);
                        rng.inc = pcg_inc(path_stream((uint64_t)f2u(m.x) * a.rp.spp_total + f2u(x.z)));
                        tr = mk3(1); pd = mk3(1); pn = mk3(1);
                        if (track_begin<GRID>(sc.media[medium], o, d, rd.w, rng, ts)) busy = true;
                        else a.pool.meta[idx] = mk4(m.x, m.y, u2f((uint32_t)rng.state), u2f((uint32_t)(rng.state >> 32)));  // channel draw only
                    }
                }
            }
        }
        if (!__ballot_sync(0xffffffffu, busy)) {
            if (drained) break;
            continue;
        }
        for (bool first = true;; first = false) {
            const unsigned bm = __ballot_sync(0xffffffffu, busy);
            if (bm == 0) break;
            if (!first && !drained && __popc(bm) < a.track_refill) break;
            if (busy) {
                int r = flight_step<GRID>(sc.media[medium], o, d, t_hit, sc.options.max_null_collisions, rng, ts, tr, pd, pn);
                if (r != kTrackContinue) {
                    V4 t = a.pool.thr[slot], v0 = a.pool.vol0[slot], v1 = a.pool.vol1[slot], m = a.pool.meta[slot];
                    float inv = 1 / avg3(pd);
                    a.pool.thr[slot] = mk4(t.x * tr.x * inv, t.y * tr.y * inv, t.z * tr.z * inv, t.w);
                    a.pool.vol0[slot] = mk4(v0.x * pd.x, v0.y * pd.y, v0.z * pd.z, v0.w);
                    a.pool.vol1[slot] = mk4(v1.x * pn.x, v1.y * pn.y, v1.z * pn.z, v1.w);
                    a.pool.meta[slot] = mk4(m.x, m.y, u2f((uint32_t)rng.state), u2f((uint32_t)(rng.state >> 32)));
                    if (r == kTrackScatter) a.pool.hit[slot] = mk4(ts.accum_t, 0, 0, u2f((uint32_t)kScatter));
                    busy = false;
                }
            }
        }
    }
}

// ---- K3 for grid media, staged.  k_trace<3> runs a whole NEE walk in one lane -- traversal of a segment, ratio
// tracking over it, the index-matched test, the next segment -- and its passes alternate between traversal steps and
// tracking steps with the lanes in the other state idle (10 of 32 lanes per instruction on hetvol_colored, 77 % of the
// stall cycles waiting for instructions: profiles/r02o_*).  Here the walk state lives in the pool between stages and
// each stage is a lean kernel:  k_walk_begin (first segment of every walk the shade kernel left) -> rounds of
// [k_trace_q<0> over the "walk view" of the pool (closest hit of every pending segment: the kernel the path integrator
// extends with) -> k_walk_track (ratio tracking over the segment, one step per lane per pass as k_flight, then the
// opaque / index-matched test: contribution added, or the next segment written)] -> k_walk_finish (the few walks that
// cross more surfaces than there were rounds, whole loops).  Same step functions and the same PCG draws in the same
// order as the serial walk: results are bit-identical (test_walk_kernels_parity).
constexpr int kWalkRounds = 3;

__global__ void __launch_bounds__(256) k_walk_begin(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pool.capacity) return;
    V4 wm = mk4(0, 0, 0, 0);
    if ((a.pool.sh_mask[i / LJ_WARP_WIDTH] >> (i % LJ_WARP_WIDTH)) & 1u) {
        const V4 sd = a.pool.sh_d[i], so = a.pool.sh_o[i], pl = a.pool.sh_pl[i];
        const uint32_t mb = f2u(so.w);
        NeeWalk wk;
        wk.pc = xyz(so); wk.dir = xyz(sd); wk.pl = xyz(pl);
        const uint64_t path_id = (uint64_t)f2u(a.pool.meta[i].x) * a.rp.spp_total + f2u(a.pool.aux[i].z);
        const Pcg rng = walk_rng(path_id, f2u(pl.w), a.rp.seed);
        float tn, tf;
        nee_walk_segment(sc, wk, tn, tf);
        a.pool.w_o[i] = mk4(wk.pc, tn);
        a.pool.w_d[i] = mk4(wk.dir, tf);
        a.pool.w_T[i] = mk4(1, 1, 1, u2f(0u));
        a.pool.w_pn[i] = mk4(1, 1, 1, u2f((uint32_t)((int)(mb & 0xffffu) - 1)));
        a.pool.w_pd[i] = mk4(1, 1, 1, u2f(mb >> 16));
        wm = mk4(0, u2f(kAlive), u2f((uint32_t)rng.state), u2f((uint32_t)(rng.state >> 32)));
    }
    a.pool.w_meta[i] = wm;
}

// the walk of slot i as the stages left it (c, the two pdfs and the light point stay in the shade kernel's records)
LJ_HD void walk_load(const PathPool &p, const RenderParams &rp, int i, const V4 &wm, NeeWalk &wk, float &seg_tnear, float &seg_tfar) {
    const V4 o = p.w_o[i], d = p.w_d[i], T = p.w_T[i], pn = p.w_pn[i], pd = p.w_pd[i], pl = p.sh_pl[i];
    wk.pc = xyz(o); wk.dir = xyz(d); wk.pl = xyz(pl);
    seg_tnear = o.w; seg_tfar = d.w;
    wk.T_light = xyz(T); wk.p_nee = xyz(pn); wk.p_dir = xyz(pd);
    wk.shadow_bounces = f2u(T.w); wk.medium = (int)f2u(pn.w); wk.budget = f2u(pd.w);
    const uint64_t path_id = (uint64_t)f2u(p.meta[i].x) * rp.spp_total + f2u(p.aux[i].z);
    wk.rng.state = (uint64_t)f2u(wm.z) | ((uint64_t)f2u(wm.w) << 32);
    wk.rng.inc = pcg_inc(path_stream(path_id) + (((uint64_t)f2u(pl.w) << 1) | 1ull));  // as walk_rng
}
// the segment is tracked: opaque / index-matched test, then the contribution or the next segment
LJ_HD void walk_decide_store(const DevScene &sc, const PathPool &p, int i, NeeWalk &wk, const Hit &hit) {
    const V4 cc = p.sh_c[i];
    wk.c = xyz(cc); wk.pdf_nee = cc.w; wk.pdf_dir = p.sh_d[i].w;
    V3 contrib;
    if (nee_walk_decide(sc, wk, hit, contrib)) {
        if (max3(contrib) > 0 || min3(contrib) < 0 || contrib.x != contrib.x || contrib.y != contrib.y || contrib.z != contrib.z) {
            V4 r = p.rad[i];
            p.rad[i] = mk4(r.x + contrib.x, r.y + contrib.y, r.z + contrib.z, r.w);
        }
        p.w_meta[i] = mk4(0, 0, 0, 0);
        return;
    }
    float tn, tf;
    nee_walk_segment(sc, wk, tn, tf);
    p.w_o[i] = mk4(wk.pc, tn);
    p.w_d[i] = mk4(wk.dir, tf);
    p.w_T[i] = mk4(wk.T_light, u2f(wk.shadow_bounces));
    p.w_pn[i] = mk4(wk.p_nee, u2f((uint32_t)wk.medium));
    p.w_pd[i] = mk4(wk.p_dir, u2f(wk.budget));
    p.w_meta[i] = mk4(0, u2f(kAlive), u2f((uint32_t)wk.rng.state), u2f((uint32_t)(wk.rng.state >> 32)));
}

__global__ void __launch_bounds__(128) k_walk_track(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    const unsigned n = (unsigned)a.pool.capacity;
    unsigned int *cursor = &a.cursors[1];
    const int lane = LJ_LANE();
    const unsigned chunk = (unsigned)a.chunk;
    unsigned chunk_next = 0, chunk_end = 0;
    bool drained = false, global_out = false;
    bool busy = false, tracking = false;
    int slot = -1;
    NeeWalk wk;
    wk.medium = -1; wk.pc = mk3(0); wk.dir = mk3(0); wk.T_light = mk3(1); wk.p_nee = mk3(1); wk.p_dir = mk3(1);
    wk.rng.state = 0; wk.rng.inc = 1;
    Hit hit;
    hit.prim = kNoHit; hit.t = 0; hit.u = 0; hit.v = 0;
    float next_t = 0;
    TrackState ts;
    for (;;) {
        unsigned want = drained ? 0u : __ballot_sync(0xffffffffu, !busy);
        if (want) {
            const unsigned cnt = (unsigned)__popc(want);
            const unsigned lim = chunk_end < n ? chunk_end : n;
            const unsigned left = chunk_next < lim ? lim - chunk_next : 0u;
            unsigned nb = 0;
            bool fresh = false;
            if (cnt > left && !global_out) {
                if (lane == 0) nb = atomicAdd(cursor, chunk);
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if (nb >= n) global_out = true; else fresh = true;
            }
            const unsigned rank = (unsigned)__popc(want & ((1u << lane) - 1));
            const unsigned idx = rank < left ? chunk_next + rank : (fresh ? nb + (rank - left) : 0xffffffffu);
            if (fresh) { chunk_next = nb + (cnt - left); chunk_end = nb + chunk; }
            else chunk_next += cnt < left ? cnt : left;
            drained = global_out && chunk_next >= (chunk_end < n ? chunk_end : n);
            if (!busy && idx < n) {
                const V4 wm = a.pool.w_meta[idx];
                if (f2u(wm.y) & kAlive) {
                    slot = (int)idx;
                    float tn, seg_tfar;
                    walk_load(a.pool, a.rp, slot, wm, wk, tn, seg_tfar);
                    const V4 h = a.pool.w_hit[slot];
                    hit.t = h.x; hit.u = h.y; hit.v = h.z; hit.prim = (int)f2u(h.w);
                    next_t = nee_walk_next_t(wk, hit);
                    tracking = wk.medium >= 0 && track_begin(sc.media[wk.medium], wk.pc, wk.dir, seg_tfar, wk.rng, ts);
                    busy = true;
                }
            }
        }
        if (!__ballot_sync(0xffffffffu, busy)) {
            if (drained) break;
            continue;
        }
        for (bool first = true;; first = false) {
            const unsigned bm = __ballot_sync(0xffffffffu, busy);
            if (bm == 0) break;
            if (!first && !drained && __popc(bm) < a.track_refill) break;
            if (busy) {
                if (tracking && ratio_step(sc.media[wk.medium], wk.pc, wk.dir, next_t, sc.options.max_null_collisions, wk.rng, ts,
                                                 wk.T_light, wk.p_nee, wk.p_dir) != kTrackContinue) tracking = false;
                if (!tracking) {
                    walk_decide_store(sc, a.pool, slot, wk, hit);
                    busy = false;
                }
            }
        }
    }
}

// the walks still in flight after the last round (more index-matched surfaces than rounds): whole loops, one thread each
__global__ void __launch_bounds__(128) k_walk_finish(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.pool.capacity) return;
    const V4 wm = a.pool.w_meta[i];
    if (!(f2u(wm.y) & kAlive)) return;
    NeeWalk wk;
    float tn, tf;
    walk_load(a.pool, a.rp, i, wm, wk, tn, tf);
    const V4 cc = a.pool.sh_c[i];
    wk.c = xyz(cc); wk.pdf_nee = cc.w; wk.pdf_dir = a.pool.sh_d[i].w;
    V3 contrib = mk3(0);
    for (int guard = 0; guard < 1 << 16; guard++) {
        Hit hit;
        trace8<false>(sc.nodes8, sc.prims, wk.pc, wk.dir, tn, tf, hit);
        const float next_t = nee_walk_next_t(wk, hit);
        if (wk.medium >= 0) ratio_track(sc.media[wk.medium], wk.pc, wk.dir, tf, next_t, sc.options.max_null_collisions, wk.rng, wk.T_light, wk.p_nee, wk.p_dir);
        if (nee_walk_decide(sc, wk, hit, contrib)) break;
        nee_walk_segment(sc, wk, tn, tf);
    }
    if (max3(contrib) > 0 || min3(contrib) < 0 || contrib.x != contrib.x || contrib.y != contrib.y || contrib.z != contrib.z) {
        V4 r = a.pool.rad[i];
        a.pool.rad[i] = mk4(r.x + contrib.x, r.y + contrib.y, r.z + contrib.z, r.w);
    }
    a.pool.w_meta[i] = mk4(0, 0, 0, 0);
}

// K4 + K5 for the volpath integrator (lj_volpath.h)
// PASS 0: one thread per slot; shades the medium events (scattering, index-matched boundaries, escapes) and appends the
// slots whose vertex lies on a surface WITH a material to pool.class_queue; PASS 1: one thread per queue entry shades
// those -- full warps in the BSDF code instead of a few lanes of every warp (the same queue as the Disney pass of the
// path integrator; before it k_shade_vol ran at 13.6 lanes per instruction, profiles/r02o_*).
LJ_HD bool vol_vertex_has_material(const DevScene &sc, int prim) {
    return prim >= 0 && sc.shapes[prim_shape_id(ld4(&sc.prims[prim].c))].material_id >= 0;
}
template <int PASS>
__global__ void __launch_bounds__(128) k_shade_vol(const LJ_GRID_CONSTANT DevScene sc, const LJ_GRID_CONSTANT WaveArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    ShadeCounters cnt = {0, 0, 0, 0};
    if (PASS == 1) {
        if ((unsigned)t < a.cursors[kCursorClass]) {
            const int i = (int)a.pool.class_queue[t];
            PathState s;
            load_state_vol(a.pool, i, s);
            shade_vol_path(sc, a.rp, s, cnt);
            store_state_vol(a.pool, i, s, (s.flags & kAlive) != 0);
            if (s.sh_pdf_dir >= 0) atomicOr(&a.pool.sh_mask[i / LJ_WARP_WIDTH], 1u << (i % LJ_WARP_WIDTH));
        }
    } else {
        const int i = t;
        bool has_shadow = false, queued = false;
        if (i < a.pool.capacity) {
            uint32_t flags = f2u(a.pool.meta[i].y);
            if (flags & kAlive) {
                if (vol_vertex_has_material(sc, (int)f2u(a.pool.hit[i].w))) {
                    queued = true;
                } else {
                    PathState s;
                    load_state_vol(a.pool, i, s);
                    shade_vol_path(sc, a.rp, s, cnt);
                    store_state_vol(a.pool, i, s, (s.flags & kAlive) != 0);
                    has_shadow = s.sh_pdf_dir >= 0;
                }
            }
        }
        {
            unsigned m = __ballot_sync(0xffffffffu, has_shadow);
            if (LJ_LANE() == 0 && i < a.pool.capacity) a.pool.sh_mask[i / LJ_WARP_WIDTH] = m;
        }
        const unsigned qm = __ballot_sync(0xffffffffu, queued);
#if defined(LJ_HOSTSIM)
        const unsigned base = queued ? atomicAdd(&a.cursors[kCursorClass], 1u) : 0u;
#else
        __shared__ unsigned s_q[4], s_qbase;
        const int warp = threadIdx.x >> 5;
        if (LJ_LANE() == 0) s_q[warp] = (unsigned)__popc(qm);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
            for (int k = 0; k < 4; k++) { unsigned c = s_q[k]; s_q[k] = tot; tot += c; }
            s_qbase = tot ? atomicAdd(&a.cursors[kCursorClass], tot) : 0u;
        }
        __syncthreads();
        const unsigned base = s_qbase + s_q[warp];
#endif
        if (queued) a.pool.class_queue[base + (unsigned)__popc(qm & ((1u << LJ_LANE()) - 1u))] = (uint32_t)i;
    }
    warp_add(&a.counters[C_BOUNCES_STRIPED + ((blockIdx.x * 4 + (threadIdx.x >> 5)) & (kStripes - 1))], cnt.bounces);
}

__global__ void k_clear_pool(PathPool pool) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pool.capacity) return;
    pool.meta[i] = mk4(0, 0, 0, 0);
    pool.sh_d[i] = mk4(0, 0, 0, -1.f);
    pool.rad[i] = mk4(0, 0, 0, 1.f);
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

// The five auxiliary integrators of render.cpp:12-69 (depth, shadingNormal, meanCurvature, rayDifferential,
// mipmapLevel): one primary ray through each pixel centre, no sampling -- images are comparable pixel by pixel
// with the reference's.
__global__ void __launch_bounds__(128) k_aux(const LJ_GRID_CONSTANT DevScene sc, int integrator, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = sc.camera.width, h = sc.camera.height;
    if (i >= w * h) return;
    int x = i % w, y = i / w;
    V3 o, d;
    sample_primary(sc.camera, mk2((x + 0.5f) / w, (y + 0.5f) / h), o, d);
    const float spread = init_ray_spread(w, h);
    Hit hit;
    V3 color = mk3(0);
    if (trace8<false>(sc.nodes8, sc.prims, o, d, 0.f, LJ_INF, hit)) {
        Vertex vx = make_vertex(sc, o, d, hit, 0.f, spread);
        if (integrator == LJ_INT_DEPTH) {
            color = mk3(distance(vx.position, o));
        } else if (integrator == LJ_INT_SHADING_NORMAL) {
            color = vx.shading_frame.n;
        } else if (integrator == LJ_INT_MEAN_CURVATURE) {
            color = mk3(vx.mean_curvature);
        } else if (integrator == LJ_INT_RAY_DIFFERENTIAL) {
            color = mk3(0.f, spread, 0.f);
        } else if (vx.material_id >= 0) {  // LJ_INT_MIPMAP_LEVEL: the texture get_texture() names (material.cpp:68-88)
            const DevMaterial &m = sc.materials[vx.material_id];
            const DevTexture &t = m.tex[m.type == LJ_MAT_ROUGHDIELECTRIC ? 1 : 0];
            if (t.kind == LJ_TEX_IMAGE && m.type != LJ_MAT_DISNEY_CLEARCOAT) {
                const DevImage &im = sc.images3[t.image_id];
                float scaled = (float)(im.w[0] > im.h[0] ? im.w[0] : im.h[0]) * fmaxf(t.uscale, t.vscale) * vx.uv_screen_size;
                color = mk3(log2f(fmaxf(scaled, 1e-8f)));
            }
        }
    }
    out[3 * i] = color.x; out[3 * i + 1] = color.y; out[3 * i + 2] = color.z;
}

static int ensure_pool(lj_scene *s, int capacity, bool vol) {
    if (s->pool_capacity == capacity && s->pool_block && (s->pool.vol0 != nullptr) == vol) return LJ_OK;
    if (s->pool_block) { pool_block_give(s->device, s->pool_block, s->pool_bytes); s->pool_block = nullptr; }
    const int kFields = vol ? 21 : 9;
    const size_t mask_words = ((size_t)capacity + LJ_WARP_WIDTH - 1) / LJ_WARP_WIDTH;
    const size_t queue_words = (vol || s->has_disney) ? (size_t)capacity : 0;  // class_queue of the second shade pass
    s->pool_block = pool_block_take(s->device, (size_t)capacity * sizeof(V4) * kFields + (mask_words + queue_words) * sizeof(uint32_t), &s->pool_bytes);
    if (!s->pool_block) { s->pool_capacity = 0; return cuda_fail(cudaErrorMemoryAllocation, "path pool allocation"); }
    V4 *base = (V4 *)s->pool_block;
    PathPool &p = s->pool;
    p.ray_o = base + (size_t)capacity * 0; p.ray_d = base + (size_t)capacity * 1; p.hit = base + (size_t)capacity * 2;
    p.thr = base + (size_t)capacity * 3; p.rad = base + (size_t)capacity * 4; p.sh_d = base + (size_t)capacity * 5;
    p.sh_c = base + (size_t)capacity * 6; p.meta = base + (size_t)capacity * 7; p.aux = base + (size_t)capacity * 8;
    p.vol0 = p.vol1 = p.vol2 = p.sh_o = p.sh_pl = nullptr;
    p.w_o = p.w_d = p.w_hit = p.w_meta = p.w_T = p.w_pn = p.w_pd = nullptr;
    if (vol) {
        p.w_o = base + (size_t)capacity * 14; p.w_d = base + (size_t)capacity * 15; p.w_hit = base + (size_t)capacity * 16;
        p.w_meta = base + (size_t)capacity * 17; p.w_T = base + (size_t)capacity * 18; p.w_pn = base + (size_t)capacity * 19;
        p.w_pd = base + (size_t)capacity * 20;
        p.vol0 = base + (size_t)capacity * 9; p.vol1 = base + (size_t)capacity * 10; p.vol2 = base + (size_t)capacity * 11;
        p.sh_o = base + (size_t)capacity * 12; p.sh_pl = base + (size_t)capacity * 13;
    }
    p.sh_mask = (uint32_t *)(base + (size_t)capacity * kFields);
    p.class_queue = queue_words ? p.sh_mask + mask_words : nullptr;
    p.capacity = capacity;
    s->pool_capacity = capacity;
    return LJ_OK;
}

struct EventPool {  // hands out the scene's events in order; they live until lj_scene_destroy
    std::vector<cudaEvent_t> &ev;
    size_t used = 0;
    explicit EventPool(std::vector<cudaEvent_t> &pool) : ev(pool) {}
    cudaEvent_t next() {
        if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
        return ev[used++];
    }
};

// Tuning overrides, read from the environment ONCE per process (first render); the defaults are the measured
// optimum on sponza / hetvol (profiles/).  LJ_TRACE_KERNEL=0 selects the one-ray-per-lane k_trace<0|1> instead of
// k_trace_q for the path integrator (A/B runs and the parity tests that compare the two).
struct Tuning {
    int prim_min_lanes = kPrimMinLanes, refill = kRefillThreshold, track_refill = 24, shadow_chunk = 128, trav_min = 4, chunk = 64;
    int trace_kernel = 1, q_refill = 48, q_chunk = 128;
    int lambert_kernel = 1;  // scenes with Lambertian materials only: k_shade<.., kMatLambert> (0: the general kernel, A/B)
    int walk_kernel = 1;  // volpath NEE walk: 0 = k_trace<2|3> (one lane per walk), 1 = staged kernels for grid media, 2 = staged for every scene
    // resident CTAs of k_trace_q per SM (0: what fits); shared-memory carve-out in % (-1: the maximum).  8 CTAs need 8 x 17.9 KB =
    // 63 % of the SM's 228 KB; asking for just that leaves 64 KB more L1 than the maximum carve-out does: sponza extend stage
    // 283 -> 264 ms per 256 spp (profiles/r02ze_sweeps.txt)
    int q_blocks = 0, q_carveout = 66;
    bool host_prof = false;
};
static const Tuning &tuning() {
    static const Tuning t = [] {
        Tuning v;
        auto geti = [](const char *name, int &dst, int lo, int hi) {
            if (const char *e = getenv(name)) dst = std::min(hi, std::max(lo, atoi(e)));
        };
        geti("LJ_PRIM_MIN_LANES", v.prim_min_lanes, 0, 32);
        geti("LJ_REFILL", v.refill, 0, 32);
        geti("LJ_TRACK_REFILL", v.track_refill, 0, 32);
        geti("LJ_SHADOW_CHUNK", v.shadow_chunk, 32, 1 << 16);
        geti("LJ_TRAV_MIN", v.trav_min, 1, 32);
        geti("LJ_CHUNK", v.chunk, 32, 1 << 16);
        geti("LJ_TRACE_KERNEL", v.trace_kernel, 0, 1);
        geti("LJ_WALK_KERNEL", v.walk_kernel, 0, 2);
        geti("LJ_LAMBERT_KERNEL", v.lambert_kernel, 0, 1);
        geti("LJ_Q_REFILL", v.q_refill, 1, kQRays);
        geti("LJ_Q_CHUNK", v.q_chunk, kQRays, 1 << 16);
        geti("LJ_Q_BLOCKS", v.q_blocks, 0, 32);
        geti("LJ_Q_CARVEOUT", v.q_carveout, -1, 100);
        v.host_prof = getenv("LJ_PROFILE_HOST") != nullptr;
        return v;
    }();
    return t;
}

// Persistent grids: exactly one wave of resident CTAs (SM count x the occupancy of each kernel), per device.
static int ensure_launch_geometry(lj_scene *s) {
    LaunchGeom &g = s->geom;
    if (g.trace_blocks > 0) return LJ_OK;
#if defined(LJ_HOSTSIM)
    g.trace_blocks = g.walk_blocks = g.step_blocks = g.flight_blocks = g.q_blocks = g.wtrack_blocks = 1;
#else
    int sms = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, q0 = 0, q1 = 0;
    LJ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    // k_trace_q keeps its ray tables in shared memory: ask for the largest carve-out so 8 CTAs stay resident
    const int carve = tuning().q_carveout >= 0 ? tuning().q_carveout : (int)cudaSharedmemCarveoutMaxShared;
    cudaFuncSetAttribute(k_trace_q<0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(k_trace_q<1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a0, k_trace<0>, 128, 0));
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a1, k_trace<1>, 128, 0));
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a2, k_trace<2>, 128, 0));
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a4, k_trace<3>, 128, 0));
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a3, k_flight<true>, 128, 0));
    {   // (the variant for scenes without grid media never needs more registers)
        int a3h = 0;
        LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a3h, k_flight<false>, 128, 0));
        a3 = std::min(a3, a3h);
    }
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q0, k_trace_q<0>, kQWarps * LJ_WARP_WIDTH, 0));
    LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q1, k_trace_q<1>, kQWarps * LJ_WARP_WIDTH, 0));
    sms = std::max(1, sms);
    g.trace_blocks = sms * std::max(1, std::min(a0, a1));
    g.walk_blocks = sms * std::max(1, a2);
    g.step_blocks = sms * std::max(1, a4);
    g.flight_blocks = sms * std::max(1, a3);
    {
        int w0 = 0;
        LJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w0, k_walk_track, 128, 0));
        g.wtrack_blocks = sms * std::max(1, w0);
    }
    g.q_blocks = sms * std::max(1, tuning().q_blocks > 0 ? std::min(tuning().q_blocks, std::min(q0, q1)) : std::min(q0, q1));
#endif
    return LJ_OK;
}

// scratch stacks of k_trace_q: (tree depth + 2) entries x kQRays rays x 8 B per warp of the persistent grid
static int ensure_qstack(lj_scene *s) {
    int r = ensure_launch_geometry(s);
    if (r != LJ_OK) return r;
    if (s->d_qstack) return LJ_OK;
    s->qdepth = std::max(4, s->info.bvh_depth + 2);
    size_t bytes = (size_t)s->geom.q_blocks * kQWarps * s->qdepth * kQRays * sizeof(U2);
    // (plain cudaMalloc memory, recycled through the spares of scene.cu: the kernel's hottest global buffer stays out
    //  of the stream-ordered pool, whose mappings a host process's NCCL may widen to its peers)
    s->qstack_bytes = bytes;
    if (!(s->d_qstack = spare_take(kSpareDevice, s->device, bytes))) return cuda_fail(cudaErrorMemoryAllocation, "traversal scratch allocation");
    return LJ_OK;
}

static void fill_trace_args(lj_scene *s, WaveArgs &a) {
    const Tuning &t = tuning();
    a.prim_min_lanes = t.prim_min_lanes;
    a.refill_threshold = t.refill;
    a.track_refill = t.track_refill;
    a.shadow_chunk = t.shadow_chunk;
    a.trav_min = t.trav_min;
    a.chunk = t.chunk;
    a.q_refill = std::min(t.q_refill, kQRays);  // (the host simulation has 2 ray slots per "warp")
    a.q_chunk = t.q_chunk;
    a.one_bits = 0x3f800000u;
    a.walk_whole_groups = s->info.num_bvh_nodes <= 4 ? 1 : 0;
    a.closest_counter = C_CLOSEST;
    a.qstack = (U2 *)s->d_qstack;
    a.qdepth = s->qdepth;
    a.cursors = s->d_cursors;
    a.counters = s->d_counters;
}

// closest-hit / shadow stage of one wave of the path integrator, as lj_render launches it
static void launch_trace(lj_scene *s, const WaveArgs &a, int mode, cudaStream_t stream) {
    const DevScene &sc = s->dev;
    if (tuning().trace_kernel == 1) {
        if (mode == 0) LJ_LAUNCH(k_trace_q<0>, s->geom.q_blocks, kQWarps * LJ_WARP_WIDTH, stream, sc, a);
        else LJ_LAUNCH(k_trace_q<1>, s->geom.q_blocks, kQWarps * LJ_WARP_WIDTH, stream, sc, a);
    } else {
        if (mode == 0) LJ_LAUNCH(k_trace<0>, s->geom.trace_blocks, 128, stream, sc, a);
        else LJ_LAUNCH(k_trace<1>, s->geom.trace_blocks, 128, stream, sc, a);
    }
}

// NEE walk stage of one volpath wave in its staged form (k_walk_begin ... k_walk_finish above)
// (default: scenes with grid media, and scenes with a real hierarchy -- there the queue-form traversal kernel beats the
//  one-ray-per-lane loop of k_trace<2>: vol_cbox_teapot walk stage 265 -> 227 ms per 256 spp; on the few-primitive test
//  scenes k_trace<2> with whole primitive groups is faster, volpath_test6 71 vs 115 ms)
static bool walk_staged(const lj_scene *s) {
    return tuning().walk_kernel == 2 || (tuning().walk_kernel == 1 && (s->has_grid_media || s->info.num_bvh_nodes > 64));
}
static uint64_t launch_walk_staged(lj_scene *s, const WaveArgs &a, cudaStream_t stream, int rounds = kWalkRounds) {
    const DevScene &sc = s->dev;
    const int nb256 = (a.pool.capacity + 255) / 256, nb128 = (a.pool.capacity + 127) / 128;
    LJ_LAUNCH(k_walk_begin, nb256, 256, stream, sc, a);
    WaveArgs v = a;  // the walk view of the pool: the pending segments are the "rays", w_meta says which slots hold one
    v.pool.ray_o = a.pool.w_o; v.pool.ray_d = a.pool.w_d; v.pool.hit = a.pool.w_hit; v.pool.meta = a.pool.w_meta;
    v.closest_counter = C_SHADOW;
    for (int r = 0; r < rounds; r++) {
        cudaMemsetAsync(s->d_cursors, 0, 2 * sizeof(unsigned int), stream);
        LJ_LAUNCH(k_trace_q<0>, s->geom.q_blocks, kQWarps * LJ_WARP_WIDTH, stream, sc, v);
        LJ_LAUNCH(k_walk_track, s->geom.wtrack_blocks, 128, stream, sc, a);
    }
    LJ_LAUNCH(k_walk_finish, nb128, 128, stream, sc, a);
    return 2 + 2 * (uint64_t)rounds;
}

static int ensure_render_buffers(lj_scene *s, int npix, bool want_sq) {
    // films and the pinned counters come from the process-wide spares (scene.cu); the rest from the stream-ordered pool
    s->film_bytes = (size_t)npix * 16;
    s->h_counters_bytes = sizeof(unsigned long long) * (C_TOTAL + 8);
    if (!s->d_film && !(s->d_film = (float *)spare_take(kSpareDevice, s->device, s->film_bytes))) return cuda_fail(cudaErrorMemoryAllocation, "film allocation");
    if (want_sq && !s->d_film_sq && !(s->d_film_sq = (float *)spare_take(kSpareDevice, s->device, s->film_bytes))) return cuda_fail(cudaErrorMemoryAllocation, "film allocation");
    const bool fresh = !s->d_counters || !s->d_cursors || !s->d_qstack;
    if (!s->d_counters) LJ_CUDA(lj_dev_alloc((void **)&s->d_counters, sizeof(unsigned long long) * C_TOTAL));
    if (!s->h_counters && !(s->h_counters = (unsigned long long *)spare_take(kSpareHost, s->device, s->h_counters_bytes))) return cuda_fail(cudaErrorMemoryAllocation, "pinned counters");
    if (!s->d_cursors) LJ_CUDA(lj_dev_alloc((void **)&s->d_cursors, 4 * sizeof(unsigned int)));
    int r = ensure_qstack(s);
    // (the stream-ordered allocations were made on the default stream and are used on the caller's, which may be a
    //  non-blocking one: once per scene)
    if (r == LJ_OK && fresh) LJ_CUDA(cudaStreamSynchronize((cudaStream_t)0));
    return r;
}

static int render_impl(lj_scene *s, const lj_render_opts *opts_in, float *d_out, float *d_var, cudaStream_t stream, lj_stats *stats) {
    lj_render_opts opts;
    memset(&opts, 0, sizeof(opts));
    if (opts_in) opts = *opts_in;
    const DevScene &sc = s->dev;
    if (sc.options.integrator >= LJ_INT_DEPTH && sc.options.integrator <= LJ_INT_MIPMAP_LEVEL) {  // render.cpp:157-163
        int npx = sc.camera.width * sc.camera.height;
        LJ_LAUNCH(k_aux, (npx + 127) / 128, 128, stream, sc, sc.options.integrator, d_out);
        LJ_CUDA(cudaStreamSynchronize(stream));
        LJ_CUDA(cudaGetLastError());
        if (stats) { memset(stats, 0, sizeof(*stats)); stats->kernel_launches = 1; stats->closest_rays = (uint64_t)npx; }
        return LJ_OK;
    }
    if (sc.options.integrator != LJ_INT_PATH && sc.options.integrator != LJ_INT_VOLPATH) {
        set_error("unknown integrator");
        return LJ_ERR_UNSUPPORTED;
    }
    const bool vol = sc.options.integrator == LJ_INT_VOLPATH;
    if (vol && sc.envmap_light_id >= 0) {  // homework2.tex:196
        set_error("volpath does not support environment maps");
        return LJ_ERR_UNSUPPORTED;
    }
    int spp = opts.spp > 0 ? opts.spp : sc.options.spp;
    int sb = opts.sample_begin, se = opts.sample_end;
    if (sb == 0 && se == 0) se = spp;
    if (sb < 0 || se > spp || sb >= se) { set_error("bad sample range"); return LJ_ERR_INVALID; }
    if (opts.pool_paths < 0) { set_error("pool_paths < 0"); return LJ_ERR_INVALID; }
    const int tile_stride = opts.tile_stride > 0 ? opts.tile_stride : 1;
    if (opts.tile_offset < 0 || opts.tile_offset >= tile_stride) { set_error("bad tile split"); return LJ_ERR_INVALID; }
    int w = sc.camera.width, h = sc.camera.height, npix = w * h;
    const int tiles_x = (w + 7) / 8, tiles_y = (h + 3) / 4, tiles = tiles_x * tiles_y;
    const int tiles_local = tiles > opts.tile_offset ? (tiles - opts.tile_offset + tile_stride - 1) / tile_stride : 0;
    const unsigned long long total_items = (unsigned long long)tiles_local * 32ull * (unsigned)(se - sb);
    // Pool capacity: a multiple of 256 slots (whole blocks of k_regen, whole warps and sh_mask words everywhere), at
    // least 1024, and no more slots than there are samples to start.  The default is sized to the work: 4 or 8 Mi slots for
    // a full-size render, fewer when the call renders a small share (a rank of a strong-scaling run, a small image),
    // so that short renders do not pay for clearing and sweeping an almost empty pool.
    // (8 Mi slots for long renders: the persistent traversal kernels lose a fixed ramp-down per launch, and a pool twice
    //  the size halves its share -- sponza at 1024 spp 203 -> 213 Msamples/s, profiles/r02x_pool.txt; at 128 spp the two
    //  sizes tie, and larger pools lose to the sweep of an emptying pool at the end of the render)
    long long capacity = opts.pool_paths > 0 ? opts.pool_paths : (total_items >= (64ull << 20) ? (1 << 23) : (1 << 22));
    if ((unsigned long long)capacity > total_items) capacity = (long long)total_items;
    capacity = std::max<long long>(1024, (capacity + 255) / 256 * 256);
    const Tuning &tune = tuning();
    const bool host_prof = tune.host_prof;  // host-side phase times on stderr
    auto hp_t0 = std::chrono::steady_clock::now();
    int r = ensure_pool(s, (int)capacity, vol);
    if (r != LJ_OK) return r;
    r = ensure_render_buffers(s, npix, d_var != nullptr);
    if (r != LJ_OK) return r;
    auto hp_t1 = std::chrono::steady_clock::now();
    unsigned long long *d_counters = s->d_counters;
    unsigned long long *h_counters = s->h_counters;  // C_COUNT final counters, then the ring of per-wave live-path counts
    unsigned long long *h_active = h_counters + C_TOTAL;

    WaveArgs a;
    memset(&a, 0, sizeof(a));
    a.pool = s->pool;
    a.rp.spp_total = (uint32_t)spp;
    a.rp.sample_begin = (uint32_t)sb;
    a.rp.sample_end = (uint32_t)se;
    a.rp.seed = opts.seed ? opts.seed : kPcgDefaultSeed;
    a.rp.width = w;
    a.rp.height = h;
    a.film = s->d_film;
    a.film_sq = d_var ? s->d_film_sq : nullptr;
    a.tiles_x = tiles_x;
    a.tiles_y = tiles_y;
    a.tile_stride = tile_stride;
    a.tile_offset = opts.tile_offset;
    a.tiles_local = tiles_local;
    a.total_items = total_items;
    fill_trace_args(s, a);
    // (k_shade is compiled for 4 resident CTAs per SM, 124 registers: caps of 5 / 6 / 8 CTAs spilled and were slower)

    const int nb256 = ((int)capacity + 255) / 256, nb128 = ((int)capacity + 127) / 128;
    const LaunchGeom &g = s->geom;
    unsigned int *d_cursors = s->d_cursors;
    EventPool evp(s->event_pool);
    std::vector<cudaEvent_t> marks;  // 5 per wave: before regen, extend, shade, shadow, after shadow
    uint64_t launches = 0, waves = 0;

    cudaEvent_t ev_begin = evp.next(), ev_end = evp.next();
    LJ_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * C_TOTAL, stream));
    LJ_CUDA(cudaMemsetAsync(s->d_film, 0, (size_t)npix * 16, stream));
    if (a.film_sq) LJ_CUDA(cudaMemsetAsync(s->d_film_sq, 0, (size_t)npix * 16, stream));
    LJ_CUDA(cudaEventRecord(ev_begin, stream));
    LJ_LAUNCH(k_clear_pool, nb256, 256, stream, s->pool);
    launches++;
    // The host runs kLookahead waves ahead of the device: the live-path count of wave w is copied to pinned memory
    // right after its regen kernel and only read (after waiting for that copy) when wave w + kLookahead is being
    // queued, so the stream never drains between waves.  The waves queued behind the first empty one find an
    // empty pool and cost a few microseconds each.
    constexpr int kLookahead = 2, kRing = 4;
    cudaEvent_t ev_copied[kRing];
    for (auto &e : ev_copied) e = evp.next();
    uint64_t queued = 0;
    for (;;) {
        cudaEvent_t e0 = evp.next(), e1 = evp.next(), e2 = evp.next(), e3 = evp.next(), e4 = evp.next();
        LJ_CUDA(cudaMemsetAsync(&d_counters[C_ACTIVE], 0, sizeof(unsigned long long), stream));
        LJ_CUDA(cudaEventRecord(e0, stream));
        LJ_LAUNCH(k_regen, nb256, 256, stream, sc, a);
        LJ_CUDA(cudaEventRecord(e1, stream));
        LJ_CUDA(cudaMemcpyAsync(&h_active[queued % kRing], &d_counters[C_ACTIVE], sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        LJ_CUDA(cudaEventRecord(ev_copied[queued % kRing], stream));
        launches++;
        if (queued >= (uint64_t)kLookahead) {
            uint64_t wv = queued - kLookahead;
            LJ_CUDA(cudaEventSynchronize(ev_copied[wv % kRing]));
            if (h_active[wv % kRing] == 0) { waves = wv; marks.push_back(e0); marks.push_back(e1); marks.push_back(nullptr); break; }
        }
        LJ_CUDA(cudaMemsetAsync(d_cursors, 0, 4 * sizeof(unsigned int), stream));
        // (volpath scenes of a few primitives keep the one-ray-per-lane kernel for their extension rays; with a real
        //  hierarchy the queue form wins there as it does for the path integrator -- same hits either way)
        if (vol && s->info.num_bvh_nodes <= 64) LJ_LAUNCH(k_trace<0>, g.trace_blocks, 128, stream, sc, a);
        else launch_trace(s, a, 0, stream);
        LJ_CUDA(cudaEventRecord(e2, stream));
        if (vol) {
            if (sc.num_media > 0) {
                if (s->has_grid_media) LJ_LAUNCH(k_flight<true>, g.flight_blocks, 128, stream, sc, a);
                else LJ_LAUNCH(k_flight<false>, g.flight_blocks, 128, stream, sc, a);
                launches++;
            }
            LJ_LAUNCH(k_shade_vol<0>, nb128, 128, stream, sc, a);
            LJ_LAUNCH(k_shade_vol<1>, nb128, 128, stream, sc, a);
            launches++;
        } else {
            // one pass per material class; the classes are disjoint and a path's class is read from its hit record,
            // which the shade passes do not modify: every live path is shaded exactly once per wave
            if (s->only_lambertian && tune.lambert_kernel) {
                LJ_LAUNCH((k_shade<LJ_LAMBERT_MIN_BLOCKS, kMatLambert>), nb128, 128, stream, sc, a, 0);
            } else {
                LJ_LAUNCH((k_shade<4, kMatSmall>), nb128, 128, stream, sc, a, s->has_disney ? 1 : 0);
                if (s->has_disney) { LJ_LAUNCH((k_shade<4, kMatDisney>), nb128, 128, stream, sc, a, 1); launches++; }
            }
        }
        LJ_CUDA(cudaEventRecord(e3, stream));
        if (vol && walk_staged(s)) launches += launch_walk_staged(s, a, stream) - 1;
        else if (vol && s->has_grid_media) LJ_LAUNCH(k_trace<3>, g.step_blocks, 128, stream, sc, a);
        else if (vol) LJ_LAUNCH(k_trace<2>, g.walk_blocks, 128, stream, sc, a);
        else launch_trace(s, a, 1, stream);
        LJ_CUDA(cudaEventRecord(e4, stream));
        launches += 3;
        queued++;
        marks.push_back(e0); marks.push_back(e1); marks.push_back(e2); marks.push_back(e3); marks.push_back(e4);
        if (queued > 1000000) { set_error("wavefront loop did not terminate"); return LJ_ERR_CUDA; }
    }
    auto hp_t2 = std::chrono::steady_clock::now();
    LJ_CUDA(cudaEventRecord(ev_end, stream));
    if (d_out) {
        LJ_LAUNCH(k_resolve, (npix + 255) / 256, 256, stream, s->d_film, a.film_sq, npix, 1.f / (float)(se - sb), opts.normalize, d_out, d_var);
        launches++;
    }
    LJ_CUDA(cudaMemcpyAsync(h_counters, d_counters, sizeof(unsigned long long) * C_TOTAL, cudaMemcpyDeviceToHost, stream));
    LJ_CUDA(cudaStreamSynchronize(stream));
    LJ_CUDA(cudaGetLastError());
    auto hp_t3 = std::chrono::steady_clock::now();
    if (host_prof) {
        auto ms_ = [](auto a_, auto b_) { return std::chrono::duration<double, std::milli>(b_ - a_).count(); };
        fprintf(stderr, "lj_render host: alloc %.2f ms, queue loop %.2f ms, drain %.2f ms, events %zu\n", ms_(hp_t0, hp_t1), ms_(hp_t1, hp_t2),
                ms_(hp_t2, hp_t3), s->event_pool.size());
    }
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
        stats->samples = h_counters[C_SAMPLES];
        stats->closest_rays = h_counters[C_CLOSEST];
        stats->shadow_rays = h_counters[C_SHADOW];
        stats->bounces = h_counters[C_BOUNCES];
        for (int k2 = 0; k2 < kStripes; k2++) stats->bounces += h_counters[C_BOUNCES_STRIPED + k2];
        stats->node_steps = h_counters[C_NODE_STEPS];
        stats->prim_tests = h_counters[C_PRIM_TESTS];
        stats->node_passes = h_counters[C_NODE_PASSES];
        stats->prim_passes = h_counters[C_PRIM_PASSES];
        stats->kernel_launches = launches;
        stats->waves = waves;
        stats->extend_launches = stats->shade_launches = stats->shadow_launches = waves;
        stats->regen_launches = waves + 1;
        stats->pool_paths = (uint64_t)capacity;
        stats->gpus_used = 1;
    }
    return LJ_OK;
}

// ---- query seam S2 through the wavefront kernels: a ray batch is loaded into the path pool the way k_regen / k_shade
// leave it and traced by the kernels lj_render launches (same grid, same cursors, same scheduling), so the ray-parity
// tests cover the refill / compaction / postponing logic and not only the plain per-thread loop of api.cu.
__global__ void k_pool_load_rays(PathPool pool, const lj_ray *rays, int m, int stride, int shadow, float shadow_eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool has = false;
    if (i < pool.capacity) {
        int k = i / stride;
        has = (i % stride) == 0 && k < m;
        V4 meta = mk4(0, 0, 0, 0), sh_d = mk4(0, 0, 0, -1.f);
        if (has) {
            lj_ray r = rays[k];
            meta = mk4(u2f((uint32_t)k), u2f(1u | kAlive | kOccupied), 0, 0);
            if (shadow) {
                // a path that continued: the NEE segment starts at ray_o (k_trace<1>); hit / ray_d are not read
                pool.ray_o[i] = mk4(r.org[0], r.org[1], r.org[2], shadow_eps);
                pool.ray_d[i] = mk4(0, 0, 1, 0);
                sh_d = mk4(r.dir[0], r.dir[1], r.dir[2], r.tfar);
                pool.sh_c[i] = mk4(1, 0, 0, 0);
            } else {
                pool.ray_o[i] = mk4(r.org[0], r.org[1], r.org[2], r.tnear);
                pool.ray_d[i] = mk4(r.dir[0], r.dir[1], r.dir[2], r.tfar);
            }
        }
        pool.meta[i] = meta;
        pool.sh_d[i] = sh_d;
        pool.rad[i] = mk4(0, 0, 0, 1.f);
        pool.hit[i] = mk4(0, 0, 0, u2f((uint32_t)kNoHit));
    }
    if (shadow) {
        unsigned mask = __ballot_sync(0xffffffffu, has);
        if (LJ_LANE() == 0 && i < pool.capacity) pool.sh_mask[i / LJ_WARP_WIDTH] = mask;
    }
}
__global__ void k_pool_store_hits(const LJ_GRID_CONSTANT DevScene sc, PathPool pool, int m, int stride, lj_hit *hits, uint8_t *occluded) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    size_t i = (size_t)k * stride;
    if (occluded) { occluded[k] = pool.rad[i].x == 0.f ? 1 : 0; return; }
    V4 h = pool.hit[i], o = pool.ray_o[i], d = pool.ray_d[i];
    Hit hit;
    hit.t = h.x; hit.u = h.y; hit.v = h.z; hit.prim = (int)f2u(h.w);
    if (hit.prim == kNoHit) hit.t = d.w;  // a miss reports the ray's tfar, like lj_trace_closest
    lj_hit out;
    hit_to_abi(sc, xyz(o), xyz(d), hit, out);
    hits[k] = out;
}

static int trace_pool_impl(lj_scene *s, const lj_ray *rays, int64_t n, const lj_trace_opts *opts, lj_hit *hits, uint8_t *occluded, double *kernel_ms) {
    const bool shadow = occluded != nullptr;
    const int stride = opts->slot_stride > 1 ? opts->slot_stride : 1;
    if (shadow)
        for (int64_t i = 0; i < n; i++)
            if (rays[i].tnear != s->dev.shadow_eps) { set_error("the shadow kernel starts every segment at the scene's shadow epsilon: rays[].tnear must equal it"); return LJ_ERR_INVALID; }
    long long capacity = opts->pool_paths > 0 ? opts->pool_paths : std::min<long long>((long long)n * stride, 1 << 22);
    capacity = std::max<long long>(1024, (capacity + 255) / 256 * 256);
    const int per_round = (int)((capacity + stride - 1) / stride);
    int r = ensure_pool(s, (int)capacity, false);
    if (r != LJ_OK) return r;
    r = ensure_render_buffers(s, s->dev.camera.width * s->dev.camera.height, false);
    if (r != LJ_OK) return r;
    WaveArgs a;
    memset(&a, 0, sizeof(a));
    a.pool = s->pool;
    fill_trace_args(s, a);
    lj_ray *d_rays = nullptr;
    lj_hit *d_hits = nullptr;
    uint8_t *d_occ = nullptr;
    cudaStream_t stream = s->stream;
    auto cleanup = [&]() { lj_dev_free(d_rays); lj_dev_free(d_hits); lj_dev_free(d_occ); };
    cudaError_t e = lj_dev_alloc((void **)&d_rays, (size_t)per_round * sizeof(lj_ray));
    if (e == cudaSuccess && !shadow) e = lj_dev_alloc((void **)&d_hits, (size_t)per_round * sizeof(lj_hit));
    if (e == cudaSuccess && shadow) e = lj_dev_alloc((void **)&d_occ, (size_t)per_round);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "ray batch allocation"); }
    const int saved_kernel = opts->kernel;
    double total_ms = 0;
    const int nb256 = ((int)capacity + 255) / 256;
    for (int64_t done = 0; done < n && e == cudaSuccess; done += per_round) {
        const int m = (int)std::min<int64_t>(per_round, n - done);
        e = cudaMemcpyAsync(d_rays, rays + done, (size_t)m * sizeof(lj_ray), cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) break;
        LJ_LAUNCH(k_pool_load_rays, nb256, 256, stream, s->pool, d_rays, m, stride, shadow ? 1 : 0, s->dev.shadow_eps);
        cudaMemsetAsync(s->d_cursors, 0, 4 * sizeof(unsigned int), stream);
        cudaMemsetAsync(s->d_counters, 0, sizeof(unsigned long long) * C_TOTAL, stream);
        cudaEventRecord(s->ev[0], stream);
        if (saved_kernel == LJ_TRACE_WAVEFRONT) {
            if (shadow) LJ_LAUNCH(k_trace_q<1>, s->geom.q_blocks, kQWarps * LJ_WARP_WIDTH, stream, s->dev, a);
            else LJ_LAUNCH(k_trace_q<0>, s->geom.q_blocks, kQWarps * LJ_WARP_WIDTH, stream, s->dev, a);
        } else {
            if (shadow) LJ_LAUNCH(k_trace<1>, s->geom.trace_blocks, 128, stream, s->dev, a);
            else LJ_LAUNCH(k_trace<0>, s->geom.trace_blocks, 128, stream, s->dev, a);
        }
        cudaEventRecord(s->ev[1], stream);
        LJ_LAUNCH(k_pool_store_hits, (m + 255) / 256, 256, stream, s->dev, s->pool, m, stride, d_hits, d_occ);
        if (shadow) e = cudaMemcpyAsync(occluded + done, d_occ, (size_t)m, cudaMemcpyDeviceToHost, stream);
        else e = cudaMemcpyAsync(hits + done, d_hits, (size_t)m * sizeof(lj_hit), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        float ms = 0;
        if (e == cudaSuccess) { cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); total_ms += ms; }
    }
    cleanup();
    if (e != cudaSuccess) return cuda_fail(e, "wavefront ray batch");
    if (kernel_ms) *kernel_ms = total_ms;
    return LJ_OK;
}

// ---- the volpath NEE segment walk (homework2.tex:459-510, 771-810) through the persistent walk kernels, and as a plain
// one-thread-per-walk loop over the same step functions: the two must agree bit for bit (same hits by the tie policy,
// same PCG draws in the same order), which is what the parity test of k_trace<2> / k_trace<3> asserts.
__global__ void __launch_bounds__(128) k_walk_plain(const LJ_GRID_CONSTANT DevScene sc, const lj_walk_query *q, int n, uint64_t seed, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lj_walk_query in = q[i];
    NeeWalk wk;
    wk.pc = mk3(in.origin[0], in.origin[1], in.origin[2]);
    wk.pl = mk3(in.light_point[0], in.light_point[1], in.light_point[2]);
    wk.dir = normalize(wk.pl - wk.pc);
    wk.T_light = mk3(1); wk.p_nee = mk3(1); wk.p_dir = mk3(1);
    wk.c = mk3(in.c[0], in.c[1], in.c[2]);
    wk.pdf_nee = in.pdf_nee; wk.pdf_dir = in.pdf_dir;
    wk.medium = in.medium_id;
    wk.budget = in.budget < 0 ? kWalkNoBudget : (uint32_t)in.budget;
    wk.shadow_bounces = 0;
    wk.rng = walk_rng((uint64_t)i, in.seed, seed);
    V3 contrib = mk3(0);
    for (int guard = 0; guard < 1 << 16; guard++) {
        float tn, tf;
        nee_walk_segment(sc, wk, tn, tf);
        Hit hit;
        trace8<false>(sc.nodes8, sc.prims, wk.pc, wk.dir, tn, tf, hit);
        float next_t = nee_walk_next_t(wk, hit);
        if (wk.medium >= 0) ratio_track(sc.media[wk.medium], wk.pc, wk.dir, tf, next_t, sc.options.max_null_collisions, wk.rng, wk.T_light, wk.p_nee, wk.p_dir);
        if (nee_walk_decide(sc, wk, hit, contrib)) break;
    }
    out[3 * i] = contrib.x; out[3 * i + 1] = contrib.y; out[3 * i + 2] = contrib.z;
}
__global__ void k_pool_load_walks(PathPool pool, const lj_walk_query *q, int m, int stride, int first) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool has = false;
    if (i < pool.capacity) {
        int k = i / stride;
        has = (i % stride) == 0 && k < m;
        V4 meta = mk4(0, 0, 0, 0), sh_d = mk4(0, 0, 0, -1.f);
        if (has) {
            lj_walk_query in = q[k];
            V3 o = mk3(in.origin[0], in.origin[1], in.origin[2]), pl = mk3(in.light_point[0], in.light_point[1], in.light_point[2]);
            V3 d = normalize(pl - o);
            meta = mk4(u2f((uint32_t)(first + k)), u2f(1u | kOccupied), 0, 0);  // (pixel = walk index: the walk's RNG stream)
            sh_d = mk4(d, in.pdf_dir);
            uint32_t budget = in.budget < 0 ? kWalkNoBudget : (uint32_t)in.budget;
            pool.sh_o[i] = mk4(o, u2f((uint32_t)((in.medium_id + 1) & 0xffff) | (budget << 16)));
            pool.sh_pl[i] = mk4(pl, u2f(in.seed));
            pool.sh_c[i] = mk4(in.c[0], in.c[1], in.c[2], in.pdf_nee);
            pool.aux[i] = mk4(1, 0, u2f(0u), u2f((uint32_t)in.medium_id));
        }
        pool.meta[i] = meta;
        pool.sh_d[i] = sh_d;
        pool.rad[i] = mk4(0, 0, 0, 1.f);
    }
    unsigned mask = __ballot_sync(0xffffffffu, has);
    if (LJ_LANE() == 0 && i < pool.capacity) pool.sh_mask[i / LJ_WARP_WIDTH] = mask;
}
__global__ void k_pool_store_walks(PathPool pool, int m, int stride, float *out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    V4 r = pool.rad[(size_t)k * stride];
    out[3 * k] = r.x; out[3 * k + 1] = r.y; out[3 * k + 2] = r.z;
}

static int walk_batch_impl(lj_scene *s, const lj_walk_query *q, int64_t n, const lj_trace_opts *opts, float *out, double *kernel_ms) {
    const int kernel = opts ? opts->kernel : LJ_TRACE_PLAIN;
    const int stride = opts && opts->slot_stride > 1 ? opts->slot_stride : 1;
    for (int64_t i = 0; i < n; i++)
        if (q[i].medium_id < -1 || q[i].medium_id >= s->dev.num_media || q[i].budget >= (int)kWalkNoBudget) { set_error("walk query out of range"); return LJ_ERR_INVALID; }
    if (n > (int64_t)1 << 30) { set_error("too many walks in one batch"); return LJ_ERR_INVALID; }
    lj_walk_query *d_q = nullptr;
    float *d_out = nullptr;
    cudaStream_t stream = s->stream;
    auto cleanup = [&]() { lj_dev_free(d_q); lj_dev_free(d_out); };
    cudaError_t e = lj_dev_alloc((void **)&d_q, (size_t)n * sizeof(lj_walk_query));
    if (e == cudaSuccess) e = lj_dev_alloc((void **)&d_out, (size_t)n * 3 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_q, q, (size_t)n * sizeof(lj_walk_query), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "walk batch allocation"); }
    double total_ms = 0;
    if (kernel == LJ_TRACE_PLAIN) {
        cudaEventRecord(s->ev[0], stream);
        LJ_LAUNCH(k_walk_plain, (int)((n + 127) / 128), 128, stream, s->dev, d_q, (int)n, kPcgDefaultSeed, d_out);
        cudaEventRecord(s->ev[1], stream);
        e = cudaStreamSynchronize(stream);
        float ms = 0;
        if (e == cudaSuccess) { cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); total_ms = ms; }
    } else {
        long long capacity = opts->pool_paths > 0 ? opts->pool_paths : std::min<long long>((long long)n * stride, 1 << 22);
        capacity = std::max<long long>(1024, (capacity + 255) / 256 * 256);
        const int per_round = (int)((capacity + stride - 1) / stride);
        int r = ensure_pool(s, (int)capacity, true);
        if (r == LJ_OK) r = ensure_render_buffers(s, s->dev.camera.width * s->dev.camera.height, false);
        if (r != LJ_OK) { cleanup(); return r; }
        WaveArgs a;
        memset(&a, 0, sizeof(a));
        a.pool = s->pool;
        a.rp.spp_total = 1;
        a.rp.seed = kPcgDefaultSeed;
        fill_trace_args(s, a);
        const int nb256 = ((int)capacity + 255) / 256;
        for (int64_t done = 0; done < n && e == cudaSuccess; done += per_round) {
            const int m = (int)std::min<int64_t>(per_round, n - done);
            LJ_LAUNCH(k_pool_load_walks, nb256, 256, stream, s->pool, d_q + done, m, stride, (int)done);
            cudaMemsetAsync(s->d_cursors, 0, 4 * sizeof(unsigned int), stream);
            cudaMemsetAsync(s->d_counters, 0, sizeof(unsigned long long) * C_TOTAL, stream);
            cudaEventRecord(s->ev[0], stream);
            if (kernel == LJ_TRACE_WALK_STAGED) launch_walk_staged(s, a, stream, opts->walk_rounds > 0 ? std::min(opts->walk_rounds, 64) : kWalkRounds);
            else if (kernel == LJ_TRACE_WALK_STEP) LJ_LAUNCH(k_trace<3>, s->geom.step_blocks, 128, stream, s->dev, a);
            else LJ_LAUNCH(k_trace<2>, s->geom.walk_blocks, 128, stream, s->dev, a);
            cudaEventRecord(s->ev[1], stream);
            LJ_LAUNCH(k_pool_store_walks, (m + 255) / 256, 256, stream, s->pool, m, stride, d_out + 3 * done);
            e = cudaStreamSynchronize(stream);
            if (e == cudaSuccess) e = cudaGetLastError();
            float ms = 0;
            if (e == cudaSuccess) { cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]); total_ms += ms; }
        }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return cuda_fail(e, "walk batch");
    if (kernel_ms) *kernel_ms = total_ms;
    return LJ_OK;
}

}  // namespace lj

using namespace lj;

extern "C" int lj_trace_closest_ex(lj_scene *s, const lj_ray *rays, int64_t n, const lj_trace_opts *opts, lj_hit *hits, double *kernel_ms) {
    if (!s || !rays || !hits || n < 0) { set_error("invalid argument"); return LJ_ERR_INVALID; }
    if (!opts || opts->kernel == LJ_TRACE_PLAIN) return lj_trace_closest(s, rays, n, hits, kernel_ms);
    if (opts->kernel != LJ_TRACE_WAVEFRONT && opts->kernel != LJ_TRACE_WAVEFRONT_LANE) { set_error("unknown trace kernel"); return LJ_ERR_INVALID; }
    if (n == 0) return LJ_OK;
    DeviceGuard guard(s->device);
    return trace_pool_impl(s, rays, n, opts, hits, nullptr, kernel_ms);
}

extern "C" int lj_trace_any_ex(lj_scene *s, const lj_ray *rays, int64_t n, const lj_trace_opts *opts, uint8_t *occluded, double *kernel_ms) {
    if (!s || !rays || !occluded || n < 0) { set_error("invalid argument"); return LJ_ERR_INVALID; }
    if (!opts || opts->kernel == LJ_TRACE_PLAIN) return lj_trace_any(s, rays, n, occluded, kernel_ms);
    if (opts->kernel != LJ_TRACE_WAVEFRONT && opts->kernel != LJ_TRACE_WAVEFRONT_LANE) { set_error("unknown trace kernel"); return LJ_ERR_INVALID; }
    if (n == 0) return LJ_OK;
    DeviceGuard guard(s->device);
    return trace_pool_impl(s, rays, n, opts, nullptr, occluded, kernel_ms);
}

extern "C" int lj_nee_walk_batch(lj_scene *s, const lj_walk_query *q, int64_t n, const lj_trace_opts *opts, float *contribution_rgb, double *kernel_ms) {
    if (!s || !q || !contribution_rgb || n < 0) { set_error("invalid argument"); return LJ_ERR_INVALID; }
    if (opts && opts->kernel != LJ_TRACE_PLAIN && opts->kernel != LJ_TRACE_WALK_WHOLE && opts->kernel != LJ_TRACE_WALK_STEP && opts->kernel != LJ_TRACE_WALK_STAGED) { set_error("unknown walk kernel"); return LJ_ERR_INVALID; }
    if (n == 0) return LJ_OK;
    DeviceGuard guard(s->device);
    return walk_batch_impl(s, q, n, opts, contribution_rgb, kernel_ms);
}

extern "C" int lj_render_device(lj_scene *s, const lj_render_opts *opts, float *d_out_rgb, void *stream, lj_stats *stats) {
    if (!s || !d_out_rgb) { set_error("null argument"); return LJ_ERR_INVALID; }
    if (opts && opts->variance_out) { set_error("variance_out needs lj_render (host buffers)"); return LJ_ERR_INVALID; }
    if (opts && opts->num_gpus > 1) { set_error("lj_render_device renders on the scene's primary device: use lj_render for the multi-GPU split"); return LJ_ERR_INVALID; }
    DeviceGuard guard(s->device);
    return render_impl(s, opts, d_out_rgb, nullptr, stream ? (cudaStream_t)stream : s->stream, stats);
}

// ---- multi-GPU render inside the library -------------------------------------------------------------
// The scene is replicated on every device of lj_init (lj_scene_create); lj_render(num_gpus = G) gives each of G devices
// a share of the samples -- a contiguous block of the per-pixel sample indices (LJ_SPLIT_SPP) or every G-th 8x4 pixel
// tile (LJ_SPLIT_TILES) -- renders the shares concurrently, one host thread per GPU, and sums the per-GPU films on the
// primary device.  Two reductions: LJ_REDUCE_P2P, one kernel on the primary device that LOADS the other devices' films
// through their peer mappings (NVLink) while it resolves -- reduce and resolve fused, no staging copy; LJ_REDUCE_NCCL,
// ncclReduce of the fp32 films to the primary device, then the usual resolve.
#if !defined(LJ_HOSTSIM)
constexpr int kMaxGpus = 16;
struct FilmSet { const float *film[kMaxGpus]; const float *film_sq[kMaxGpus]; int n; };

__global__ void k_resolve_multi(const LJ_GRID_CONSTANT FilmSet fs, int npix, float inv_n, int normalize, float *out, float *var_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 f = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    for (int g = 0; g < fs.n; g++) {  // (peer loads for g > 0: 16 bytes per pixel per GPU over NVLink)
        float4 v = reinterpret_cast<const float4 *>(fs.film[g])[i];
        f.x += v.x; f.y += v.y; f.z += v.z; f.w += v.w;
        if (var_out) { float4 w = reinterpret_cast<const float4 *>(fs.film_sq[g])[i]; q.x += w.x; q.y += w.y; q.z += w.z; }
    }
    float s = normalize ? inv_n : 1.f;
    out[3 * i] = f.x * s; out[3 * i + 1] = f.y * s; out[3 * i + 2] = f.z * s;
    if (var_out) {
        float n = f.w, v[3] = {0, 0, 0};
        if (n > 1) {
            v[0] = fmaxf(q.x - f.x * f.x / n, 0.f) / (n - 1) / n;
            v[1] = fmaxf(q.y - f.y * f.y / n, 0.f) / (n - 1) / n;
            v[2] = fmaxf(q.z - f.z * f.z / n, 0.f) / (n - 1) / n;
        }
        var_out[3 * i] = v[0]; var_out[3 * i + 1] = v[1]; var_out[3 * i + 2] = v[2];
    }
}

// NCCL is loaded at run time, and only when a render asks for it: libljb200.so must load next to any NCCL the host
// process already carries (PyTorch bundles its own) without pulling a second copy in at link time.
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::vector<void *> comms;
    std::vector<int> devices;
};
static std::mutex g_nccl_mutex;
static NcclApi g_nccl;

static int nccl_prepare(const std::vector<int> &devices) {
    if (!g_nccl.lib) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { set_error(std::string("LJ_REDUCE_NCCL: cannot load libnccl.so.2: ") + dlerror()); return LJ_ERR_UNSUPPORTED; }
        g_nccl.lib = h;
        *(void **)&g_nccl.CommInitAll = dlsym(h, "ncclCommInitAll");
        *(void **)&g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
        *(void **)&g_nccl.GroupStart = dlsym(h, "ncclGroupStart");
        *(void **)&g_nccl.GroupEnd = dlsym(h, "ncclGroupEnd");
        *(void **)&g_nccl.Reduce = dlsym(h, "ncclReduce");
        *(void **)&g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
        if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.Reduce) {
            set_error("LJ_REDUCE_NCCL: libnccl lacks the expected entry points");
            return LJ_ERR_UNSUPPORTED;
        }
    }
    if (g_nccl.devices != devices) {
        for (void *c : g_nccl.comms) g_nccl.CommDestroy(c);
        g_nccl.comms.assign(devices.size(), nullptr);
        int r = g_nccl.CommInitAll(g_nccl.comms.data(), (int)devices.size(), devices.data());
        if (r != 0) {
            g_nccl.comms.clear(); g_nccl.devices.clear();
            set_error(std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
            return LJ_ERR_CUDA;
        }
        g_nccl.devices = devices;
    }
    return LJ_OK;
}

static int render_multi(lj_scene *s, const lj_render_opts *opts_in, int G, float *out_rgb, lj_stats *stats) {
    lj_render_opts base;
    memset(&base, 0, sizeof(base));
    if (opts_in) base = *opts_in;
    const DevScene &sc = s->dev;
    if (sc.options.integrator != LJ_INT_PATH && sc.options.integrator != LJ_INT_VOLPATH) G = 1;  // one launch: nothing to split
    if (base.tile_stride > 1) { set_error("tile_stride / tile_offset and num_gpus > 1 exclude each other"); return LJ_ERR_INVALID; }
    const int spp = base.spp > 0 ? base.spp : sc.options.spp;
    int sb = base.sample_begin, se = base.sample_end;
    if (sb == 0 && se == 0) se = spp;
    if (sb < 0 || se > spp || sb >= se) { set_error("bad sample range"); return LJ_ERR_INVALID; }
    int split = base.split;
    if (split == LJ_SPLIT_AUTO) split = (se - sb) >= G ? LJ_SPLIT_SPP : LJ_SPLIT_TILES;
    if (split == LJ_SPLIT_SPP && (se - sb) < G) G = se - sb;
    std::vector<lj_scene *> sc_g(G);
    sc_g[0] = s;
    for (int g = 1; g < G; g++) sc_g[g] = s->replicas[g - 1];
    const int npix = sc.camera.width * sc.camera.height;
    const bool want_var = base.variance_out != nullptr;
    // peer access from the primary device to the replicas' films (once per scene)
    if (!s->peer_checked) {
        s->peer_checked = true;
        s->peer_ok = true;
        for (lj_scene *r : s->replicas) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, s->device, r->device) != cudaSuccess || !can) { s->peer_ok = false; break; }
            cudaError_t e = cudaDeviceEnablePeerAccess(r->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { s->peer_ok = false; break; }
        }
        cudaGetLastError();
    }
    int reduce = base.reduce;
    if (reduce == LJ_REDUCE_AUTO) reduce = s->peer_ok ? LJ_REDUCE_P2P : LJ_REDUCE_NCCL;
    if (reduce == LJ_REDUCE_P2P && !s->peer_ok) { set_error("LJ_REDUCE_P2P: the primary device cannot access its peers"); return LJ_ERR_UNSUPPORTED; }
    if (reduce != LJ_REDUCE_P2P && reduce != LJ_REDUCE_NCCL) { set_error("unknown reduce mode"); return LJ_ERR_INVALID; }
    // ---- the shares, one host thread per GPU
    std::vector<lj_stats> st(G);
    std::vector<int> codes(G, LJ_OK);
    std::vector<std::string> errors(G);
    std::vector<float *> d_dummy_var(G, nullptr);
    auto work = [&](int g) {
        lj_scene *sg = sc_g[g];
        DeviceGuard guard(sg->device);
        lj_render_opts o = base;
        o.spp = spp;
        o.normalize = 0;
        o.variance_out = nullptr;
        o.num_gpus = 1;
        if (split == LJ_SPLIT_SPP) {
            const int total = se - sb, per = total / G, extra = total % G;
            o.sample_begin = sb + g * per + std::min(g, extra);
            o.sample_end = o.sample_begin + per + (g < extra ? 1 : 0);
            o.tile_stride = 0; o.tile_offset = 0;
        } else {
            o.sample_begin = sb; o.sample_end = se;
            o.tile_stride = G; o.tile_offset = g;
        }
        // render_impl accumulates into sg->d_film (and d_film_sq when a variance buffer is named); no resolve here
        float *var_flag = nullptr;
        if (want_var) {
            if (lj_dev_alloc((void **)&d_dummy_var[g], 16) != cudaSuccess) { codes[g] = LJ_ERR_CUDA; errors[g] = "allocation failed"; return; }
            var_flag = d_dummy_var[g];
        }
        codes[g] = render_impl(sg, &o, nullptr, var_flag, sg->stream, &st[g]);
        if (codes[g] != LJ_OK) errors[g] = lj_last_error();
    };
    {
        std::vector<std::thread> threads;
        for (int g = 1; g < G; g++) threads.emplace_back(work, g);
        work(0);
        for (auto &t : threads) t.join();
    }
    for (int g = 0; g < G; g++) if (d_dummy_var[g]) { DeviceGuard guard(sc_g[g]->device); lj_dev_free(d_dummy_var[g]); }
    for (int g = 0; g < G; g++)
        if (codes[g] != LJ_OK) { set_error("GPU " + std::to_string(sc_g[g]->device) + ": " + errors[g]); return codes[g]; }
    // ---- reduce on the primary device
    float *d_out = nullptr, *d_var = nullptr;
    LJ_CUDA(lj_dev_alloc((void **)&d_out, (size_t)npix * 3 * sizeof(float)));
    if (want_var) {
        cudaError_t e = lj_dev_alloc((void **)&d_var, (size_t)npix * 3 * sizeof(float));
        if (e != cudaSuccess) { lj_dev_free(d_out); return cuda_fail(e, "variance buffer allocation"); }
    }
    cudaStream_t stream = s->stream;
    const float inv_n = 1.f / (float)(se - sb);
    cudaEventRecord(s->ev[2], stream);
    int rc = LJ_OK;
    if (reduce == LJ_REDUCE_P2P) {
        FilmSet fs;
        fs.n = G;
        for (int g = 0; g < G; g++) { fs.film[g] = sc_g[g]->d_film; fs.film_sq[g] = sc_g[g]->d_film_sq; }
        LJ_LAUNCH(k_resolve_multi, (npix + 255) / 256, 256, stream, fs, npix, inv_n, base.normalize, d_out, d_var);
    } else {
        std::lock_guard<std::mutex> lock(g_nccl_mutex);
        std::vector<int> devs(G);
        for (int g = 0; g < G; g++) devs[g] = sc_g[g]->device;
        rc = nccl_prepare(devs);
        if (rc == LJ_OK) {
            for (int pass = 0; pass < (want_var ? 2 : 1) && rc == LJ_OK; pass++) {
                g_nccl.GroupStart();
                for (int g = 0; g < G; g++) {
                    float *buf = pass == 0 ? sc_g[g]->d_film : sc_g[g]->d_film_sq;
                    int r = g_nccl.Reduce(buf, buf, (size_t)npix * 4, 7 /*ncclFloat32*/, 0 /*ncclSum*/, 0, g_nccl.comms[g], sc_g[g]->stream);
                    if (r != 0) { set_error(std::string("ncclReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error")); rc = LJ_ERR_CUDA; }
                }
                g_nccl.GroupEnd();
            }
            if (rc == LJ_OK) LJ_LAUNCH(k_resolve, (npix + 255) / 256, 256, stream, s->d_film, want_var ? s->d_film_sq : nullptr, npix, inv_n, base.normalize, d_out, d_var);
        }
    }
    cudaEventRecord(s->ev[3], stream);
    if (rc == LJ_OK) {
        cudaError_t e = cudaStreamSynchronize(stream);
        for (int g = 1; g < G && e == cudaSuccess; g++) { DeviceGuard guard(sc_g[g]->device); e = cudaStreamSynchronize(sc_g[g]->stream); }
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(out_rgb, d_out, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && want_var) e = cudaMemcpy(base.variance_out, d_var, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "film reduction");
    }
    if (rc == LJ_OK && stats) {
        memset(stats, 0, sizeof(*stats));
        for (int g = 0; g < G; g++) {
            const lj_stats &x = st[g];
            stats->render_ms = std::max(stats->render_ms, x.render_ms);
            stats->extend_ms = std::max(stats->extend_ms, x.extend_ms); stats->shadow_ms = std::max(stats->shadow_ms, x.shadow_ms);
            stats->shade_ms = std::max(stats->shade_ms, x.shade_ms); stats->regen_ms = std::max(stats->regen_ms, x.regen_ms);
            stats->samples += x.samples; stats->closest_rays += x.closest_rays; stats->shadow_rays += x.shadow_rays; stats->bounces += x.bounces;
            stats->kernel_launches += x.kernel_launches; stats->waves = std::max(stats->waves, x.waves);
            stats->extend_launches += x.extend_launches; stats->shadow_launches += x.shadow_launches;
            stats->shade_launches += x.shade_launches; stats->regen_launches += x.regen_launches;
            stats->node_steps += x.node_steps; stats->prim_tests += x.prim_tests; stats->node_passes += x.node_passes; stats->prim_passes += x.prim_passes;
            stats->pool_paths += x.pool_paths;
        }
        stats->kernel_launches += 1;
        stats->gpus_used = G;
        float ms = 0;
        cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]);
        stats->reduce_ms = ms;
    }
    lj_dev_free(d_out);
    if (d_var) lj_dev_free(d_var);
    return rc;
}
#endif

extern "C" int lj_render(lj_scene *s, const lj_render_opts *opts, float *out_rgb, lj_stats *stats) {
    if (!s || !out_rgb) { set_error("null argument"); return LJ_ERR_INVALID; }
    DeviceGuard guard(s->device);
#if !defined(LJ_HOSTSIM)
    {   // GPUs for this call: 0 = every device lj_init named (the scene's replicas), n = at most n of them
        const int have = 1 + (int)s->replicas.size();
        int want = opts ? opts->num_gpus : 1;
        if (want < 0 || want > kMaxGpus) { set_error("bad num_gpus"); return LJ_ERR_INVALID; }
        if (want > have) { set_error("num_gpus exceeds the devices given to lj_init before this scene was created"); return LJ_ERR_INVALID; }
        const int G = want == 0 ? have : want;
        if (G > 1) return render_multi(s, opts, G, out_rgb, stats);
    }
#endif
    int npix = s->dev.camera.width * s->dev.camera.height;
    const bool host_prof = tuning().host_prof;
    auto t0 = std::chrono::steady_clock::now();
    float *d_out = nullptr, *d_var = nullptr;
    LJ_CUDA(lj_dev_alloc((void **)&d_out, (size_t)npix * 3 * sizeof(float)));
    bool want_var = opts && opts->variance_out;
    if (want_var) {
        cudaError_t e = lj_dev_alloc((void **)&d_var, (size_t)npix * 3 * sizeof(float));
        if (e != cudaSuccess) { lj_dev_free(d_out); return cuda_fail(e, "variance buffer allocation"); }
    }
    auto t1 = std::chrono::steady_clock::now();
    int r = render_impl(s, opts, d_out, d_var, s->stream, stats);
    auto t2 = std::chrono::steady_clock::now();
    if (r == LJ_OK) {
        cudaError_t e = cudaMemcpy(out_rgb, d_out, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && want_var) e = cudaMemcpy(opts->variance_out, d_var, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) r = cuda_fail(e, "framebuffer download");
    }
    auto t3 = std::chrono::steady_clock::now();
    lj_dev_free(d_out);
    if (d_var) lj_dev_free(d_var);
    if (host_prof) {
        auto ms_ = [](auto a_, auto b_) { return std::chrono::duration<double, std::milli>(b_ - a_).count(); };
        fprintf(stderr, "lj_render: out alloc %.2f ms, render_impl %.2f ms, download %.2f ms, free %.2f ms\n", ms_(t0, t1), ms_(t1, t2), ms_(t2, t3),
                ms_(t3, std::chrono::steady_clock::now()));
    }
    return r;
}
