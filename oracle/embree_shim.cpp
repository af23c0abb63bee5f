// oracle/embree_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the 21 Embree 4.3.0 C entry points that lajolla calls, so that the
// UNMODIFIED reference sources under /root/reference/src link and run here (the shipped
// libembree4.so.4 is absent: /root/reference/.MISSING_LARGE_BLOBS).  Embree 4.3.0 is pinned by
// /root/reference/embree/include/embree4/rtcore_config.h:10-14.  Call sites this file serves:
//   scene.cpp:20-31,58            rtcNewScene / BuildQuality / SceneFlags / Commit / GetSceneBounds / Release
//   shapes/triangle_mesh.inl:2-20 rtcNewGeometry / Attach / SetNewGeometryBuffer / VertexAttributeCount / Commit / Release
//   shapes/sphere.inl:152-160     user geometry: PrimitiveCount / UserData / Bounds / Intersect / Occluded callbacks
//   intersection.cpp:32,83        rtcIntersect1 / rtcOccluded1
//   main.cpp:30,48                rtcNewDevice / rtcReleaseDevice
//
// Published Embree semantics restated here (API manual for RTC_GEOMETRY_TYPE_TRIANGLE and the
// RTC_SCENE_FLAG_ROBUST Pluecker intersector):
//   * geomID = attach order (0,1,2,...).
//   * triangle hit: p = (1-u-v) v0 + u v1 + v v2 ; Ng = (v1-v0) x (v2-v0), unnormalised.
//   * edge functions U,V,W from Pluecker coordinates relative to the ray origin; a hit needs all
//     three of one sign (zero included), den != 0 and tnear <= t <= tfar; then ray.tfar = t.
//   * rtcOccluded1 sets ray.tfar = -inf on any hit.
//   * user geometry callbacks are invoked with N = 1, valid[0] = -1.
// The triangle test is evaluated in DOUBLE on the float-rounded inputs: it is the ground truth
// for the "t within 1e-5 relative" ray-parity test of the CUDA traversal.
// Edge-tie policy (documented): among hits with exactly equal t the lowest
// (geomID, primID) wins.
//
// Parity status: "parity unpinned" against real Embree (no binary available here); pinned
// against the reference's own tests/intersection.cpp (one ray / one triangle) only.

#include <embree4/rtcore.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int i = 0; i < 3; i++) { lo[i] = FLT_MAX; hi[i] = -FLT_MAX; }
    }
    void grow(const Box &b) {
        for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], b.lo[i]); hi[i] = std::max(hi[i], b.hi[i]); }
    }
    void grow(const float *p) {
        for (int i = 0; i < 3; i++) { lo[i] = std::min(lo[i], p[i]); hi[i] = std::max(hi[i], p[i]); }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct DeviceImpl {
    std::atomic<int> refs{1};
};

struct GeomImpl {
    std::atomic<int> refs{1};
    RTCGeometryType type;
    // triangle buffers (owned, as rtcSetNewGeometryBuffer allocates them)
    std::vector<unsigned char> vbuf, ibuf;
    size_t vstride = 0, vcount = 0, istride = 0, icount = 0;
    // user geometry
    unsigned int user_prims = 0;
    void *user_ptr = nullptr;
    RTCBoundsFunction bounds_fn = nullptr;
    void *bounds_user = nullptr;
    RTCIntersectFunctionN intersect_fn = nullptr;
    RTCOccludedFunctionN occluded_fn = nullptr;

    const float *vertex(size_t i) const { return (const float *)(vbuf.data() + i * vstride); }
    const unsigned *tri(size_t i) const { return (const unsigned *)(ibuf.data() + i * istride); }
    size_t num_prims() const { return type == RTC_GEOMETRY_TYPE_TRIANGLE ? icount : user_prims; }
};

struct PrimRef {
    unsigned geom, prim;
    Box box;
    float c[3];
};

struct Node {
    Box box;
    int left;   // internal: index of left child (right = left+1); leaf: first prim
    int count;  // 0 for internal
};

struct SceneImpl {
    std::atomic<int> refs{1};
    std::vector<GeomImpl *> geoms;
    std::vector<PrimRef> prims;
    std::vector<Node> nodes;
    Box bounds;
    bool committed = false;
};

std::atomic<unsigned long long> g_closest{0}, g_any{0};
// time-stamp-counter ticks spent inside rtcIntersect1 / rtcOccluded1, all threads (flushed every 1024 calls per thread):
// the shim's share of the CPU render time, reported next to every "x CPU" figure (bench.py cpu_baseline.shim_share)
std::atomic<unsigned long long> g_ticks{0};
static inline unsigned long long shim_rdtsc() {
#if defined(__x86_64__) || defined(__i386__)
    unsigned lo, hi;
    __asm__ __volatile__("rdtsc" : "=a"(lo), "=d"(hi));
    return ((unsigned long long)hi << 32) | lo;
#else
    return 0;
#endif
}

inline void release(GeomImpl *g) {
    if (g->refs.fetch_sub(1) == 1) delete g;
}

// ---------------------------------------------------------------- BVH build (binned SAH)
struct Builder {
    std::vector<PrimRef> &prims;
    std::vector<Node> &nodes;

    void build() {
        nodes.clear();
        nodes.reserve(prims.size() * 2 + 1);
        nodes.push_back(Node{});
        subdivide(0, 0, (int)prims.size());
    }

    void subdivide(int ni, int begin, int end) {
        Box box, cbox;
        box.reset();
        cbox.reset();
        for (int i = begin; i < end; i++) {
            box.grow(prims[i].box);
            cbox.grow(prims[i].c);
        }
        nodes[ni].box = box;
        int n = end - begin;
        if (n <= 2) {
            nodes[ni].left = begin;
            nodes[ni].count = n;
            return;
        }
        constexpr int NB = 16;
        float best = FLT_MAX;
        int best_axis = -1, best_split = -1;
        for (int ax = 0; ax < 3; ax++) {
            float lo = cbox.lo[ax], hi = cbox.hi[ax];
            if (!(hi > lo)) continue;
            Box bb[NB];
            int bc[NB];
            for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
            float k = NB * (1 - 1e-6f) / (hi - lo);
            for (int i = begin; i < end; i++) {
                int b = std::min(NB - 1, std::max(0, (int)(k * (prims[i].c[ax] - lo))));
                bb[b].grow(prims[i].box);
                bc[b]++;
            }
            float la[NB], ra[NB];
            int lc[NB], rc[NB];
            Box acc;
            acc.reset();
            int cnt = 0;
            for (int b = 0; b < NB - 1; b++) {
                acc.grow(bb[b]);
                cnt += bc[b];
                la[b] = acc.half_area();
                lc[b] = cnt;
            }
            acc.reset();
            cnt = 0;
            for (int b = NB - 1; b > 0; b--) {
                acc.grow(bb[b]);
                cnt += bc[b];
                ra[b - 1] = acc.half_area();
                rc[b - 1] = cnt;
            }
            for (int b = 0; b < NB - 1; b++) {
                if (lc[b] == 0 || rc[b] == 0) continue;
                float cost = la[b] * lc[b] + ra[b] * rc[b];
                if (cost < best) { best = cost; best_axis = ax; best_split = b; }
            }
        }
        int mid;
        if (best_axis < 0) {
            if (n <= 8) {
                nodes[ni].left = begin;
                nodes[ni].count = n;
                return;
            }
            mid = (begin + end) / 2;  // coincident centroids: median split
        } else {
            float leaf_cost = box.half_area() * n;
            if (n <= 4 && leaf_cost <= best + box.half_area() * 0.5f) {
                nodes[ni].left = begin;
                nodes[ni].count = n;
                return;
            }
            float lo = cbox.lo[best_axis], hi = cbox.hi[best_axis];
            float k = NB * (1 - 1e-6f) / (hi - lo);
            auto it = std::partition(prims.begin() + begin, prims.begin() + end, [&](const PrimRef &p) {
                int b = std::min(NB - 1, std::max(0, (int)(k * (p.c[best_axis] - lo))));
                return b <= best_split;
            });
            mid = (int)(it - prims.begin());
            if (mid == begin || mid == end) mid = (begin + end) / 2;
        }
        int l = (int)nodes.size();
        nodes.push_back(Node{});
        nodes.push_back(Node{});
        nodes[ni].left = l;
        nodes[ni].count = 0;
        subdivide(l, begin, mid);
        subdivide(l + 1, mid, end);
    }
};

// ---------------------------------------------------------------- traversal
struct RayD {
    double o[3], d[3], inv[3];
    double tnear, tfar;
};

// Conservative slab test in double on float boxes.
inline bool hit_box(const Box &b, const RayD &r, double tfar, double &tmin_out) {
    double t0 = r.tnear, t1 = tfar;
    for (int a = 0; a < 3; a++) {
        double ta = ((double)b.lo[a] - r.o[a]) * r.inv[a];
        double tb = ((double)b.hi[a] - r.o[a]) * r.inv[a];
        if (ta > tb) std::swap(ta, tb);
        // NaN (0 * inf) must not reject: comparisons below are false for NaN
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
    }
    tmin_out = t0;
    // pad: boxes are float, ray math double; widen by a relative epsilon
    return t0 <= t1 * (1 + 1e-9) + 1e-12;
}

struct TriHit {
    double t, u, v;
    double ng[3];
};

// Pluecker edge-function triangle test (Embree ROBUST convention), double precision.
inline bool hit_triangle(const float *a, const float *b, const float *c, const RayD &r, double tfar, TriHit &h) {
    double v0[3], v1[3], v2[3];
    for (int i = 0; i < 3; i++) {
        v0[i] = (double)a[i] - r.o[i];
        v1[i] = (double)b[i] - r.o[i];
        v2[i] = (double)c[i] - r.o[i];
    }
    double e0[3], e1[3], e2[3];
    for (int i = 0; i < 3; i++) {
        e0[i] = v2[i] - v0[i];
        e1[i] = v0[i] - v1[i];
        e2[i] = v1[i] - v2[i];
    }
    auto cross = [](const double *x, const double *y, double *o) {
        o[0] = x[1] * y[2] - x[2] * y[1];
        o[1] = x[2] * y[0] - x[0] * y[2];
        o[2] = x[0] * y[1] - x[1] * y[0];
    };
    auto dot = [](const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
    double s[3], cr[3];
    for (int i = 0; i < 3; i++) s[i] = v2[i] + v0[i];
    cross(e0, s, cr);
    double U = dot(cr, r.d);
    for (int i = 0; i < 3; i++) s[i] = v0[i] + v1[i];
    cross(e1, s, cr);
    double V = dot(cr, r.d);
    for (int i = 0; i < 3; i++) s[i] = v1[i] + v2[i];
    cross(e2, s, cr);
    double W = dot(cr, r.d);
    double mn = std::min(U, std::min(V, W)), mx = std::max(U, std::max(V, W));
    if (!(mn >= 0 || mx <= 0)) return false;
    double UVW = U + V + W;
    if (UVW == 0) return false;
    // Ng = (v1-v0) x (v2-v0) = e0 x e1  (e0 = v2-v0, e1 = v0-v1)
    double Ng[3];
    cross(e0, e1, Ng);
    double den = 2 * dot(Ng, r.d);
    if (den == 0) return false;
    double T = 2 * dot(v0, Ng);
    double t = T / den;
    if (!(t >= r.tnear && t <= tfar)) return false;
    h.t = t;
    h.u = std::min(U / UVW, 1.0);
    h.v = std::min(V / UVW, 1.0);
    h.ng[0] = Ng[0];
    h.ng[1] = Ng[1];
    h.ng[2] = Ng[2];
    return true;
}

inline RayD make_ray(const RTCRay &ray) {
    RayD r;
    r.o[0] = ray.org_x; r.o[1] = ray.org_y; r.o[2] = ray.org_z;
    r.d[0] = ray.dir_x; r.d[1] = ray.dir_y; r.d[2] = ray.dir_z;
    for (int i = 0; i < 3; i++) r.inv[i] = 1.0 / r.d[i];
    r.tnear = ray.tnear;
    r.tfar = ray.tfar;
    return r;
}

template <bool ANY>
void traverse(SceneImpl *s, RTCRayHit *rayhit, RTCRay *ray_only) {
    RTCRay &ray = ANY ? *ray_only : rayhit->ray;
    if (s->nodes.empty() || s->prims.empty()) return;
    if (!(ray.tnear <= ray.tfar)) return;
    RayD r = make_ray(ray);
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    double best_t = r.tfar;
    unsigned best_geom = RTC_INVALID_GEOMETRY_ID, best_prim = RTC_INVALID_GEOMETRY_ID;
    RTCRayQueryContext ctx;
    rtcInitRayQueryContext(&ctx);
    while (sp > 0) {
        const Node &n = s->nodes[stack[--sp]];
        double tb;
        if (!hit_box(n.box, r, best_t, tb)) continue;
        if (n.count == 0) {
            // push far child first
            double ta, tc;
            bool hl = hit_box(s->nodes[n.left].box, r, best_t, ta);
            bool hr = hit_box(s->nodes[n.left + 1].box, r, best_t, tc);
            if (hl && hr) {
                if (ta <= tc) { stack[sp++] = n.left + 1; stack[sp++] = n.left; }
                else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
            } else if (hl) stack[sp++] = n.left;
            else if (hr) stack[sp++] = n.left + 1;
            continue;
        }
        for (int i = 0; i < n.count; i++) {
            const PrimRef &p = s->prims[n.left + i];
            GeomImpl *g = s->geoms[p.geom];
            if (g->type == RTC_GEOMETRY_TYPE_TRIANGLE) {
                const unsigned *idx = g->tri(p.prim);
                TriHit h;
                if (!hit_triangle(g->vertex(idx[0]), g->vertex(idx[1]), g->vertex(idx[2]), r, best_t, h)) continue;
                if (ANY) {
                    ray.tfar = -std::numeric_limits<float>::infinity();
                    return;
                }
                // tie policy: exactly equal t -> lowest (geomID, primID)
                if (h.t == best_t && best_geom != RTC_INVALID_GEOMETRY_ID &&
                    !(p.geom < best_geom || (p.geom == best_geom && p.prim < best_prim)))
                    continue;
                best_t = h.t;
                best_geom = p.geom;
                best_prim = p.prim;
                rayhit->hit.Ng_x = (float)h.ng[0];
                rayhit->hit.Ng_y = (float)h.ng[1];
                rayhit->hit.Ng_z = (float)h.ng[2];
                rayhit->hit.u = (float)h.u;
                rayhit->hit.v = (float)h.v;
                rayhit->hit.primID = p.prim;
                rayhit->hit.geomID = p.geom;
                rayhit->hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
                ray.tfar = (float)h.t;
            } else {
                int valid = -1;
                if (ANY) {
                    RTCOccludedFunctionNArguments a;
                    a.valid = &valid;
                    a.geometryUserPtr = g->user_ptr;
                    a.primID = p.prim;
                    a.context = &ctx;
                    a.ray = (RTCRayN *)&ray;
                    a.N = 1;
                    a.geomID = p.geom;
                    if (g->occluded_fn) g->occluded_fn(&a);
                    if (ray.tfar < 0) return;
                } else {
                    RTCIntersectFunctionNArguments a;
                    a.valid = &valid;
                    a.geometryUserPtr = g->user_ptr;
                    a.primID = p.prim;
                    a.context = &ctx;
                    a.rayhit = (RTCRayHitN *)rayhit;
                    a.N = 1;
                    a.geomID = p.geom;
                    float before = ray.tfar;
                    if (g->intersect_fn) g->intersect_fn(&a);
                    if (ray.tfar != before) {
                        best_t = ray.tfar;
                        best_geom = rayhit->hit.geomID;
                        best_prim = rayhit->hit.primID;
                    }
                }
            }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------- exported C API
RTC_NAMESPACE_BEGIN

RTC_API RTCDevice rtcNewDevice(const char *) { return (RTCDevice) new DeviceImpl(); }

RTC_API void rtcReleaseDevice(RTCDevice device) {
    DeviceImpl *d = (DeviceImpl *)device;
    if (getenv("LJ_SHIM_STATS")) {
        fprintf(stderr, "[embree_shim] closest=%llu any=%llu\n", (unsigned long long)g_closest.load(),
                (unsigned long long)g_any.load());
    }
    if (d && d->refs.fetch_sub(1) == 1) delete d;
}

RTC_API RTCScene rtcNewScene(RTCDevice) { return (RTCScene) new SceneImpl(); }

RTC_API void rtcReleaseScene(RTCScene scene) {
    SceneImpl *s = (SceneImpl *)scene;
    if (!s) return;
    if (s->refs.fetch_sub(1) == 1) {
        for (GeomImpl *g : s->geoms) release(g);
        delete s;
    }
}

RTC_API void rtcSetSceneBuildQuality(RTCScene, enum RTCBuildQuality) {}
RTC_API void rtcSetSceneFlags(RTCScene, enum RTCSceneFlags) {}

RTC_API void rtcCommitScene(RTCScene scene) {
    SceneImpl *s = (SceneImpl *)scene;
    s->prims.clear();
    s->bounds.reset();
    for (unsigned gi = 0; gi < s->geoms.size(); gi++) {
        GeomImpl *g = s->geoms[gi];
        for (unsigned pi = 0; pi < g->num_prims(); pi++) {
            PrimRef p;
            p.geom = gi;
            p.prim = pi;
            p.box.reset();
            if (g->type == RTC_GEOMETRY_TYPE_TRIANGLE) {
                const unsigned *idx = g->tri(pi);
                for (int k = 0; k < 3; k++) p.box.grow(g->vertex(idx[k]));
            } else {
                RTCBounds b;
                RTCBoundsFunctionArguments a;
                a.geometryUserPtr = g->user_ptr;
                a.primID = pi;
                a.timeStep = 0;
                a.bounds_o = &b;
                g->bounds_fn(&a);
                p.box.lo[0] = b.lower_x; p.box.lo[1] = b.lower_y; p.box.lo[2] = b.lower_z;
                p.box.hi[0] = b.upper_x; p.box.hi[1] = b.upper_y; p.box.hi[2] = b.upper_z;
            }
            for (int k = 0; k < 3; k++) p.c[k] = 0.5f * (p.box.lo[k] + p.box.hi[k]);
            s->bounds.grow(p.box);
            s->prims.push_back(p);
        }
    }
    Builder b{s->prims, s->nodes};
    if (!s->prims.empty()) b.build();
    s->committed = true;
}

RTC_API void rtcGetSceneBounds(RTCScene scene, struct RTCBounds *o) {
    SceneImpl *s = (SceneImpl *)scene;
    o->lower_x = s->bounds.lo[0]; o->lower_y = s->bounds.lo[1]; o->lower_z = s->bounds.lo[2];
    o->upper_x = s->bounds.hi[0]; o->upper_y = s->bounds.hi[1]; o->upper_z = s->bounds.hi[2];
    o->align0 = o->align1 = 0;
}

RTC_API RTCGeometry rtcNewGeometry(RTCDevice, enum RTCGeometryType type) {
    GeomImpl *g = new GeomImpl();
    g->type = type;
    return (RTCGeometry)g;
}

RTC_API void rtcReleaseGeometry(RTCGeometry geometry) { release((GeomImpl *)geometry); }
RTC_API void rtcCommitGeometry(RTCGeometry) {}
RTC_API void rtcSetGeometryVertexAttributeCount(RTCGeometry, unsigned int) {}

RTC_API unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry) {
    SceneImpl *s = (SceneImpl *)scene;
    GeomImpl *g = (GeomImpl *)geometry;
    g->refs.fetch_add(1);
    s->geoms.push_back(g);
    return (unsigned)s->geoms.size() - 1;
}

RTC_API void *rtcSetNewGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int, enum RTCFormat,
                                      size_t byteStride, size_t itemCount) {
    GeomImpl *g = (GeomImpl *)geometry;
    if (type == RTC_BUFFER_TYPE_VERTEX) {
        g->vbuf.assign(byteStride * itemCount + 16, 0);
        g->vstride = byteStride;
        g->vcount = itemCount;
        return g->vbuf.data();
    } else if (type == RTC_BUFFER_TYPE_INDEX) {
        g->ibuf.assign(byteStride * itemCount + 16, 0);
        g->istride = byteStride;
        g->icount = itemCount;
        return g->ibuf.data();
    }
    return nullptr;
}

RTC_API void rtcSetGeometryUserPrimitiveCount(RTCGeometry geometry, unsigned int n) { ((GeomImpl *)geometry)->user_prims = n; }
RTC_API void rtcSetGeometryUserData(RTCGeometry geometry, void *ptr) { ((GeomImpl *)geometry)->user_ptr = ptr; }
RTC_API void rtcSetGeometryBoundsFunction(RTCGeometry geometry, RTCBoundsFunction f, void *userPtr) {
    ((GeomImpl *)geometry)->bounds_fn = f;
    ((GeomImpl *)geometry)->bounds_user = userPtr;
}
RTC_API void rtcSetGeometryIntersectFunction(RTCGeometry geometry, RTCIntersectFunctionN f) { ((GeomImpl *)geometry)->intersect_fn = f; }
RTC_API void rtcSetGeometryOccludedFunction(RTCGeometry geometry, RTCOccludedFunctionN f) { ((GeomImpl *)geometry)->occluded_fn = f; }

RTC_API void rtcIntersect1(RTCScene scene, struct RTCRayHit *rayhit, struct RTCIntersectArguments *) {
    static thread_local unsigned long long local = 0, ticks = 0;
    const unsigned long long t0 = shim_rdtsc();
    traverse<false>((SceneImpl *)scene, rayhit, nullptr);
    ticks += shim_rdtsc() - t0;
    if ((++local & 1023) == 0) { g_closest.fetch_add(1024, std::memory_order_relaxed); g_ticks.fetch_add(ticks, std::memory_order_relaxed); ticks = 0; }
}

RTC_API void rtcOccluded1(RTCScene scene, struct RTCRay *ray, struct RTCOccludedArguments *) {
    static thread_local unsigned long long local = 0, ticks = 0;
    const unsigned long long t0 = shim_rdtsc();
    traverse<true>((SceneImpl *)scene, nullptr, ray);
    ticks += shim_rdtsc() - t0;
    if ((++local & 1023) == 0) { g_any.fetch_add(1024, std::memory_order_relaxed); g_ticks.fetch_add(ticks, std::memory_order_relaxed); ticks = 0; }
}

RTC_NAMESPACE_END

// Ray counters (approximate to 1024 per thread) for the CPU-baseline Mrays/s figure.
extern "C" void ljshim_get_counters(unsigned long long *closest, unsigned long long *any) {
    *closest = g_closest.load();
    *any = g_any.load();
}
extern "C" void ljshim_reset_counters(void) {
    g_closest.store(0);
    g_any.store(0);
    g_ticks.store(0);
}
// ticks inside the two ray-cast entry points (summed over threads) and the current time-stamp counter
extern "C" void ljshim_get_ticks(unsigned long long *inside, unsigned long long *now) {
    *inside = g_ticks.load();
    *now = shim_rdtsc();
}
