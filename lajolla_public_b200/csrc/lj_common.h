// lj_common.h -- fp32 vector math shared by every device function.
//
// All per-ray / per-vertex arithmetic lives in LJ_HD inline functions so the same source is
// compiled by nvcc for sm_100a (the product) and by g++ for tests/hostsim (unit tests of the device
// math against the oracle on the GPU-less authoring box; never shipped, never benchmarked).
// Restates the parts of the reference's vector.h / frame.h / lajolla.h the hot path uses, in fp32.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LJ_HD __host__ __device__ __forceinline__
#define LJ_D __device__ __forceinline__
// One copy per translation unit instead of one per call site: the material dispatchers are called five times per
// shade and inlining them made k_shade 1.5 MB of SASS, which the instruction caches cannot hold when the warps of an
// SM sit in different materials (disney_bsdf: 8 % of issue slots used, the rest waiting for instructions).
#define LJ_HD_CALL __attribute__((unused)) static __host__ __device__ __noinline__
#else
#define LJ_HD inline
#define LJ_D inline
#define LJ_HD_CALL static inline
#endif

namespace lj {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kInvTwoPi = 0.15915494309189533577f;
constexpr float kInvFourPi = 0.07957747154594766788f;
#define LJ_INF (__builtin_huge_valf())

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct alignas(16) V4 { float x, y, z, w; };

LJ_HD V2 mk2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
LJ_HD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
LJ_HD V3 mk3(float s) { return mk3(s, s, s); }
LJ_HD V4 mk4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
LJ_HD V4 mk4(V3 a, float w) { return mk4(a.x, a.y, a.z, w); }
LJ_HD V3 xyz(V4 a) { return mk3(a.x, a.y, a.z); }

LJ_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
LJ_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
LJ_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
LJ_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
LJ_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
LJ_HD V3 operator*(float s, V3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
LJ_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
LJ_HD V3 operator/(V3 a, V3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
LJ_HD V3 &operator+=(V3 &a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
LJ_HD V3 &operator*=(V3 &a, V3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
LJ_HD V3 &operator*=(V3 &a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
LJ_HD V2 operator+(V2 a, V2 b) { return mk2(a.x + b.x, a.y + b.y); }
LJ_HD V2 operator-(V2 a, V2 b) { return mk2(a.x - b.x, a.y - b.y); }
LJ_HD V2 operator*(V2 a, float s) { return mk2(a.x * s, a.y * s); }
LJ_HD V2 operator*(float s, V2 a) { return mk2(a.x * s, a.y * s); }

LJ_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LJ_HD V3 cross(V3 a, V3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
LJ_HD float length_squared(V3 a) { return dot(a, a); }
LJ_HD float length(V3 a) { return sqrtf(dot(a, a)); }
LJ_HD float distance(V3 a, V3 b) { return length(a - b); }
LJ_HD float distance_squared(V3 a, V3 b) { return length_squared(a - b); }
// The same sums with the rounding of every step spelled out.  `x*x + y*y + z*z` may be contracted to FMAs in more
// than one way, and the compiler's pick depends on the code around the expression: quantities that several KERNELS
// must reproduce bit for bit (the length of a walk segment, wavefront.cu) are formed with these.
LJ_HD float dot_fixed(V3 a, V3 b) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x)));
#else
    return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x));
#endif
}
LJ_HD float distance_fixed(V3 a, V3 b) { V3 v = a - b; return sqrtf(dot_fixed(v, v)); }
LJ_HD float sum_squares_fixed(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(b, b, __fmul_rn(a, a));
#else
    return fmaf(b, b, a * a);
#endif
}
// vector.h:249-257: zero vector stays zero.
LJ_HD V3 normalize(V3 a) {
    float l = length(a);
    if (l <= 0) return mk3(0, 0, 0);
    return a / l;
}
LJ_HD float max3(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
LJ_HD float min3(V3 a) { return fminf(a.x, fminf(a.y, a.z)); }
LJ_HD float avg3(V3 a) { return (a.x + a.y + a.z) / 3; }
LJ_HD float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
LJ_HD V3 exp3(V3 a) { return mk3(expf(a.x), expf(a.y), expf(a.z)); }
LJ_HD V3 sqrt3(V3 a) { return mk3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
LJ_HD V3 max3v(V3 a, V3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
LJ_HD V3 min3v(V3 a, V3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
LJ_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
LJ_HD int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
LJ_HD bool is_finite(float x) { return fabsf(x) <= 3.402823466e+38f; }

// spectrum.h:32-34
LJ_HD float luminance(V3 s) { return s.x * 0.212671f + s.y * 0.715160f + s.z * 0.072169f; }

// lajolla.h:48-56
LJ_HD int modulo_i(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }
LJ_HD float modulo_f(float a, float b) { float r = fmodf(a, b); return r < 0.0f ? r + b : r; }

// frame.h:6-17 (Duff et al. style branch at n.z ~ -1)
struct Frame { V3 x, y, n; };
LJ_HD void coordinate_system(V3 n, V3 &a, V3 &b) {
    if (n.z < -1.f + 1e-6f) {
        a = mk3(0, -1, 0);
        b = mk3(-1, 0, 0);
    } else {
        // Same frame as frame.h:6-17.  In fp32 1 + n.z cancels near the pole (the reference is double), so
        // below the equator 1 / (1 + n.z) is taken from the unit-length identity (1 - n.z) / (n.x^2 + n.y^2).
        float s = n.z < -0.5f ? (1 - n.z) / (n.x * n.x + n.y * n.y) : 1 / (1 + n.z);
        float t = -n.x * n.y * s;
        a = mk3(1 - n.x * n.x * s, t, -n.x);
        b = mk3(t, 1 - n.y * n.y * s, -n.y);
    }
}
LJ_HD Frame make_frame(V3 n) { Frame f; f.n = n; coordinate_system(n, f.x, f.y); return f; }
LJ_HD Frame make_frame(V3 x, V3 y, V3 n) { Frame f; f.x = x; f.y = y; f.n = n; return f; }
LJ_HD Frame flip(Frame f) { return make_frame(-f.x, -f.y, -f.n); }
LJ_HD V3 to_local(const Frame &f, V3 v) { return mk3(dot(v, f.x), dot(v, f.y), dot(v, f.n)); }
LJ_HD V3 to_world(const Frame &f, V3 v) { return f.x * v.x + f.y * v.y + f.n * v.z; }

// Row-major 4x4 (matrix.h); only the two transforms the hot path applies (transform.cpp).
struct M44 { float m[16]; };
LJ_HD V3 xform_point(const M44 &t, V3 p) {
    float x = t.m[0] * p.x + t.m[1] * p.y + t.m[2] * p.z + t.m[3];
    float y = t.m[4] * p.x + t.m[5] * p.y + t.m[6] * p.z + t.m[7];
    float z = t.m[8] * p.x + t.m[9] * p.y + t.m[10] * p.z + t.m[11];
    float w = t.m[12] * p.x + t.m[13] * p.y + t.m[14] * p.z + t.m[15];
    float inv = 1 / w;
    return mk3(x * inv, y * inv, z * inv);
}
LJ_HD V3 xform_vector(const M44 &t, V3 v) {
    return mk3(t.m[0] * v.x + t.m[1] * v.y + t.m[2] * v.z,
               t.m[4] * v.x + t.m[5] * v.y + t.m[6] * v.z,
               t.m[8] * v.x + t.m[9] * v.y + t.m[10] * v.z);
}

LJ_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
LJ_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// 16-byte vector load through the read-only path on the device.
LJ_HD V4 ld4(const V4 *p) {
#if defined(__CUDA_ARCH__)
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    return mk4(v.x, v.y, v.z, v.w);
#else
    return *p;
#endif
}

}  // namespace lj
