// Micro-benchmark: cost of divergent record fetches through the L1 data pipe on sm_100a.
// Each lane fetches a pseudo-random record (BVH-node-like access: half of the fetches go to the first 256 records)
// and folds it into a checksum; variants differ in how the record is loaded (5 x LDG.128 at stride 80, 3 x LDG.256 at
// stride 96, 2 x LDG.256 + LDG.128, ...).  Prints ns per warp-level record fetch and fetches per second.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

struct F4 { float x, y, z, w; };
struct F8 { float v[8]; };
__device__ __forceinline__ F4 ld128(const void *p) {
    F4 r; asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ F8 ld256(const void *p) {
    F8 r; asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p)); return r;
}
__device__ __forceinline__ float sum4(F4 a) { return a.x + a.y + a.z + a.w; }
__device__ __forceinline__ float sum8(F8 a) { float s = 0; for (int i = 0; i < 8; i++) s += a.v[i]; return s; }

// VARIANT: 0 = 5x128 @80, 1 = 3x256 @96, 2 = 2x256+128 @96, 3 = 3x128 @48 (prim), 4 = 256+128 @64 (prim), 5 = 2x256 @64, 6 = 4x256 @128, 7 = 5x128 @96
template <int VARIANT>
__global__ void __launch_bounds__(128, 8) k(const char *tab, uint32_t nrec, int iters, float *out, int dep) {
    uint32_t s = hash32(blockIdx.x * 128 + threadIdx.x + 1);
    float acc = 0;
    constexpr int stride = VARIANT == 0 ? 80 : VARIANT == 1 ? 96 : VARIANT == 2 ? 96 : VARIANT == 3 ? 48 : VARIANT == 4 ? 64 : VARIANT == 5 ? 64 : VARIANT == 6 ? 128 : 96;
    for (int it = 0; it < iters; it++) {
        s = hash32(s + it);
        uint32_t r = (s & 1u) ? (s >> 1) % 256u : (s >> 1) % nrec;
        if (dep) r = (r + (__float_as_uint(acc) & 1u)) % nrec;  // dependent chain like a traversal
        const char *p = tab + (size_t)r * stride;
        if (VARIANT == 0 || VARIANT == 7) acc += sum4(ld128(p)) + sum4(ld128(p + 16)) + sum4(ld128(p + 32)) + sum4(ld128(p + 48)) + sum4(ld128(p + 64));
        if (VARIANT == 1) acc += sum8(ld256(p)) + sum8(ld256(p + 32)) + sum8(ld256(p + 64));
        if (VARIANT == 2) acc += sum8(ld256(p)) + sum8(ld256(p + 32)) + sum4(ld128(p + 64));
        if (VARIANT == 3) acc += sum4(ld128(p)) + sum4(ld128(p + 16)) + sum4(ld128(p + 32));
        if (VARIANT == 4) acc += sum8(ld256(p)) + sum4(ld128(p + 32));
        if (VARIANT == 5) acc += sum8(ld256(p)) + sum8(ld256(p + 32));
        if (VARIANT == 6) acc += sum8(ld256(p)) + sum8(ld256(p + 32)) + sum8(ld256(p + 64)) + sum8(ld256(p + 96));
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <int V>
void run(const char *name, const char *tab, uint32_t nrec, float *out, int dep) {
    const int blocks = 148 * 8, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<blocks, 128>>>(tab, nrec, 200, out, dep);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<V><<<blocks, 128>>>(tab, nrec, iters, out, dep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fetches = (double)blocks * 128 * iters;
    printf("%-28s nrec=%7u dep=%d  %8.3f ms  %7.2f Gfetch/s  %6.2f SM-cycles per warp fetch (at 1.965 GHz)\n", name, nrec, dep, ms, fetches / ms / 1e6,
           ms * 1e-3 * 1.965e9 * 148 / (fetches / 32));
    cudaError_t err = cudaGetLastError(); if (err) printf("CUDA error %s\n", cudaGetErrorString(err));
}

int main() {
    const size_t bytes = 64u << 20;
    char *tab; cudaMalloc(&tab, bytes);
    float *h = (float *)malloc(bytes);
    for (size_t i = 0; i < bytes / 4; i++) h[i] = (float)(i & 1023) * 1e-6f;
    cudaMemcpy(tab, h, bytes, cudaMemcpyHostToDevice);
    float *out; cudaMalloc(&out, 148 * 8 * 128 * 4);
    for (int dep = 0; dep < 2; dep++)
        for (uint32_t nrec : {9333u, 66446u}) {
            run<0>("node 5x128 @80", tab, nrec, out, dep);
            run<7>("node 5x128 @96", tab, nrec, out, dep);
            run<1>("node 3x256 @96", tab, nrec, out, dep);
            run<2>("node 2x256+128 @96", tab, nrec, out, dep);
            run<5>("node 2x256 @64", tab, nrec, out, dep);
            run<6>("node 4x256 @128", tab, nrec, out, dep);
            run<3>("prim 3x128 @48", tab, nrec, out, dep);
            run<4>("prim 256+128 @64", tab, nrec, out, dep);
        }
    return 0;
}
