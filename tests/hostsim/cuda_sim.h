// tests/hostsim/cuda_sim.h -- TEST INFRASTRUCTURE ONLY.
// Minimal serial stand-in for the CUDA runtime calls and device intrinsics that
// lajolla_public_b200/csrc uses, so the *.cu sources compile with g++ (-x c++ -DLJ_HOSTSIM) into
// tests/hostsim/_build/libljsim.so and the host orchestration + device functions can be exercised on
// a machine without a GPU.  Kernels run one "thread" at a time with warps of width 1.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define LJ_GRID_CONSTANT
#define LJ_LANE() 0
#define LJ_WARP_WIDTH 1

struct SimDim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local SimDim3 threadIdx, blockIdx, blockDim, gridDim;
struct float4 { float x, y, z, w; };

#define LJ_LAUNCH(kernel, grid, block, stream, ...)                              \
    do {                                                                         \
        gridDim.x = (unsigned)(grid);                                            \
        blockDim.x = (unsigned)(block);                                          \
        for (unsigned b__ = 0; b__ < gridDim.x; b__++)                           \
            for (unsigned t__ = 0; t__ < blockDim.x; t__++) {                    \
                blockIdx.x = b__;                                                \
                threadIdx.x = t__;                                               \
                kernel(__VA_ARGS__);                                             \
            }                                                                    \
    } while (0)

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef std::chrono::steady_clock::time_point *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };

inline const char *cudaGetErrorString(cudaError_t) { return "hostsim error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t lj_dev_alloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t lj_dev_free(void *p) { free(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new std::chrono::steady_clock::time_point(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { *e = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(*b - *a).count();
    return cudaSuccess;
}

// device intrinsics, warp width 1
template <typename T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
inline int atomicMin(int *p, int v) { int o = *p; *p = std::min(o, v); return o; }
inline int atomicMax(int *p, int v) { int o = *p; *p = std::max(o, v); return o; }
inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
template <typename T> inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline void __threadfence() {}
inline void __syncwarp() {}
using std::max;
using std::min;
inline float lj_warp_min(float x) { return x; }
inline float lj_warp_max(float x) { return x; }
inline double lj_warp_sum(double x) { return x; }
inline int lj_float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline unsigned __activemask() { return 1u; }
// position of the offset-th set bit of mask at or above base (offset >= 1), 0xffffffff if there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    for (unsigned b = base; b < 32; b++) if ((mask >> b) & 1u) { if (--offset == 0) return b; }
    return 0xffffffffu;
}
inline bool __any_sync(unsigned, bool p) { return p; }
