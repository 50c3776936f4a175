// Shared helpers for libodil_b200 (sm_100a). Error reporting, launch accounting, reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "odil_b200.h"

namespace odil {

std::string& last_error_ref();
std::atomic<int64_t>& launch_counter();

inline int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return -1;
}

#define ODIL_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return ::odil::fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define ODIL_LAUNCHED()                                                                                  \
    do {                                                                                                 \
        ::odil::launch_counter()++;                                                                      \
        cudaError_t e_ = cudaGetLastError(); /* clears a refused launch so that it does not stick */   \
        if (e_ != cudaSuccess)                                                                           \
            return ::odil::fail("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

#define ODIL_REQUIRE(cond, ...)                       \
    do {                                              \
        if (!(cond)) return ::odil::fail(__VA_ARGS__); \
    } while (0)

// Per-device scratch for reductions that have no plan (sum_squares, dot). Allocated on first use
// (call once before CUDA-graph capture).
double* reduction_scratch(int nslots);

constexpr int kMaxPartialBlocks = 1024;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0. `red` must hold 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < nw ? red[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

// Second stage: deterministic sum of `n` block partials into out[0] (fixed order).
// `static`: one copy per translation unit, so no relocatable device code is needed.
static __global__ void k_reduce_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    v = block_sum(v, red);
    if (threadIdx.x == 0) out[0] = v;
}

template <typename T>
struct alignas(16) Vec4 {
    T x, y, z, w;
};

// One Adam update (reference optimizer.py:311-319), shared by k_adam (optim.cu) and the transposed interpolation that
// applies the update of the finest multigrid term while it streams the gradient (mg_march.cuh).
template <typename T>
__device__ __forceinline__ void adam_one(T& x, T& m, T& v, const T g, const T alpha, const T omb1, const T omb2,
                                         const T eps) {
    // Same operation order as the reference; no FMA contraction across the rounding points that
    // matter (m and v updates are written as the reference writes them).
    m = m + (g - m) * omb1;
    v = v + (g * g - v) * omb2;
    x = x - (m * alpha) / (sqrt(v) + eps);
}

}  // namespace odil
