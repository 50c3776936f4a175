// Optimizer updates and vector building blocks (sm_100a, HBM-bound, 128-bit accesses).
// adam_step restates reference optimizer.py:311-319 (AdamNativeOptimizer._step), gd_step :269-270.
#include <algorithm>

#include "common.cuh"

namespace odil {

constexpr int kMaxTensors = 16;

template <typename T>
struct AdamBatch {
    T* x[kMaxTensors];
    T* m[kMaxTensors];
    T* v[kMaxTensors];
    const T* g[kMaxTensors];
    int64_t n[kMaxTensors];
    int vec[kMaxTensors];  // 1 if all four pointers are 16-byte aligned
    T alpha, omb1, omb2, eps;
    const double* alpha_dev;  // when set, the step size is read from device memory (graph-replayable epochs)
};

template <typename T>
__global__ void __launch_bounds__(256) k_adam(AdamBatch<T> b) {
    const int t = blockIdx.y;
    const int64_t n = b.n[t];
    T* __restrict__ x = b.x[t];
    T* __restrict__ m = b.m[t];
    T* __restrict__ v = b.v[t];
    const T* __restrict__ g = b.g[t];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const T alpha = b.alpha_dev ? (T)__ldg(b.alpha_dev) : b.alpha;
    if (b.vec[t]) {
        const int64_t n4 = n / 4;
        for (int64_t i = tid; i < n4; i += stride) {
            Vec4<T> xx = reinterpret_cast<Vec4<T>*>(x)[i];
            Vec4<T> mm = reinterpret_cast<Vec4<T>*>(m)[i];
            Vec4<T> vv = reinterpret_cast<Vec4<T>*>(v)[i];
            const Vec4<T> gg = reinterpret_cast<const Vec4<T>*>(g)[i];
            adam_one(xx.x, mm.x, vv.x, gg.x, alpha, b.omb1, b.omb2, b.eps);
            adam_one(xx.y, mm.y, vv.y, gg.y, alpha, b.omb1, b.omb2, b.eps);
            adam_one(xx.z, mm.z, vv.z, gg.z, alpha, b.omb1, b.omb2, b.eps);
            adam_one(xx.w, mm.w, vv.w, gg.w, alpha, b.omb1, b.omb2, b.eps);
            reinterpret_cast<Vec4<T>*>(x)[i] = xx;
            reinterpret_cast<Vec4<T>*>(m)[i] = mm;
            reinterpret_cast<Vec4<T>*>(v)[i] = vv;
        }
        for (int64_t i = n4 * 4 + tid; i < n; i += stride) adam_one(x[i], m[i], v[i], g[i], alpha, b.omb1, b.omb2, b.eps);
    } else {
        for (int64_t i = tid; i < n; i += stride) adam_one(x[i], m[i], v[i], g[i], alpha, b.omb1, b.omb2, b.eps);
    }
}

template <typename T>
struct GdBatch {
    T* x[kMaxTensors];
    const T* g[kMaxTensors];
    int64_t n[kMaxTensors];
    T lr;
};

template <typename T>
__global__ void __launch_bounds__(256) k_gd(GdBatch<T> b) {
    const int t = blockIdx.y;
    const int64_t n = b.n[t];
    T* __restrict__ x = b.x[t];
    const T* __restrict__ g = b.g[t];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = x[i] - g[i] * b.lr;
}

template <typename T>
__global__ void __launch_bounds__(256) k_axpby(int64_t n, T a, const T* __restrict__ x, T bb, T* __restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T yi = bb == T(0) ? T(0) : bb * y[i];
        y[i] = a * x[i] + yi;
    }
}

// Conjugate-gradient vector updates; the step scalars are read from device memory (num[0] / den[0]), so an
// iteration never waits for the host.
template <typename T>
__global__ void __launch_bounds__(256) k_cg_update_xr(int64_t n, const double* __restrict__ num,
                                                      const double* __restrict__ den, const T* __restrict__ p,
                                                      const T* __restrict__ q, T* __restrict__ x, T* __restrict__ r) {
    const double d = den[0];
    const T alpha = d != 0.0 ? (T)(num[0] / d) : T(0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        x[i] += alpha * p[i];
        r[i] -= alpha * q[i];
    }
}
template <typename T>
__global__ void __launch_bounds__(256) k_cg_update_p(int64_t n, const double* __restrict__ num,
                                                     const double* __restrict__ den, const T* __restrict__ r,
                                                     T* __restrict__ p) {
    const double d = den[0];
    const T beta = d != 0.0 ? (T)(num[0] / d) : T(0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = r[i] + beta * p[i];
}

template <typename T, bool DOT>
__global__ void __launch_bounds__(256) k_dot_partial(const T* __restrict__ x, const T* __restrict__ y, int64_t n,
                                                     double* __restrict__ partials) {
    __shared__ double red[32];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a = (double)x[i];
        acc += DOT ? a * (double)y[i] : a * a;
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}


// ------------------------------------------------------------------------------------------------
// L-BFGS building blocks (compact / Byrd-Nocedal-Schnabel form): one pass over the 2m history vectors
// for all inner products [S Y]^T g, one pass for the direction d = a0*g + sum_r coef_r V_r.
// Replaces the host-side SciPy L-BFGS-B vector algebra of the reference (optimizer.py:95-105).
// V is a row-major [k][n] matrix (row stride `ld` elements).
// ------------------------------------------------------------------------------------------------
constexpr int kMdChunk = 2048;     // elements of g staged per block iteration
constexpr int kMdMaxRows = 256;

template <typename T>
__global__ void __launch_bounds__(256) k_multi_dot(const T* __restrict__ V, int64_t ld, int k, const T* __restrict__ g,
                                                   int64_t n, double* __restrict__ partials /*[k][gridDim.x]*/) {
    __shared__ T gs[kMdChunk];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = 8;
    double acc[kMdMaxRows / NW];
#pragma unroll
    for (int j = 0; j < kMdMaxRows / NW; ++j) acc[j] = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * kMdChunk; base < n; base += (int64_t)gridDim.x * kMdChunk) {
        const int len = (int)min((int64_t)kMdChunk, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += blockDim.x) gs[i] = g[base + i];
        __syncthreads();
#pragma unroll 1
        for (int j = 0; j * NW + warp < k; ++j) {
            const T* row = V + (int64_t)(j * NW + warp) * ld + base;
            double a = 0.0;
            for (int i = lane; i < len; i += 32) a += (double)row[i] * (double)gs[i];
            acc[j] += a;
        }
    }
    for (int j = 0; j * NW + warp < k; ++j) {
        const double a = warp_sum(acc[j]);
        if (lane == 0) partials[(int64_t)(j * NW + warp) * gridDim.x + blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(256) k_multi_dot_final(const double* __restrict__ partials, int nblocks,
                                                         double* __restrict__ out) {
    __shared__ double red[32];
    const double* row = partials + (int64_t)blockIdx.x * nblocks;
    double v = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += row[i];
    v = block_sum(v, red);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
}

template <typename T>
__global__ void __launch_bounds__(256) k_multi_axpy(const T* __restrict__ V, int64_t ld, int k,
                                                    const double* __restrict__ coef /*[k] device*/, double a0,
                                                    const T* __restrict__ g, T* __restrict__ d, int64_t n) {
    __shared__ double cs[kMdMaxRows];
    for (int i = threadIdx.x; i < k; i += blockDim.x) cs[i] = coef[i];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double acc = g ? a0 * (double)g[i] : 0.0;
        for (int r = 0; r < k; ++r) acc += cs[r] * (double)V[(int64_t)r * ld + i];
        d[i] = (T)acc;
    }
}

static unsigned blocks_for(int64_t n, int per_thread) {
    int64_t nb = (n + 256ll * per_thread - 1) / (256ll * per_thread);
    if (nb < 1) nb = 1;
    const int64_t cap = 148ll * 16;
    return (unsigned)(nb > cap ? cap : nb);
}

template <typename T>
static int run_adam(int nt, void* const* x, void* const* m, void* const* v, const void* const* g,
                    const int64_t* counts, double alpha, const double* alpha_dev, double omb1, double omb2, double eps,
                    cudaStream_t st) {
    for (int base = 0; base < nt; base += kMaxTensors) {
        AdamBatch<T> b;
        b.alpha_dev = alpha_dev;
        const int k = nt - base < kMaxTensors ? nt - base : kMaxTensors;
        int64_t nmax = 0;
        for (int i = 0; i < k; ++i) {
            b.x[i] = (T*)x[base + i];
            b.m[i] = (T*)m[base + i];
            b.v[i] = (T*)v[base + i];
            b.g[i] = (const T*)g[base + i];
            b.n[i] = counts[base + i];
            ODIL_REQUIRE(b.n[i] >= 0 && (b.n[i] == 0 || (b.x[i] && b.m[i] && b.v[i] && b.g[i])), "adam: null tensor %d",
                         base + i);
            b.vec[i] = (((uintptr_t)b.x[i] | (uintptr_t)b.m[i] | (uintptr_t)b.v[i] | (uintptr_t)b.g[i]) % 16) == 0;
            nmax = b.n[i] > nmax ? b.n[i] : nmax;
        }
        if (nmax == 0) continue;
        b.alpha = (T)alpha;
        b.omb1 = (T)omb1;
        b.omb2 = (T)omb2;
        b.eps = (T)eps;
        dim3 grid(blocks_for(nmax, 8), k);
        k_adam<T><<<grid, 256, 0, st>>>(b);
        ODIL_LAUNCHED();
    }
    return 0;
}

template <typename T>
static int run_gd(int nt, void* const* x, const void* const* g, const int64_t* counts, double lr, cudaStream_t st) {
    for (int base = 0; base < nt; base += kMaxTensors) {
        GdBatch<T> b;
        const int k = nt - base < kMaxTensors ? nt - base : kMaxTensors;
        int64_t nmax = 0;
        for (int i = 0; i < k; ++i) {
            b.x[i] = (T*)x[base + i];
            b.g[i] = (const T*)g[base + i];
            b.n[i] = counts[base + i];
            nmax = b.n[i] > nmax ? b.n[i] : nmax;
        }
        if (nmax == 0) continue;
        b.lr = (T)lr;
        dim3 grid(blocks_for(nmax, 8), k);
        k_gd<T><<<grid, 256, 0, st>>>(b);
        ODIL_LAUNCHED();
    }
    return 0;
}

template <typename T, bool DOT>
static int run_dot(const void* x, const void* y, int64_t n, double* out, cudaStream_t st) {
    double* scratch = reduction_scratch(kMaxPartialBlocks);
    ODIL_REQUIRE(scratch != nullptr, "reduction scratch allocation failed");
    int nb = (int)blocks_for(n, 16);
    if (nb > kMaxPartialBlocks) nb = kMaxPartialBlocks;
    k_dot_partial<T, DOT><<<nb, 256, 0, st>>>((const T*)x, (const T*)y, n, scratch);
    ODIL_LAUNCHED();
    k_reduce_partials<<<1, 1024, 0, st>>>(scratch, nb, out);
    ODIL_LAUNCHED();
    return 0;
}

}  // namespace odil

using namespace odil;

extern "C" {

int odil_b200_adam_step(int ntensors, void* const* x, void* const* m, void* const* v, const void* const* g,
                        const int64_t* counts, int dtype, double alpha, double one_minus_beta1,
                        double one_minus_beta2, double epsilon, void* stream) {
    ODIL_REQUIRE(ntensors >= 0 && (ntensors == 0 || (x && m && v && g && counts)), "adam: bad arguments");
    if (dtype == ODIL_B200_F32)
        return run_adam<float>(ntensors, x, m, v, g, counts, alpha, nullptr, one_minus_beta1, one_minus_beta2, epsilon,
                               (cudaStream_t)stream);
    if (dtype == ODIL_B200_F64)
        return run_adam<double>(ntensors, x, m, v, g, counts, alpha, nullptr, one_minus_beta1, one_minus_beta2, epsilon,
                                (cudaStream_t)stream);
    return fail("dtype=%d unsupported", dtype);
}

// Head of a replayed Adam epoch: out[0] = table[step[0]]; step[0] += 1 -- the per-epoch step size (optimizer.py:307-309)
// picked from a device table in ONE launch, so that a captured epoch reads nothing from the host.
static __global__ void k_table_pick(const double* __restrict__ table, long long* __restrict__ step, double* __restrict__ out) {
    const long long s = step[0];
    out[0] = table[s];
    step[0] = s + 1;
}

int odil_b200_table_pick(const double* table, int64_t* step, double* out, void* stream) {
    ODIL_REQUIRE(table && step && out, "table_pick: null pointer");
    k_table_pick<<<1, 1, 0, (cudaStream_t)stream>>>(table, reinterpret_cast<long long*>(step), out);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_adam_step_dev(int ntensors, void* const* x, void* const* m, void* const* v, const void* const* g,
                            const int64_t* counts, int dtype, const double* alpha_dev, double one_minus_beta1,
                            double one_minus_beta2, double epsilon, void* stream) {
    ODIL_REQUIRE(alpha_dev && ntensors >= 0 && (ntensors == 0 || (x && m && v && g && counts)),
                 "adam_dev: bad arguments");
    if (dtype == ODIL_B200_F32)
        return run_adam<float>(ntensors, x, m, v, g, counts, 0.0, alpha_dev, one_minus_beta1, one_minus_beta2, epsilon,
                               (cudaStream_t)stream);
    if (dtype == ODIL_B200_F64)
        return run_adam<double>(ntensors, x, m, v, g, counts, 0.0, alpha_dev, one_minus_beta1, one_minus_beta2,
                                epsilon, (cudaStream_t)stream);
    return fail("dtype=%d unsupported", dtype);
}

int odil_b200_gd_step(int ntensors, void* const* x, const void* const* g, const int64_t* counts, int dtype,
                      double lr, void* stream) {
    ODIL_REQUIRE(ntensors >= 0 && (ntensors == 0 || (x && g && counts)), "gd: bad arguments");
    if (dtype == ODIL_B200_F32) return run_gd<float>(ntensors, x, g, counts, lr, (cudaStream_t)stream);
    if (dtype == ODIL_B200_F64) return run_gd<double>(ntensors, x, g, counts, lr, (cudaStream_t)stream);
    return fail("dtype=%d unsupported", dtype);
}

int odil_b200_axpby(int64_t count, int dtype, double a, const void* x, double b, void* y, void* stream) {
    ODIL_REQUIRE(count >= 0 && (count == 0 || (x && y)), "axpby: bad arguments");
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ODIL_B200_F32)
        k_axpby<float><<<blocks_for(count, 4), 256, 0, st>>>(count, (float)a, (const float*)x, (float)b, (float*)y);
    else if (dtype == ODIL_B200_F64)
        k_axpby<double><<<blocks_for(count, 4), 256, 0, st>>>(count, a, (const double*)x, b, (double*)y);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_cg_update_xr(int64_t count, int dtype, const double* num, const double* den, const void* p, const void* q,
                           void* x, void* r, void* stream) {
    ODIL_REQUIRE(num && den && p && q && x && r && count >= 0, "cg_update_xr: bad arguments");
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ODIL_B200_F32)
        k_cg_update_xr<float><<<blocks_for(count, 4), 256, 0, st>>>(count, num, den, (const float*)p, (const float*)q,
                                                                    (float*)x, (float*)r);
    else if (dtype == ODIL_B200_F64)
        k_cg_update_xr<double><<<blocks_for(count, 4), 256, 0, st>>>(count, num, den, (const double*)p, (const double*)q,
                                                                     (double*)x, (double*)r);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_cg_update_p(int64_t count, int dtype, const double* num, const double* den, const void* r, void* p,
                          void* stream) {
    ODIL_REQUIRE(num && den && r && p && count >= 0, "cg_update_p: bad arguments");
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ODIL_B200_F32)
        k_cg_update_p<float><<<blocks_for(count, 4), 256, 0, st>>>(count, num, den, (const float*)r, (float*)p);
    else if (dtype == ODIL_B200_F64)
        k_cg_update_p<double><<<blocks_for(count, 4), 256, 0, st>>>(count, num, den, (const double*)r, (double*)p);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_sum_squares(const void* x, int64_t count, int dtype, double* sumsq_out, void* stream) {
    ODIL_REQUIRE(x && sumsq_out && count >= 0, "sum_squares: bad arguments");
    if (dtype == ODIL_B200_F32) return run_dot<float, false>(x, x, count, sumsq_out, (cudaStream_t)stream);
    if (dtype == ODIL_B200_F64) return run_dot<double, false>(x, x, count, sumsq_out, (cudaStream_t)stream);
    return fail("dtype=%d unsupported", dtype);
}

int odil_b200_dot(const void* x, const void* y, int64_t count, int dtype, double* out, void* stream) {
    ODIL_REQUIRE(x && y && out && count >= 0, "dot: bad arguments");
    if (dtype == ODIL_B200_F32) return run_dot<float, true>(x, y, count, out, (cudaStream_t)stream);
    if (dtype == ODIL_B200_F64) return run_dot<double, true>(x, y, count, out, (cudaStream_t)stream);
    return fail("dtype=%d unsupported", dtype);
}


int odil_b200_multi_dot(const void* V, int64_t ld, int k, const void* g, int64_t count, int dtype, double* out,
                        void* stream) {
    ODIL_REQUIRE(V && g && out && k >= 1 && k <= kMdMaxRows && count >= 0 && ld >= count, "multi_dot: bad arguments");
    const int nb = (int)std::min<int64_t>(512, (count + kMdChunk - 1) / kMdChunk > 0 ? (count + kMdChunk - 1) / kMdChunk : 1);
    double* scratch = reduction_scratch(kMdMaxRows * 512);
    ODIL_REQUIRE(scratch != nullptr, "reduction scratch allocation failed");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ODIL_B200_F32)
        k_multi_dot<float><<<nb, 256, 0, st>>>((const float*)V, ld, k, (const float*)g, count, scratch);
    else if (dtype == ODIL_B200_F64)
        k_multi_dot<double><<<nb, 256, 0, st>>>((const double*)V, ld, k, (const double*)g, count, scratch);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    k_multi_dot_final<<<k, 256, 0, st>>>(scratch, nb, out);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_multi_axpy(const void* V, int64_t ld, int k, const double* coef, double a0, const void* g, void* d,
                         int64_t count, int dtype, void* stream) {
    ODIL_REQUIRE(V && coef && d && k >= 1 && k <= kMdMaxRows && count >= 0 && ld >= count, "multi_axpy: bad arguments");
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ODIL_B200_F32)
        k_multi_axpy<float><<<blocks_for(count, 2), 256, 0, st>>>((const float*)V, ld, k, coef, a0, (const float*)g,
                                                                  (float*)d, count);
    else if (dtype == ODIL_B200_F64)
        k_multi_axpy<double><<<blocks_for(count, 2), 256, 0, st>>>((const double*)V, ld, k, coef, a0, (const double*)g,
                                                                   (double*)d, count);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    return 0;
}

}  // extern "C"
