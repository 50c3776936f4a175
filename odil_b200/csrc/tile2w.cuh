// k_tile2w: fused residual + loss + adjoint gradient on 2-D grids with ANY offset set of radius <= 2, one WARP per
// work item and no synchronisation between warps.
//
// k_tile2d (tile2d.cuh) stages a 32 x 64 tile behind CTA barriers with index-wrapped scalar loads and looks every
// coefficient up in shared memory: ncu shows 85 M warp instructions for 2048 x 4096 cells at 86 % of the issue slots
// (92 us = 17 % of the HBM roofline).  This kernel needs 25 M (measured: 59 us = 27 %; 4096^2 star: 88 us = 35 % against
// 167 us; 1024^2: 16.5 us against 17.4 us -- both are launch- and latency-bound there).  Here:
//   * a warp owns a strip of 120 columns (lanes 1..30, four consecutive cells each; lanes 0 and 31 carry the four halo
//     columns on either side: twice the largest column radius) and a chunk of rows, and MARCHES down the rows.  Per
//     step it stages one new U row (one 16-byte load per lane, issued PF steps ahead), computes the F row H0 rows
//     behind it and the g row 2*H0 rows behind it;
//   * U rows and F rows live in two rings of 2*H0 + 1 rows each, PRIVATE to the warp in shared memory, stored cell-major
//     ([j][lane]: the j-th cells of all lanes are contiguous), so a neighbour at any column distance is one
//     conflict-free LDS at `ring slot + constant`; the constants (per offset and cell) sit in the kernel parameters,
//     i.e. in the constant bank.  Only __syncwarp() orders the two phases: a slow warp delays nobody;
//   * rows of F that neighbouring chunks / strips need are recomputed, never exchanged (2*H0 extra rows per chunk,
//     8 extra columns per strip: they hit L1 / L2);
//   * warps whose cells (F) / source cells (g) all have the interior class take the coefficients from registers; any
//     other warp reads them per cell and offset from the table (__ldg, L1-resident).
// Per cell the operations and their order are those of k_tile2d / k_generic (c first, then the offsets in table order,
// fused multiply-adds), so F and g are bit-identical to them; the loss partial is summed per lane and row in T, then
// in fp64.  Requires: wrap_free plan (F is 0 outside the array and U reads as 0 there), <= 8 offsets, radii <= 2,
// N1 % 4 == 0, 16-byte aligned arrays, no slab.
// Reference: ctx.field() = roll (core.py:910-975), the operator's arithmetic (examples/poisson/poisson.py:57-68,
// 100-113; examples/wave/wave.py:29-75), mean(square(F)) (core.py:1093) and its reverse-mode gradient (core.py:1100).
#pragma once
#include "tile2d.cuh"
#include "tile3t.cuh"

namespace odil {

constexpr int kT2wN = 8;                   // offsets
constexpr int kT2wVW = 4;                  // cells per lane
constexpr int kT2wOwn = 30 * kT2wVW;       // owned columns per strip
constexpr int kT2wW = 34;                  // words per cell plane of a ring slot: lane + 1, one pad word on either side
constexpr int kT2wSlot = kT2wVW * kT2wW;   // words per ring slot
constexpr int kT2wWarps = 4;               // warps per CTA
constexpr int kT2wHM = 2;                  // largest radius

template <typename T>
struct Tile2wParams {
    const T* U;
    const T* c;        // nullable
    T* G;
    T* Fout;           // nullable
    const T* table;    // [ncls][noff]
    double* partials;  // one per CTA
    double* sumsq;     // sum of the partials, written by the last CTA to finish
    unsigned* counter; // CTAs done (zero between launches: the last CTA's atomicInc wraps it)
    T scale;
    int N0, N1;
    int R0, R1;
    int H0, H1;
    int noff, ncls;
    int nstrips, rows_per_chunk, nitems;
    signed char dy[kT2wN], dx[kT2wN];
    int kU[kT2wN][kT2wVW];  // byte offset of U[.][x_j + dx_o] from (cell plane 0, own lane) of a ring slot
    int kF[kT2wN][kT2wVW];  // byte offset of F[.][x_j - dx_o]
    int rU[kT2wN];          // ring slot of U row jf + dy_o relative to the newest staged row (jf + H0): dy_o - H0 <= 0
    int rF[kT2wN];          // ring slot of F row k - dy_o relative to the newest F row (k + H0): -H0 - dy_o <= 0
};

template <typename T>
struct T2wPack {
    T v[kT2wVW];
};

template <typename T>
__device__ __forceinline__ T2wPack<T> t2w_zero() {
    T2wPack<T> z;
#pragma unroll
    for (int j = 0; j < kT2wVW; ++j) z.v[j] = T(0);
    return z;
}
__device__ __forceinline__ T2wPack<float> t2w_ldg(const float* p) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    T2wPack<float> r;
    r.v[0] = q.x, r.v[1] = q.y, r.v[2] = q.z, r.v[3] = q.w;
    return r;
}
__device__ __forceinline__ T2wPack<double> t2w_ldg(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    T2wPack<double> r;
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = b.x, r.v[3] = b.y;
    return r;
}
__device__ __forceinline__ void t2w_stg(float* p, const T2wPack<float>& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ void t2w_stg(double* p, const T2wPack<double>& r) {
    reinterpret_cast<double2*>(p)[0] = make_double2(r.v[0], r.v[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(r.v[2], r.v[3]);
}

// both rings of one warp: 2 * (2*H0 + 1) rows
template <typename T>
inline size_t t2w_smem_bytes(int warps, int H0) {
    return (size_t)warps * 2 * (2 * H0 + 1) * kT2wSlot * sizeof(T);
}

// PF: rows of U (and of c) in flight per warp, one register buffer each; the row loop is unrolled PF times so that the
// buffers are never copied (a copy would wait for the load it copies -- ncu showed exactly that for a conditional
// refill).  WARPS: warps per CTA.  MINB: CTAs per SM the register allocation must allow.
template <typename T, int NOFF, int PF, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) k_tile2w(const __grid_constant__ Tile2wParams<T> p) {
    constexpr int S = (int)sizeof(T);
    constexpr uint32_t SLOTB = kT2wSlot * S, PLANEB = kT2wW * S;
    extern __shared__ __align__(16) unsigned char t2w_raw[];
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int noff = NOFF;  // the offset loops are unrolled exactly (one instantiation per offset count)
    const int N0 = p.N0, N1 = p.N1, H0 = p.H0, H1 = p.H1;
    const int NR = 2 * H0 + 1;  // rows per ring
    T* ring = reinterpret_cast<T*>(t2w_raw) + (size_t)warp * 2 * NR * kT2wSlot;
    for (int i = lane; i < 2 * NR * kT2wSlot; i += 32) ring[i] = T(0);  // pads included
    __syncwarp();
    // shared byte address of (slot 0, cell plane 0, own lane) of the U ring / the F ring
    const uint32_t sU = smem_u32(ring) + (uint32_t)((lane + 1) * S);
    const uint32_t sF = sU + (uint32_t)NR * SLOTB;
    const int C1 = 2 * p.R1 + 1;
    const int item = blockIdx.x * WARPS + warp;
    double acc = 0.0;
    if (item < p.nitems) {
        const int strip = item % p.nstrips, chunk = item / p.nstrips;
        const int ys = chunk * p.rows_per_chunk, ye = min(ys + p.rows_per_chunk, N0);
        const int x0 = strip * kT2wOwn + kT2wVW * (lane - 1);
        const bool xin = x0 >= 0 && x0 < N1;           // N1 % 4 == 0: the four columns are inside or outside together
        const bool own = xin && lane >= 1 && lane <= 30;
        int ccls[kT2wVW];
#pragma unroll
        for (int j = 0; j < kT2wVW; ++j) ccls[j] = xin ? t2_class(x0 + j, N1, p.R1) : 0;
        const unsigned full = 0xffffffffu;
        const bool fastXF = __all_sync(full, !xin || (x0 >= p.R1 && x0 + kT2wVW - 1 < N1 - p.R1));
        const bool fastXG = __all_sync(full, !own || (x0 >= p.R1 + H1 && x0 + kT2wVW - 1 < N1 - p.R1 - H1));
        T wi[NOFF];
        {
            const int CI = p.R0 * C1 + p.R1;
#pragma unroll
            for (int o = 0; o < NOFF; ++o) wi[o] = __ldg(p.table + CI * noff + o);
        }
        const T* ucol = p.U + (xin ? x0 : 0);
        const T* ccol = p.c ? p.c + (xin ? x0 : 0) : nullptr;
        auto load_u = [&](int r) {
            return (xin && r >= 0 && r < N0) ? t2w_ldg(ucol + (int64_t)r * N1) : t2w_zero<T>();
        };
        auto load_c = [&](int r) {  // c of F row r (rows this warp computes only)
            return (ccol && xin && r >= 0 && r < N0 && r < ye + H0) ? t2w_ldg(ccol + (int64_t)r * N1) : t2w_zero<T>();
        };
        auto stage_u = [&](const T2wPack<T>& u, int slot) {
            const uint32_t su = sU + (uint32_t)slot * SLOTB;
#pragma unroll
            for (int j = 0; j < kT2wVW; ++j) t3t_sts(su + j * PLANEB, u.v[j]);
        };
        // prologue: the 2*H0 rows below the first F row's newest row (ring slots 0 .. 2*H0 - 1); the main loop then
        // stages row ru, computes F row ru - H0 and g row ru - 2*H0 in EVERY step (no conditional refill of a buffer)
        const int ru1 = ye - 1 + 2 * H0;
        T2wPack<T> un[PF], cn[PF];
        {
            T2wPack<T> pro[2 * kT2wHM];  // all loads of the prologue are issued before the first one is consumed
#pragma unroll
            for (int q = 0; q < 2 * kT2wHM; ++q) pro[q] = q < 2 * H0 ? load_u(ys - 2 * H0 + q) : t2w_zero<T>();
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                un[q] = load_u(ys + q);
                cn[q] = load_c(ys - H0 + q);
            }
#pragma unroll
            for (int q = 0; q < 2 * kT2wHM; ++q)
                if (q < 2 * H0) stage_u(pro[q], q);
        }
        int cu = NR - 1;  // ring slot of the newest U row (row ru) and of the newest F row (row ru - H0): both 2*H0
        auto step = [&](const int ru, T2wPack<T>& ubuf, T2wPack<T>& cbuf) {
            stage_u(ubuf, cu);
            ubuf = load_u(ru + PF);
            __syncwarp();
            // ---------------- F row jf = ru - H0 from the U rows jf - H0 .. jf + H0
            const int jf = ru - H0;
            {
                const bool rin = jf >= 0 && jf < N0;
                T f[kT2wVW];
#pragma unroll
                for (int j = 0; j < kT2wVW; ++j) f[j] = cbuf.v[j];
                cbuf = load_c(jf + PF);
                if (rin) {
                    uint32_t so[NOFF];
#pragma unroll
                    for (int o = 0; o < NOFF; ++o) {
                        int t = cu + p.rU[o];
                        t += t < 0 ? NR : 0;
                        so[o] = sU + (uint32_t)t * SLOTB;
                    }
                    if (fastXF && jf >= p.R0 && jf < N0 - p.R0) {
#pragma unroll
                        for (int o = 0; o < NOFF; ++o) {
#pragma unroll
                            for (int j = 0; j < kT2wVW; ++j)
                                f[j] = fma(wi[o], t3t_lds(so[o] + (uint32_t)p.kU[o][j], (T*)nullptr), f[j]);
                        }
                    } else {
                        const int rc = t2_class(jf, N0, p.R0) * C1;
#pragma unroll
                        for (int o = 0; o < NOFF; ++o) {
#pragma unroll
                            for (int j = 0; j < kT2wVW; ++j) {
                                const T w = __ldg(p.table + (rc + ccls[j]) * noff + o);
                                f[j] = fma(w, t3t_lds(so[o] + (uint32_t)p.kU[o][j], (T*)nullptr), f[j]);
                            }
                        }
                    }
                }
                if (!(rin && xin)) {
#pragma unroll
                    for (int j = 0; j < kT2wVW; ++j) f[j] = T(0);
                }
                if (own && jf >= ys && jf < ye) {
                    T accp = T(0);
#pragma unroll
                    for (int j = 0; j < kT2wVW; ++j) accp = fma(f[j], f[j], accp);
                    acc += (double)accp;
                    if (p.Fout) {
                        T2wPack<T> fo;
#pragma unroll
                        for (int j = 0; j < kT2wVW; ++j) fo.v[j] = f[j];
                        t2w_stg(p.Fout + (int64_t)jf * N1 + x0, fo);
                    }
                }
                const uint32_t sf = sF + (uint32_t)cu * SLOTB;
#pragma unroll
                for (int j = 0; j < kT2wVW; ++j) t3t_sts(sf + j * PLANEB, f[j]);
            }
            __syncwarp();
            // ---------------- g row k = ru - 2 H0 from the F rows k - H0 .. k + H0
            const int k = ru - 2 * H0;
            if (k >= ys) {
                T g[kT2wVW];
#pragma unroll
                for (int j = 0; j < kT2wVW; ++j) g[j] = T(0);
                uint32_t so[NOFF];
#pragma unroll
                for (int o = 0; o < NOFF; ++o) {
                    int t = cu + p.rF[o];
                    t += t < 0 ? NR : 0;
                    so[o] = sF + (uint32_t)t * SLOTB;
                }
                if (fastXG && k >= p.R0 + H0 && k < N0 - p.R0 - H0) {
#pragma unroll
                    for (int o = 0; o < NOFF; ++o) {
#pragma unroll
                        for (int j = 0; j < kT2wVW; ++j)
                            g[j] = fma(wi[o], t3t_lds(so[o] + (uint32_t)p.kF[o][j], (T*)nullptr), g[j]);
                    }
                } else {
                    // coefficient of the SOURCE cell's class; sources outside the array have F = 0 and are skipped
#pragma unroll
                    for (int o = 0; o < NOFF; ++o) {
                        const int sy = k - p.dy[o];
                        if (sy >= 0 && sy < N0) {
                            const int rc = t2_class(sy, N0, p.R0) * C1;
#pragma unroll
                            for (int j = 0; j < kT2wVW; ++j) {
                                const int sx = x0 + j - p.dx[o];
                                const int cx = (sx >= 0 && sx < N1) ? t2_class(sx, N1, p.R1) : 0;
                                const T w = __ldg(p.table + (rc + cx) * noff + o);
                                g[j] = fma(w, t3t_lds(so[o] + (uint32_t)p.kF[o][j], (T*)nullptr), g[j]);
                            }
                        }
                    }
                }
                if (own) {
                    T2wPack<T> go;
#pragma unroll
                    for (int j = 0; j < kT2wVW; ++j) go.v[j] = g[j] * p.scale;
                    t2w_stg(p.G + (int64_t)k * N1 + x0, go);
                }
            }
            cu = cu + 1 == NR ? 0 : cu + 1;
        };
        for (int ru = ys; ru <= ru1; ru += PF) {
#pragma unroll
            for (int q = 0; q < PF; ++q)
                if (ru + q <= ru1) step(ru + q, un[q], cn[q]);
        }
    }
    const double s = block_sum(acc, red);
    // Second stage inside the same launch: every CTA publishes its partial, the last one to arrive sums them all in a
    // fixed order (thread t takes partials t, t + 128, ...; then the block tree) -- deterministic, no extra launch.
    __shared__ bool last;
    if (threadIdx.x == 0) {
        p.partials[blockIdx.x] = s;
        __threadfence();
        last = atomicInc(p.counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32 * WARPS) v += __ldcg(p.partials + i);
        v = block_sum(v, red);
        if (threadIdx.x == 0) p.sumsq[0] = v;
    }
}

}  // namespace odil
