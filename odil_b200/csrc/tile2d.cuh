// k_tile2d: region-typed affine stencil on 2-D grids with ANY offset set of small radius (sm_100a).
//
// The reference's 2-D examples are not all star stencils: examples/wave/wave.py:29-75 couples
// (t,x), (t-1,x), (t-2,x), (t-1,x-1), (t-1,x+1).  One CTA owns a TY x TX tile.  Fused mode stages the U tile with a
// halo of twice the stencil radius in shared memory, evaluates F = A U + c on the tile plus one radius (class row
// of every cell looked up once and kept as a byte), accumulates the squared loss of the owned cells and applies
// the transpose from the F tile: U and c are read once from HBM (halo re-reads hit L2), g is written once, F never
// leaves the chip.  Forward-only and adjoint-only modes (Newton products, multi-field outputs) use the same tiles
// with one radius of halo.  Periodic wrap (ctx.field = roll, core.py:963) is applied while staging, so the inner
// loops index shared memory only.  Summation order is the one of k_generic (c first, offsets in table order).
#pragma once
#include "common.cuh"

namespace odil {

constexpr int kT2Y = 32, kT2X = 64, kT2Threads = 256, kT2MaxRadius = 4;

template <typename T>
struct Tile2Params {
    const T* A;       // fused / forward: U;  adjoint: F
    const T* c;       // fused: constant term;  forward: F_in;  adjoint: G_in   (nullable)
    T* out;           // forward: F_out;  adjoint / fused: G_out
    T* Fout;          // fused: optional store of F
    const T* table;   // [ncls][noff]
    double* partials; // fused: one per CTA
    T scale;
    int N0, N1;       // rows (axis 0), columns (axis 1, contiguous)
    int R0, R1;       // region half-widths
    int H0, H1;       // stencil radius per axis
    int noff, ncls;
    unsigned magicA, magicF;  // ceil(2^32 / tile pitch): row = umulhi(e, magic) for e < 2^16
    signed char dy[ODIL_B200_MAX_OFFSETS], dx[ODIL_B200_MAX_OFFSETS];
};

__device__ __forceinline__ int t2_wrap(int i, int n) {
    if ((unsigned)i >= (unsigned)n) {
        i %= n;
        if (i < 0) i += n;
    }
    return i;
}

__device__ __forceinline__ int t2_class(int i, int n, int r) {
    if (i < r) return i;
    const int d = n - 1 - i;
    return d < r ? 2 * r - d : r;
}

template <int MODE>
__host__ __device__ inline void t2_dims(int H0, int H1, int& AH, int& AW, int& FH, int& FW) {
    const int a0 = MODE == 2 ? 2 * H0 : H0, a1 = MODE == 2 ? 2 * H1 : H1;
    AH = kT2Y + 2 * a0;
    AW = kT2X + 2 * a1;
    FH = kT2Y + 2 * H0;
    FW = kT2X + 2 * H1;
}

template <typename T, int MODE>
inline size_t t2_smem_bytes(int H0, int H1, int ncls, int noff) {
    int AH, AW, FH, FW;
    t2_dims<MODE>(H0, H1, AH, AW, FH, FW);
    size_t n = (size_t)AH * AW * sizeof(T);
    if (MODE == 2) n += (size_t)FH * FW * sizeof(T);
    n += (size_t)ncls * noff * sizeof(T);
    n += 2 * ODIL_B200_MAX_OFFSETS * sizeof(int);
    if (MODE != 0) n += (size_t)FH * FW;  // class bytes
    return n + 16;
}

template <typename T, int MODE>  // 0 forward, 1 adjoint, 2 fused
__global__ void __launch_bounds__(kT2Threads) k_tile2d(const __grid_constant__ Tile2Params<T> p) {
    extern __shared__ __align__(16) unsigned char t2_smem[];
    __shared__ double red[32];
    int AH, AW, FH, FW;
    t2_dims<MODE>(p.H0, p.H1, AH, AW, FH, FW);
    const int HA0 = (AH - kT2Y) / 2, HA1 = (AW - kT2X) / 2;
    T* sA = reinterpret_cast<T*>(t2_smem);
    T* sF = sA + AH * AW;
    T* sTab = sF + (MODE == 2 ? FH * FW : 0);
    int* sDA = reinterpret_cast<int*>(sTab + p.ncls * p.noff);
    int* sDF = sDA + ODIL_B200_MAX_OFFSETS;
    unsigned char* sC = reinterpret_cast<unsigned char*>(sDF + ODIL_B200_MAX_OFFSETS);
    const int tid = threadIdx.x;
    const int ty0 = blockIdx.y * kT2Y, tx0 = blockIdx.x * kT2X;
    const int N0 = p.N0, N1 = p.N1, noff = p.noff;
    const int C1 = 2 * p.R1 + 1;

    for (int i = tid; i < p.ncls * noff; i += kT2Threads) sTab[i] = p.table[i];
    if (tid < noff) {
        sDA[tid] = p.dy[tid] * AW + p.dx[tid];
        sDF[tid] = p.dy[tid] * FW + p.dx[tid];
    }
    // stage the input tile (periodic wrap resolved here)
    for (int e = tid; e < AH * AW; e += kT2Threads) {
        const int r = (int)__umulhi((unsigned)e, p.magicA);
        const int cc = e - r * AW;
        const int gy = t2_wrap(ty0 - HA0 + r, N0), gx = t2_wrap(tx0 - HA1 + cc, N1);
        sA[e] = p.A[(int64_t)gy * N1 + gx];
        if (MODE == 1) sC[e] = (unsigned char)(t2_class(gy, N0, p.R0) * C1 + t2_class(gx, N1, p.R1));
    }
    __syncthreads();

    if (MODE == 0) {
        for (int e = tid; e < kT2Y * kT2X; e += kT2Threads) {
            const int r = e / kT2X, cc = e % kT2X;
            const int y = ty0 + r, x = tx0 + cc;
            if (y >= N0 || x >= N1) continue;
            const T* trow = sTab + (t2_class(y, N0, p.R0) * C1 + t2_class(x, N1, p.R1)) * noff;
            const T* a = sA + (r + HA0) * AW + cc + HA1;
            const int64_t lin = (int64_t)y * N1 + x;
            T f = p.c ? p.c[lin] : T(0);
            for (int o = 0; o < noff; ++o) f += trow[o] * a[sDA[o]];
            p.out[lin] = f;
        }
        return;
    }
    if (MODE == 1) {
        for (int e = tid; e < kT2Y * kT2X; e += kT2Threads) {
            const int r = e / kT2X, cc = e % kT2X;
            const int y = ty0 + r, x = tx0 + cc;
            if (y >= N0 || x >= N1) continue;
            const int at = (r + HA0) * AW + cc + HA1;
            T g = T(0);
            for (int o = 0; o < noff; ++o) {
                const int j = at - sDA[o];
                g += sTab[(int)sC[j] * noff + o] * sA[j];
            }
            g *= p.scale;
            const int64_t lin = (int64_t)y * N1 + x;
            if (p.c) g += p.c[lin];
            p.out[lin] = g;
        }
        return;
    }

    // fused: F on the tile plus one radius
    double acc = 0.0;
    for (int e = tid; e < FH * FW; e += kT2Threads) {
        const int r = (int)__umulhi((unsigned)e, p.magicF);
        const int cc = e - r * FW;
        const int ly = ty0 - p.H0 + r, lx = tx0 - p.H1 + cc;
        const int gy = t2_wrap(ly, N0), gx = t2_wrap(lx, N1);
        const int cls = t2_class(gy, N0, p.R0) * C1 + t2_class(gx, N1, p.R1);
        const T* trow = sTab + cls * noff;
        const T* a = sA + (r + p.H0) * AW + cc + p.H1;
        T f = p.c ? p.c[(int64_t)gy * N1 + gx] : T(0);
        for (int o = 0; o < noff; ++o) f += trow[o] * a[sDA[o]];
        sF[e] = f;
        sC[e] = (unsigned char)cls;
        const bool owned = r >= p.H0 && r < p.H0 + kT2Y && cc >= p.H1 && cc < p.H1 + kT2X && ly < N0 && lx < N1;
        if (owned) {
            acc += (double)f * (double)f;
            if (p.Fout) p.Fout[(int64_t)ly * N1 + lx] = f;
        }
    }
    __syncthreads();
    for (int e = tid; e < kT2Y * kT2X; e += kT2Threads) {
        const int r = e / kT2X, cc = e % kT2X;
        const int y = ty0 + r, x = tx0 + cc;
        if (y >= N0 || x >= N1) continue;
        const int at = (r + p.H0) * FW + cc + p.H1;
        T g = T(0);
        for (int o = 0; o < noff; ++o) {
            const int j = at - sDF[o];
            g += sTab[(int)sC[j] * noff + o] * sF[j];
        }
        p.out[(int64_t)y * N1 + x] = g * p.scale;
    }
    const double s = block_sum(acc, red);
    if (tid == 0) p.partials[blockIdx.y * gridDim.x + blockIdx.x] = s;
}

}  // namespace odil
