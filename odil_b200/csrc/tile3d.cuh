// k_tile3d: fused residual + loss + adjoint gradient for 3-D grids with ANY offset set of small radius (sm_100a).
//
// The wave operator in two space dimensions (BASELINE configs[2], (t, x, y) grid) couples (t,x,y), (t-1,x,y),
// (t-2,x,y), (t-1,x+-1,y), (t-1,x,y+-1): not a star, so k_star8 does not apply.  This kernel is k_tile2d carried along
// axis 0: a CTA owns a 16 x 64 tile of (axis 1, axis 2) and marches over a chunk of planes.  Two rings of 2*H0+1
// planes live in shared memory (H0 = stencil radius along axis 0): the U tile with a halo of twice the in-plane
// radius, and the F tile with one radius plus the class byte of every cell.  Per plane j:
//     stage U plane j+H0  ->  F plane j from U planes j-H0..j+H0  ->  g plane j-H0 from F planes j-2H0..j
// so U and c are read once per chunk (plus 4*H0 lead-in planes), g is written once and F never leaves the chip.
// Periodic wrap along all three axes is resolved while staging; ring slots are addressed by the unwrapped plane
// number.  Summation order is the one of k_generic.  tests/test_tile_emulation_cpu.py restates the addressing in NumPy.
//
// Default for non-star 3-D plans on one GPU (odil_b200_stencil_plan_tune variant 81 or ODIL_B200_TILE3D=0 selects
// k_generic).  First measurement (128 x 256 x 256 fp32, wave footprint): 0.307 ms vs 1.791 ms in k_generic; the loop is
// latency-bound (staging, F and g of a plane run back to back with two barriers and no prefetch of the next plane).
#pragma once
#include "tile2d.cuh"

namespace odil {

constexpr int kT3Y = 16, kT3X = 64, kT3Threads = 256, kT3MaxRadius = 2;

template <typename T>
struct Tile3Params {
    const T* U;
    const T* c;       // nullable
    T* G;
    T* Fout;          // nullable
    const T* table;   // [ncls][noff]
    double* partials; // one per CTA
    T scale;
    int N0, N1, N2;
    int R0, R1, R2;
    int H0, H1, H2;
    int noff, ncls;
    int zchunk;
    unsigned magicA, magicF;
    signed char dz[ODIL_B200_MAX_OFFSETS], dy[ODIL_B200_MAX_OFFSETS], dx[ODIL_B200_MAX_OFFSETS];
};

struct Tile3Dims {
    int AH, AW, FH, FW, NR;
};

__host__ __device__ inline Tile3Dims t3_dims(int H0, int H1, int H2) {
    Tile3Dims d;
    d.AH = kT3Y + 4 * H1;
    d.AW = kT3X + 4 * H2;
    d.FH = kT3Y + 2 * H1;
    d.FW = kT3X + 2 * H2;
    d.NR = 2 * H0 + 1;
    return d;
}

template <typename T>
inline size_t t3_smem_bytes(int H0, int H1, int H2, int ncls, int noff) {
    const Tile3Dims d = t3_dims(H0, H1, H2);
    size_t n = (size_t)d.NR * (d.AH * d.AW + d.FH * d.FW) * sizeof(T);
    n += (size_t)ncls * noff * sizeof(T);
    n += 4 * ODIL_B200_MAX_OFFSETS * sizeof(int);
    n += (size_t)d.NR * d.FH * d.FW;  // class bytes
    return n + 16;
}

__device__ __forceinline__ int t3_slot(int p, int nr) {  // non-negative p mod nr for small |p| / nr ranges
    p %= nr;
    return p < 0 ? p + nr : p;
}

template <typename T>
__global__ void __launch_bounds__(kT3Threads) k_tile3d(const __grid_constant__ Tile3Params<T> p) {
    extern __shared__ __align__(16) unsigned char t3_smem[];
    __shared__ double red[32];
    const Tile3Dims d = t3_dims(p.H0, p.H1, p.H2);
    const int AHW = d.AH * d.AW, FHW = d.FH * d.FW, NR = d.NR;
    T* sU = reinterpret_cast<T*>(t3_smem);                  // [NR][AHW]
    T* sF = sU + NR * AHW;                                  // [NR][FHW]
    T* sTab = sF + NR * FHW;                                // [ncls][noff]
    int* sOU = reinterpret_cast<int*>(sTab + p.ncls * p.noff);  // [2][MAX_OFFSETS], by plane parity
    int* sOF = sOU + 2 * ODIL_B200_MAX_OFFSETS;                 // [2][MAX_OFFSETS]
    unsigned char* sC = reinterpret_cast<unsigned char*>(sOF + 2 * ODIL_B200_MAX_OFFSETS);  // [NR][FHW]
    const int tid = threadIdx.x;
    const int ty0 = blockIdx.y * kT3Y, tx0 = blockIdx.x * kT3X;
    const int zs = blockIdx.z * p.zchunk, ze = min(zs + p.zchunk, p.N0);
    const int N0 = p.N0, N1 = p.N1, N2 = p.N2, noff = p.noff;
    const int H0 = p.H0, H1 = p.H1, H2 = p.H2;
    const int C1 = 2 * p.R1 + 1, C2 = 2 * p.R2 + 1;
    const int64_t plane = (int64_t)N1 * N2;

    for (int i = tid; i < p.ncls * noff; i += kT3Threads) sTab[i] = p.table[i];

    auto stage = [&](int pz) {  // U plane pz (unwrapped number) -> its ring slot
        const T* src = p.U + (int64_t)t2_wrap(pz, N0) * plane;
        T* dst = sU + t3_slot(pz, NR) * AHW;
        for (int e = tid; e < AHW; e += kT3Threads) {
            const int r = (int)__umulhi((unsigned)e, p.magicA);
            const int cc = e - r * d.AW;
            dst[e] = src[(int64_t)t2_wrap(ty0 - 2 * H1 + r, N1) * N2 + t2_wrap(tx0 - 2 * H2 + cc, N2)];
        }
    };

    const int j0 = zs - H0, j1 = ze - 1 + H0;
    for (int pz = j0 - H0; pz < j0 + H0; ++pz) stage(pz);
    double acc = 0.0;
    for (int j = j0; j <= j1; ++j) {
        const int k = j - H0;  // plane whose gradient becomes computable in this step
        stage(j + H0);
        int* oU = sOU + (j & 1) * ODIL_B200_MAX_OFFSETS;
        int* oF = sOF + (j & 1) * ODIL_B200_MAX_OFFSETS;
        if (tid < noff) {
            oU[tid] = t3_slot(j + p.dz[tid], NR) * AHW + p.dy[tid] * d.AW + p.dx[tid];
            oF[tid] = t3_slot(k - p.dz[tid], NR) * FHW - (p.dy[tid] * d.FW + p.dx[tid]);
        }
        __syncthreads();
        {   // F plane j on the tile plus one in-plane radius
            const int gz = t2_wrap(j, N0);
            const int czc = t2_class(gz, N0, p.R0) * C1;
            const bool own_plane = j >= zs && j < ze;
            T* fdst = sF + t3_slot(j, NR) * FHW;
            unsigned char* cdst = sC + t3_slot(j, NR) * FHW;
            const T* csrc = p.c ? p.c + (int64_t)gz * plane : nullptr;
            for (int e = tid; e < FHW; e += kT3Threads) {
                const int r = (int)__umulhi((unsigned)e, p.magicF);
                const int cc = e - r * d.FW;
                const int ly = ty0 - H1 + r, lx = tx0 - H2 + cc;
                const int gy = t2_wrap(ly, N1), gx = t2_wrap(lx, N2);
                const int cls = (czc + t2_class(gy, N1, p.R1)) * C2 + t2_class(gx, N2, p.R2);
                const T* trow = sTab + cls * noff;
                const int at = (r + H1) * d.AW + cc + H2;
                T f = csrc ? csrc[(int64_t)gy * N2 + gx] : T(0);
                for (int o = 0; o < noff; ++o) f += trow[o] * sU[oU[o] + at];
                fdst[e] = f;
                cdst[e] = (unsigned char)cls;
                if (own_plane && r >= H1 && r < H1 + kT3Y && cc >= H2 && cc < H2 + kT3X && ly < N1 && lx < N2) {
                    acc += (double)f * (double)f;
                    if (p.Fout) p.Fout[(int64_t)j * plane + (int64_t)ly * N2 + lx] = f;
                }
            }
        }
        __syncthreads();
        if (k >= zs) {  // g plane k from F planes k-H0 .. k+H0 (= j)
            T* gdst = p.G + (int64_t)k * plane;
            for (int e = tid; e < kT3Y * kT3X; e += kT3Threads) {
                const int r = e / kT3X, cc = e % kT3X;
                const int y = ty0 + r, x = tx0 + cc;
                if (y >= N1 || x >= N2) continue;
                const int at = (r + H1) * d.FW + cc + H2;
                T g = T(0);
                for (int o = 0; o < noff; ++o) {
                    const int i = oF[o] + at;
                    g += sTab[(int)sC[i] * noff + o] * sF[i];
                }
                gdst[(int64_t)y * N2 + x] = g * p.scale;
            }
        }
    }
    const double s = block_sum(acc, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------------------------
// k_tile3d8: the same sweep for plans with at most 8 offsets (the wave footprint has 7), cut from ~530 to ~70
// thread-instructions per cell (ncu on k_tile3d: 3 % DRAM, 39 % SM throughput -- bound by instruction issue):
//   * the offset loop is unrolled; the ring-slot offsets of a plane are read from shared memory once per thread and
//     plane instead of once per cell and offset;
//   * cells whose own class (F) / whose sources' classes (g) are interior on all axes use the interior coefficient row
//     held in registers -- no class computation, no class-byte or table loads.  Whether a tile position is "deep
//     interior" in the plane is decided once per thread (positions are fixed across planes); whether plane j / k is, once
//     per plane.  Everything else runs the per-cell table path of k_tile3d, statement for statement.
// Same summation order per cell as k_tile3d and k_generic: results are bit-identical.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kT3N = 8;

template <typename T>
__global__ void __launch_bounds__(kT3Threads) k_tile3d8(const __grid_constant__ Tile3Params<T> p) {
    extern __shared__ __align__(16) unsigned char t3_smem[];
    __shared__ double red[32];
    const Tile3Dims d = t3_dims(p.H0, p.H1, p.H2);
    const int AHW = d.AH * d.AW, FHW = d.FH * d.FW, NR = d.NR;
    T* sU = reinterpret_cast<T*>(t3_smem);
    T* sF = sU + NR * AHW;
    T* sTab = sF + NR * FHW;
    int* sOU = reinterpret_cast<int*>(sTab + p.ncls * p.noff);
    int* sOF = sOU + 2 * ODIL_B200_MAX_OFFSETS;
    unsigned char* sC = reinterpret_cast<unsigned char*>(sOF + 2 * ODIL_B200_MAX_OFFSETS);
    const int tid = threadIdx.x;
    const int ty0 = blockIdx.y * kT3Y, tx0 = blockIdx.x * kT3X;
    const int zs = blockIdx.z * p.zchunk, ze = min(zs + p.zchunk, p.N0);
    const int N0 = p.N0, N1 = p.N1, N2 = p.N2, noff = p.noff;
    const int H0 = p.H0, H1 = p.H1, H2 = p.H2;
    const int C1 = 2 * p.R1 + 1, C2 = 2 * p.R2 + 1;
    const int64_t plane = (int64_t)N1 * N2;
    const int CI = (p.R0 * C1 + p.R1) * C2 + p.R2;  // class of a cell that is interior on every axis

    for (int i = tid; i < p.ncls * noff; i += kT3Threads) sTab[i] = p.table[i];
    T wi[kT3N];
#pragma unroll
    for (int o = 0; o < kT3N; ++o) wi[o] = o < noff ? p.table[CI * noff + o] : T(0);

    // deep-interior flags of this thread's tile positions (bit i = i-th element it handles in the F / g loops)
    unsigned fastF = 0, fastG = 0;
    for (int e = tid, i = 0; e < FHW; e += kT3Threads, ++i) {
        const int r = (int)__umulhi((unsigned)e, p.magicF);
        const int cc = e - r * d.FW;
        const int ly = ty0 - H1 + r, lx = tx0 - H2 + cc;
        if (ly >= p.R1 && ly < N1 - p.R1 && lx >= p.R2 && lx < N2 - p.R2) fastF |= 1u << i;
    }
    for (int e = tid, i = 0; e < kT3Y * kT3X; e += kT3Threads, ++i) {
        const int y = ty0 + e / kT3X, x = tx0 + e % kT3X;
        if (y >= p.R1 + H1 && y < N1 - p.R1 - H1 && x >= p.R2 + H2 && x < N2 - p.R2 - H2) fastG |= 1u << i;
    }

    auto stage = [&](int pz) {
        const T* src = p.U + (int64_t)t2_wrap(pz, N0) * plane;
        T* dst = sU + t3_slot(pz, NR) * AHW;
        for (int e = tid; e < AHW; e += kT3Threads) {
            const int r = (int)__umulhi((unsigned)e, p.magicA);
            const int cc = e - r * d.AW;
            dst[e] = src[(int64_t)t2_wrap(ty0 - 2 * H1 + r, N1) * N2 + t2_wrap(tx0 - 2 * H2 + cc, N2)];
        }
    };

    const int j0 = zs - H0, j1 = ze - 1 + H0;
    for (int pz = j0 - H0; pz < j0 + H0; ++pz) stage(pz);
    double acc = 0.0;
    for (int j = j0; j <= j1; ++j) {
        const int k = j - H0;
        stage(j + H0);
        int* oU = sOU + (j & 1) * ODIL_B200_MAX_OFFSETS;
        int* oF = sOF + (j & 1) * ODIL_B200_MAX_OFFSETS;
        if (tid < kT3N) {  // entries beyond noff point at slot 0 and meet a zero coefficient
            const bool on = tid < noff;
            oU[tid] = on ? t3_slot(j + p.dz[tid], NR) * AHW + p.dy[tid] * d.AW + p.dx[tid] : 0;
            oF[tid] = on ? t3_slot(k - p.dz[tid], NR) * FHW - (p.dy[tid] * d.FW + p.dx[tid]) : 0;
        }
        __syncthreads();
        {   // F plane j on the tile plus one in-plane radius
            int ou[kT3N];
#pragma unroll
            for (int o = 0; o < kT3N; ++o) ou[o] = oU[o];
            const int gz = t2_wrap(j, N0);
            const bool zfast = j >= p.R0 && j < N0 - p.R0;  // in range (no wrap) and interior z class
            const int czc = t2_class(gz, N0, p.R0) * C1;
            const bool own_plane = j >= zs && j < ze;
            T* fdst = sF + t3_slot(j, NR) * FHW;
            unsigned char* cdst = sC + t3_slot(j, NR) * FHW;
            const T* csrc = p.c ? p.c + (int64_t)gz * plane : nullptr;
            for (int e = tid, i = 0; e < FHW; e += kT3Threads, ++i) {
                const int r = (int)__umulhi((unsigned)e, p.magicF);
                const int cc = e - r * d.FW;
                const int ly = ty0 - H1 + r, lx = tx0 - H2 + cc;
                const int at = (r + H1) * d.AW + cc + H2;
                T f;
                if (zfast && ((fastF >> i) & 1u)) {
                    f = csrc ? csrc[(int64_t)ly * N2 + lx] : T(0);
#pragma unroll
                    for (int o = 0; o < kT3N; ++o)
                        if (o < noff) f += wi[o] * sU[ou[o] + at];
                    cdst[e] = (unsigned char)CI;
                } else {
                    const int gy = t2_wrap(ly, N1), gx = t2_wrap(lx, N2);
                    const int cls = (czc + t2_class(gy, N1, p.R1)) * C2 + t2_class(gx, N2, p.R2);
                    const T* trow = sTab + cls * noff;
                    f = csrc ? csrc[(int64_t)gy * N2 + gx] : T(0);
                    for (int o = 0; o < noff; ++o) f += trow[o] * sU[ou[o] + at];
                    cdst[e] = (unsigned char)cls;
                }
                fdst[e] = f;
                if (own_plane && r >= H1 && r < H1 + kT3Y && cc >= H2 && cc < H2 + kT3X && ly < N1 && lx < N2) {
                    acc += (double)f * (double)f;
                    if (p.Fout) p.Fout[(int64_t)j * plane + (int64_t)ly * N2 + lx] = f;
                }
            }
        }
        __syncthreads();
        if (k >= zs) {  // g plane k from F planes k-H0 .. k+H0 (= j)
            int of[kT3N];
#pragma unroll
            for (int o = 0; o < kT3N; ++o) of[o] = oF[o];
            const bool zfast = k >= p.R0 + H0 && k < N0 - p.R0 - H0;  // every source plane: in range, interior z class
            T* gdst = p.G + (int64_t)k * plane;
            for (int e = tid, i = 0; e < kT3Y * kT3X; e += kT3Threads, ++i) {
                const int r = e / kT3X, cc = e % kT3X;
                const int y = ty0 + r, x = tx0 + cc;
                if (y >= N1 || x >= N2) continue;
                const int at = (r + H1) * d.FW + cc + H2;
                T g = T(0);
                if (zfast && ((fastG >> i) & 1u)) {
#pragma unroll
                    for (int o = 0; o < kT3N; ++o)
                        if (o < noff) g += wi[o] * sF[of[o] + at];
                } else {
                    for (int o = 0; o < noff; ++o) {
                        const int ii = of[o] + at;
                        g += sTab[(int)sC[ii] * noff + o] * sF[ii];
                    }
                }
                gdst[(int64_t)y * N2 + x] = g * p.scale;
            }
        }
    }
    const double s = block_sum(acc, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

}  // namespace odil
