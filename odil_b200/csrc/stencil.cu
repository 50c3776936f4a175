// Region-typed affine stencil  F = A U + c,  loss = sum F^2,  g = scale * A^T F   (sm_100a).
//
// Replaces, on the reference side: ctx.field()=roll (core.py:910-975), the example operators'
// arithmetic (examples/poisson/poisson.py:57-68,100-113; examples/wave/wave.py:29-75), the loss
// reduction (core.py:1093) and the AD gradient (core.py:1100-1101).
//
// Two kernels:
//   k_generic  - any ndim<=4, any offsets, per-cell class lookup; works on a list of boxes.  Used for
//                (a) whole-domain evaluation of stencils the tiled kernel does not cover and
//                (b) the boundary SHELL (cells within 2r of a face) after the tiled kernel.
//   k_star3d   - the hot kernel: 3-D (or 2-D) star stencil with the INTERIOR coefficient row only,
//                2.5-D marching along axis 0, F staged in a 4-slot shared-memory ring, forward
//                residual + squared-loss partial + adjoint gradient in ONE sweep (U and c read once,
//                g written once; F never touches HBM).
#include <algorithm>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "common.cuh"

namespace odil {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
std::atomic<int64_t>& launch_counter() {
    static std::atomic<int64_t> c{0};
    return c;
}

double* reduction_scratch(int nslots) {
    static double* buf[64] = {nullptr};
    static int cap[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (cap[dev] < nslots) {
        if (buf[dev]) cudaFree(buf[dev]);
        int n = std::max(nslots, 4 * kMaxPartialBlocks);
        if (cudaMalloc(&buf[dev], sizeof(double) * n) != cudaSuccess) return nullptr;
        cap[dev] = n;
    }
    return buf[dev];
}

constexpr int kMaxBoxes = 8;
constexpr int kPartialCapacity = 1 << 16;

struct BoxList {
    int nbox;
    int64_t lo[kMaxBoxes][ODIL_B200_MAX_NDIM];  // local coords (axis 0 relative to first owned plane)
    int64_t sz[kMaxBoxes][ODIL_B200_MAX_NDIM];
    int64_t start[kMaxBoxes + 1];  // prefix sums of cell counts
};

struct GenParams {
    int ndim, noff;
    int64_t shape[ODIL_B200_MAX_NDIM];   // global shape
    int64_t stride[ODIL_B200_MAX_NDIM];  // element strides of the local arrays
    int64_t n0, z0;
    int halo;
    int off[ODIL_B200_MAX_OFFSETS][ODIL_B200_MAX_NDIM];
    int R[ODIL_B200_MAX_NDIM];
    int cstride[ODIL_B200_MAX_NDIM];  // class-index strides
    int zero_off;                     // index of the all-zero offset, or -1
    int count_mode;                   // 0: every cell counts toward the loss; 1: only cells with a non-interior class
};

}  // namespace odil

using namespace odil;

struct odil_b200_plan {
    int ndim, dtype, noff;
    int64_t shape[ODIL_B200_MAX_NDIM];
    int off[ODIL_B200_MAX_OFFSETS][ODIL_B200_MAX_NDIM];
    int R[ODIL_B200_MAX_NDIM];
    int ncls;
    std::vector<double> table;  // host copy [ncls][noff]
    void* table_dev;            // typed copy
    double* partials;           // device, kPartialCapacity doubles
    int device;
    // tiled-kernel eligibility
    int kind;      // 0 generic, 1 star tiled
    double w[7];   // interior weights: c, zm, zp, ym, yp, xm, xp  (3-D naming; 2-D uses y,x)
    int zchunk;    // 0 => auto
    int variant;   // tile shape variant
    int rmax0;     // max |off| along axis 0
    void* star_table;  // device, typed [ncls][7] (c, zm, zp, ym, yp, xm, xp) when kind == 1
    int star_has_z;    // some class row couples axis 0 (even if the interior row does not)
    int wrap_free;     // no coefficient multiplies a neighbour across a periodic boundary (TMA zero fill is exact)
    int use_tma;       // 1: TMA-fed kernel when eligible
    int use_v3;        // 1: column-group kernel (default when N2 % 4 == 0), 0: v2 tile kernel + shell
    int use_star7;     // 1: k_star7 (default when eligible)
    int no_xu;         // tuning: force the general coefficient registers
    int use_star8;     // 1: k_star8 (default when eligible): one row per warp, host-built work list
    // work list of k_star8 (device copy + the key it was built for)
    mutable void* work_dev;
    mutable int work_cap, work_n, work_key[6];
    int use_tile2d;    // 1: k_tile2d for 2-D grids (default; env ODIL_B200_TILE2D=0 or an explicit variant disables)
    int h2[2];         // stencil radius per axis (2-D plans)
    int use_tile3d;    // 1: k_tile3d for non-star 3-D plans (default; ODIL_B200_TILE3D=0 or variant 81 selects k_generic)
    int h3[3];         // stencil radius per axis (3-D plans)
    int star_xu;       // z-/y-arm coefficients of the interior y/z classes do not depend on the x class
};

namespace odil {

// ------------------------------------------------------------------------------------------------
// Generic kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int axis_class(int64_t i, int64_t n, int r) {
    if (i < r) return (int)i;
    const int64_t d = n - 1 - i;
    if (d < r) return 2 * r - (int)d;
    return r;
}

template <typename T>
struct GenIO {
    const T* U;     // forward/fused: input field; adjoint: F
    const T* c;     // fused: constant term (nullable); forward: F_in; adjoint: G_in
    T* out;         // forward: F_out; adjoint/fused: G_out
    T* Fout;        // fused: optional F store
    const T* table;
    double* partials;
    T scale;
};

// Resolves the local element offset and class index of the cell `x + sgn*off` given local coords.
struct CellRef {
    int64_t lin;
    int cls;
};

__device__ __forceinline__ CellRef neighbour(const GenParams& p, const int64_t* xc /*local coords*/, const int* off,
                                             int sgn) {
    CellRef r;
    r.lin = 0;
    r.cls = 0;
#pragma unroll
    for (int a = 0; a < ODIL_B200_MAX_NDIM; ++a) {
        if (a >= p.ndim) break;
        const int s = sgn * off[a];
        const int64_t n = p.shape[a];
        int64_t il, ig;
        if (a == 0) {
            ig = p.z0 + xc[0] + s;
            if (ig < 0) ig += n;
            if (ig >= n) ig -= n;
            if (p.halo > 0) {
                il = xc[0] + s;  // physical halo plane
            } else {
                il = xc[0] + s;
                if (il < 0) il += p.n0;
                if (il >= p.n0) il -= p.n0;
            }
        } else {
            il = xc[a] + s;
            if (il < 0) il += n;
            if (il >= n) il -= n;
            ig = il;
        }
        r.lin += il * p.stride[a];
        r.cls += axis_class(ig, n, p.R[a]) * p.cstride[a];
    }
    return r;
}

template <typename T, int MODE>  // 0 forward, 1 adjoint, 2 fused
__global__ void __launch_bounds__(256) k_generic(GenParams p, BoxList boxes, GenIO<T> io) {
    __shared__ double red[32];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc2 = 0.0;
    if (gid < boxes.start[boxes.nbox]) {
        int b = 0;
        while (gid >= boxes.start[b + 1]) ++b;
        int64_t rem = gid - boxes.start[b];
        int64_t xc[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
        for (int a = p.ndim - 1; a >= 0; --a) {
            const int64_t s = boxes.sz[b][a];
            xc[a] = boxes.lo[b][a] + rem % s;
            rem /= s;
        }
        const int zero[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
        const CellRef self = neighbour(p, xc, zero, 0);
        if (MODE == 0) {
            T f = io.c ? io.c[self.lin] : T(0);
            for (int o = 0; o < p.noff; ++o) {
                const CellRef nb = neighbour(p, xc, p.off[o], +1);
                f += io.table[self.cls * p.noff + o] * io.U[nb.lin];
            }
            io.out[self.lin] = f;
        } else if (MODE == 1) {
            T g = T(0);
            for (int o = 0; o < p.noff; ++o) {
                const CellRef nb = neighbour(p, xc, p.off[o], -1);
                g += io.table[nb.cls * p.noff + o] * io.U[nb.lin];
            }
            g *= io.scale;
            if (io.c) g += io.c[self.lin];
            io.out[self.lin] = g;
        } else {
            // F at a cell y (given by local coords yc): sum_p table[cls(y)][p] * U[y + off_p] + c[y]
            auto eval_F = [&](const int64_t* yc, const CellRef& yref) -> T {
                T f = io.c ? io.c[yref.lin] : T(0);
                for (int q = 0; q < p.noff; ++q) {
                    const CellRef nb = neighbour(p, yc, p.off[q], +1);
                    f += io.table[yref.cls * p.noff + q] * io.U[nb.lin];
                }
                return f;
            };
            T g = T(0);
            T fself = T(0);
            bool have_self = false;
            for (int o = 0; o < p.noff; ++o) {
                // y = x - off_o, in local coords with the same wrapping rule as `neighbour`
                int64_t yc[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
                for (int a = 0; a < p.ndim; ++a) {
                    int64_t v = xc[a] - p.off[o][a];
                    if (a == 0) {
                        if (p.halo == 0) {
                            if (v < 0) v += p.n0;
                            if (v >= p.n0) v -= p.n0;
                        }
                    } else {
                        if (v < 0) v += p.shape[a];
                        if (v >= p.shape[a]) v -= p.shape[a];
                    }
                    yc[a] = v;
                }
                const CellRef yref = neighbour(p, yc, zero, 0);
                const T fy = eval_F(yc, yref);
                if (o == p.zero_off) {
                    fself = fy;
                    have_self = true;
                }
                g += io.table[yref.cls * p.noff + o] * fy;
            }
            if (!have_self) fself = eval_F(xc, self);
            io.out[self.lin] = g * io.scale;
            if (io.Fout) io.Fout[self.lin] = fself;
            bool count = true;
            if (p.count_mode == 1) {
                // interior class index = sum_a R[a]*cstride[a]
                int cint = 0;
                for (int a = 0; a < p.ndim; ++a) cint += p.R[a] * p.cstride[a];
                count = self.cls != cint;
            }
            if (count) acc2 = (double)fself * (double)fself;
        }
    }
    if (MODE == 2) {
        const double s = block_sum(acc2, red);
        if (threadIdx.x == 0) io.partials[blockIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled star kernel
// ------------------------------------------------------------------------------------------------
template <typename T>
struct StarParams {
    const T* U;
    const T* c;
    T* G;
    T* Fout;
    double* partials;
    int64_t n0, N0g, z0;
    int halo;
    int N1, N2;
    T wc, wzm, wzp, wym, wyp, wxm, wxp;
    T scale;
    int R0, R1, R2;
    int zchunk;
    int has_z;
};

template <typename T, int TY, int TX, int NT, bool VEC>
__global__ void __launch_bounds__(NT) k_star3d(StarParams<T> p) {
    constexpr int FH = TY + 2;
    constexpr int FW = TX + 2;
    constexpr int PITCH = TX + 8;
    constexpr int NC = FH * FW;
    constexpr int NCOL = (NC + NT - 1) / NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Fs = reinterpret_cast<T*>(smem_raw);  // [4][FH][PITCH]
    __shared__ double red[32];

    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TX;
    const int ty0 = blockIdx.y * TY;
    const int64_t zs = (int64_t)blockIdx.z * p.zchunk;
    const int64_t ze = min(zs + (int64_t)p.zchunk, p.n0);
    const int64_t plane = (int64_t)p.N1 * p.N2;

    // Per-column precomputation (columns are fixed while marching along axis 0).
    int offc[NCOL], oym[NCOL], oyp[NCOL], oxm[NCOL], oxp[NCOL], sidx[NCOL];
    unsigned valid = 0, counted = 0, inner = 0;
    T um[NCOL], uc[NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int i = tid + j * NT;
        offc[j] = oym[j] = oyp[j] = oxm[j] = oxp[j] = 0;
        sidx[j] = 0;
        um[j] = uc[j] = T(0);
        if (i < NC) {
            valid |= 1u << j;
            const int fy = i / FW, fx = i - fy * FW;
            const int y = ty0 - 1 + fy, x = tx0 - 1 + fx;
            const int yw = ((y % p.N1) + p.N1) % p.N1;
            const int xw = ((x % p.N2) + p.N2) % p.N2;
            const int ym = yw == 0 ? p.N1 - 1 : yw - 1;
            const int yp = yw == p.N1 - 1 ? 0 : yw + 1;
            const int xm = xw == 0 ? p.N2 - 1 : xw - 1;
            const int xp = xw == p.N2 - 1 ? 0 : xw + 1;
            offc[j] = yw * p.N2 + xw;
            oym[j] = ym * p.N2 + xw;
            oyp[j] = yp * p.N2 + xw;
            oxm[j] = yw * p.N2 + xm;
            oxp[j] = yw * p.N2 + xp;
            sidx[j] = fy * PITCH + fx + 3;
            const bool in_tile = fy >= 1 && fy <= TY && fx >= 1 && fx <= TX && y < p.N1 && x < p.N2;
            if (in_tile) inner |= 1u << j;
            if (in_tile && y >= p.R1 && y < p.N1 - p.R1 && x >= p.R2 && x < p.N2 - p.R2) counted |= 1u << j;
        }
    }

    const int n0i = (int)p.n0;
    auto zoff = [&](int64_t k64) -> int64_t {
        int k = (int)k64;
        if (p.halo == 0) {
            while (k < 0) k += n0i;
            while (k >= n0i) k -= n0i;
        }
        return (int64_t)k * plane;
    };

    const int64_t kf_begin = p.has_z ? zs - 1 : zs;
    const int64_t kf_end = p.has_z ? ze + 1 : ze;
    if (p.has_z) {
        const T* Um = p.U + zoff(kf_begin - 1);
        const T* Uc = p.U + zoff(kf_begin);
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (valid >> j & 1) {
                um[j] = __ldg(Um + offc[j]);
                uc[j] = __ldg(Uc + offc[j]);
            }
    } else {
        const T* Uc = p.U + zoff(kf_begin);
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
            if (valid >> j & 1) uc[j] = __ldg(Uc + offc[j]);
    }

    double acc2 = 0.0;
    for (int64_t kf = kf_begin; kf < kf_end; ++kf) {
        const int slot = (int)((kf - kf_begin) & 3);
        const int64_t zo = zoff(kf);
        const T* Uc = p.U + zo;
        const T* Up = p.U + zoff(kf + 1);
        const T* cp = p.c ? p.c + zo : nullptr;
        T* Fslot = Fs + slot * (FH * PITCH);
        const int64_t zg = p.z0 + kf;
        const bool zcount = kf >= zs && kf < ze && zg >= p.R0 && zg < p.N0g - p.R0;
        const bool zown = kf >= zs && kf < ze;
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            if (valid >> j & 1) {
                T up = T(0);
                if (p.has_z) up = __ldg(Up + offc[j]);
                T f = cp ? __ldg(cp + offc[j]) : T(0);
                f += p.wc * uc[j];
                f += p.wzm * um[j];
                f += p.wzp * up;
                f += p.wym * __ldg(Uc + oym[j]);
                f += p.wyp * __ldg(Uc + oyp[j]);
                f += p.wxm * __ldg(Uc + oxm[j]);
                f += p.wxp * __ldg(Uc + oxp[j]);
                Fslot[sidx[j]] = f;
                if (zcount && (counted >> j & 1)) acc += f * f;
                if (p.Fout && zown && (inner >> j & 1)) p.Fout[kf * plane + offc[j]] = f;
                um[j] = uc[j];
                uc[j] = up;
            }
        }
        acc2 += (double)acc;
        __syncthreads();
        const int64_t kg = p.has_z ? kf - 1 : kf;
        if (!p.has_z || kf >= zs + 1) {
            const T* Fc = Fs + (int)((kg - kf_begin) & 3) * (FH * PITCH);
            const T* Fm = p.has_z ? Fs + (int)((kg - 1 - kf_begin) & 3) * (FH * PITCH) : Fc;
            const T* Fp = p.has_z ? Fs + (int)((kg + 1 - kf_begin) & 3) * (FH * PITCH) : Fc;
            T* Gp = p.G + kg * plane;
            if (VEC) {
                constexpr int GX = TX / 4;
                for (int t = tid; t < TY * GX; t += NT) {
                    const int gy = t / GX, gx = (t - gy * GX) * 4;
                    const int y = ty0 + gy, x = tx0 + gx;
                    if (y < p.N1 && x < p.N2) {
                        const T* r0 = Fc + (gy + 1) * PITCH + gx + 4;
                        const Vec4<T> fc = *reinterpret_cast<const Vec4<T>*>(r0);
                        const T fl = r0[-1], fr = r0[4];
                        const Vec4<T> fym = *reinterpret_cast<const Vec4<T>*>(r0 - PITCH);
                        const Vec4<T> fyp = *reinterpret_cast<const Vec4<T>*>(r0 + PITCH);
                        Vec4<T> g;
                        g.x = p.wc * fc.x + p.wxm * fc.y + p.wxp * fl + p.wym * fyp.x + p.wyp * fym.x;
                        g.y = p.wc * fc.y + p.wxm * fc.z + p.wxp * fc.x + p.wym * fyp.y + p.wyp * fym.y;
                        g.z = p.wc * fc.z + p.wxm * fc.w + p.wxp * fc.y + p.wym * fyp.z + p.wyp * fym.z;
                        g.w = p.wc * fc.w + p.wxm * fr + p.wxp * fc.z + p.wym * fyp.w + p.wyp * fym.w;
                        if (p.has_z) {
                            const Vec4<T> fzm = *reinterpret_cast<const Vec4<T>*>(Fm + (gy + 1) * PITCH + gx + 4);
                            const Vec4<T> fzp = *reinterpret_cast<const Vec4<T>*>(Fp + (gy + 1) * PITCH + gx + 4);
                            g.x += p.wzm * fzp.x + p.wzp * fzm.x;
                            g.y += p.wzm * fzp.y + p.wzp * fzm.y;
                            g.z += p.wzm * fzp.z + p.wzp * fzm.z;
                            g.w += p.wzm * fzp.w + p.wzp * fzm.w;
                        }
                        g.x *= p.scale;
                        g.y *= p.scale;
                        g.z *= p.scale;
                        g.w *= p.scale;
                        *reinterpret_cast<Vec4<T>*>(Gp + (int64_t)y * p.N2 + x) = g;
                    }
                }
            } else {
                for (int t = tid; t < TY * TX; t += NT) {
                    const int gy = t / TX, gx = t - gy * TX;
                    const int y = ty0 + gy, x = tx0 + gx;
                    if (y < p.N1 && x < p.N2) {
                        const T* r0 = Fc + (gy + 1) * PITCH + gx + 4;
                        T g = p.wc * r0[0] + p.wxm * r0[1] + p.wxp * r0[-1] + p.wym * r0[PITCH] + p.wyp * r0[-PITCH];
                        if (p.has_z)
                            g += p.wzm * Fp[(gy + 1) * PITCH + gx + 4] + p.wzp * Fm[(gy + 1) * PITCH + gx + 4];
                        Gp[(int64_t)y * p.N2 + x] = g * p.scale;
                    }
                }
            }
        }
    }
    const double s = block_sum(acc2, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}


// ------------------------------------------------------------------------------------------------
// Star kernel v3: one thread per float4 COLUMN GROUP of the F region (tile + 1-cell ring), marching
// along axis 0 with U[k-1], U[k], U[k+1] and F[k-2], F[k-1], F[k] of its own column in registers.
// Per plane and thread: 2 x LDG.128 (next U plane, c), 2 x STS.128 (own U[k], own F[k-1]),
// one __syncthreads, 4 x LDS.128 + 4 x LDS.32 (in-plane neighbours of U and F), 1 x STG.128.
// Boundary rows (non-interior region classes) are handled in the same sweep by a per-cell table
// lookup on the few threads/planes that touch them -- no separate shell pass.
// Requires N2 % 4 == 0 (16-byte column groups).
// ------------------------------------------------------------------------------------------------
template <typename T>
struct StarV3Params {
    const T* U;
    const T* c;
    T* G;
    T* Fout;
    double* partials;
    const T* table;  // [C0*C1*C2][7] in star order: c, zm, zp, ym, yp, xm, xp
    int64_t n0, N0g, z0;
    int halo;
    int N1, N2;
    int R0, R1, R2;
    T w[7];
    T scale;
    int zchunk;
    int has_z;
};

__device__ __forceinline__ int wrapi(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}

template <typename T>
__device__ __forceinline__ Vec4<T> ldg4(const T* p) {
    return __ldg(reinterpret_cast<const Vec4<T>*>(p));
}
template <>
__device__ __forceinline__ Vec4<float> ldg4<float>(const float* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return Vec4<float>{v.x, v.y, v.z, v.w};
}
template <>
__device__ __forceinline__ Vec4<double> ldg4<double>(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return Vec4<double>{a.x, a.y, b.x, b.y};
}

// Interior rows: F for four consecutive cells of one row, and the adjoint gather of g.
template <typename T>
__device__ __forceinline__ Vec4<T> star_fwd(const Vec4<T>& cc, const Vec4<T>& uc, const Vec4<T>& um, const Vec4<T>& up,
                                            const Vec4<T>& uym, const Vec4<T>& uyp, T ul, T ur, const T* w) {
    Vec4<T> f;
    f.x = cc.x + w[0] * uc.x + w[1] * um.x + w[2] * up.x + w[3] * uym.x + w[4] * uyp.x + w[5] * ul + w[6] * uc.y;
    f.y = cc.y + w[0] * uc.y + w[1] * um.y + w[2] * up.y + w[3] * uym.y + w[4] * uyp.y + w[5] * uc.x + w[6] * uc.z;
    f.z = cc.z + w[0] * uc.z + w[1] * um.z + w[2] * up.z + w[3] * uym.z + w[4] * uyp.z + w[5] * uc.y + w[6] * uc.w;
    f.w = cc.w + w[0] * uc.w + w[1] * um.w + w[2] * up.w + w[3] * uym.w + w[4] * uyp.w + w[5] * uc.z + w[6] * ur;
    return f;
}

template <typename T>
__device__ __forceinline__ Vec4<T> star_adj(const Vec4<T>& fc, const Vec4<T>& fp, const Vec4<T>& fm, const Vec4<T>& fym,
                                            const Vec4<T>& fyp, T fl, T fr, const T* w) {
    Vec4<T> g;
    g.x = w[0] * fc.x + w[1] * fp.x + w[2] * fm.x + w[3] * fyp.x + w[4] * fym.x + w[5] * fc.y + w[6] * fl;
    g.y = w[0] * fc.y + w[1] * fp.y + w[2] * fm.y + w[3] * fyp.y + w[4] * fym.y + w[5] * fc.z + w[6] * fc.x;
    g.z = w[0] * fc.z + w[1] * fp.z + w[2] * fm.z + w[3] * fyp.z + w[4] * fym.z + w[5] * fc.w + w[6] * fc.y;
    g.w = w[0] * fc.w + w[1] * fp.w + w[2] * fm.w + w[3] * fyp.w + w[4] * fym.w + w[5] * fr + w[6] * fc.z;
    return g;
}

// Boundary rows (rare): recompute the flagged cells with their own coefficient row.  `cxp` packs the
// x-classes of cells x0-1 .. x0+4 (5 bits each).  Kept out of line so that the hot loop stays lean.
template <typename T>
__device__ __noinline__ void star_patch_fwd(Vec4<T>& f, unsigned mask, const T* __restrict__ tab, int rbase,
                                            unsigned cxp, Vec4<T> cc, Vec4<T> uc, Vec4<T> um, Vec4<T> up,
                                            Vec4<T> uym, Vec4<T> uyp, T ul, T ur) {
    const T ucv[6] = {ul, uc.x, uc.y, uc.z, uc.w, ur};
    const T umv[4] = {um.x, um.y, um.z, um.w}, upv[4] = {up.x, up.y, up.z, up.w};
    const T uymv[4] = {uym.x, uym.y, uym.z, uym.w}, uypv[4] = {uyp.x, uyp.y, uyp.z, uyp.w};
    const T ccv[4] = {cc.x, cc.y, cc.z, cc.w};
    T fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (mask >> i & 1) {
            const T* row = tab + (rbase + (int)((cxp >> (5 * (i + 1))) & 31u)) * 7;
            fv[i] = ccv[i] + row[0] * ucv[i + 1] + row[1] * umv[i] + row[2] * upv[i] + row[3] * uymv[i] +
                    row[4] * uypv[i] + row[5] * ucv[i] + row[6] * ucv[i + 2];
        }
    }
    f = Vec4<T>{fv[0], fv[1], fv[2], fv[3]};
}

template <typename T>
__device__ __noinline__ void star_patch_adj(Vec4<T>& g, unsigned mask, const T* __restrict__ tab, int C1, int C2,
                                            int czm, int cz0, int czp, int cym, int cy, int cyp, unsigned cxp,
                                            bool has_z, Vec4<T> fc, Vec4<T> fp, Vec4<T> fm, Vec4<T> fym, Vec4<T> fyp,
                                            T fl, T fr) {
    const T fcv[6] = {fl, fc.x, fc.y, fc.z, fc.w, fr};
    const T fmv[4] = {fm.x, fm.y, fm.z, fm.w}, fpv[4] = {fp.x, fp.y, fp.z, fp.w};
    const T fymv[4] = {fym.x, fym.y, fym.z, fym.w}, fypv[4] = {fyp.x, fyp.y, fyp.z, fyp.w};
    T gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (mask >> i & 1) {
            const int cxm = (cxp >> (5 * i)) & 31u, cx = (cxp >> (5 * (i + 1))) & 31u, cxq = (cxp >> (5 * (i + 2))) & 31u;
            T gi = tab[((cz0 * C1 + cy) * C2 + cx) * 7 + 0] * fcv[i + 1];
            if (has_z)
                gi += tab[((czp * C1 + cy) * C2 + cx) * 7 + 1] * fpv[i] + tab[((czm * C1 + cy) * C2 + cx) * 7 + 2] * fmv[i];
            gi += tab[((cz0 * C1 + cyp) * C2 + cx) * 7 + 3] * fypv[i] + tab[((cz0 * C1 + cym) * C2 + cx) * 7 + 4] * fymv[i];
            gi += tab[((cz0 * C1 + cy) * C2 + cxq) * 7 + 5] * fcv[i + 2] + tab[((cz0 * C1 + cy) * C2 + cxm) * 7 + 6] * fcv[i];
            gv[i] = gi;
        }
    }
    g = Vec4<T>{gv[0], gv[1], gv[2], gv[3]};
}

// Threads: (TX/4 + 2) column groups x (TY + 4) rows.  Row ry holds y = ty0 - 2 + ry.
//   rows 0 and TY+3      : loaders (only publish their U plane row: the y-neighbours of the F ring)
//   rows 1 .. TY+2       : compute F for their column group (tile + 1-cell ring)
//   rows 2 .. TY+1       : additionally compute g and the loss partial (the tile itself)
// Each thread keeps U[k-1], U[k], U[k+1] (+ prefetched U[k+2]) and F[k-2], F[k-1], F[k] of its column
// group in registers; in-plane neighbours go through double-buffered shared-memory planes.
template <typename T, int TY, int TX, bool HASZ>
__global__ void __launch_bounds__((TX / 4 + 2) * (TY + 4)) k_star_v3(StarV3Params<T> p) {
    constexpr int GXN = TX / 4 + 2;
    constexpr int NR = TY + 4;
    constexpr int PITCH = GXN * 4 + 8;
    constexpr int PLN = NR * PITCH;
    constexpr int kTabSmem = 512;  // table entries kept in shared memory (27 classes x 7 = 189 for r = 1)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Us = reinterpret_cast<T*>(smem_raw);  // [2][NR][PITCH]
    T* Fs = Us + 2 * PLN;                    // [2][NR][PITCH]
    T* tab_s = Fs + 2 * PLN;                 // [kTabSmem]
    __shared__ double red[32];

    const int tid = threadIdx.x;
    const int ry = tid / GXN, gx = tid - ry * GXN;
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    const int y = ty0 - 2 + ry, x0 = tx0 - 4 + 4 * gx;
    const int yw = wrapi(y, p.N1), x0w = wrapi(x0, p.N2);
    const int n0i = (int)p.n0, N0gi = (int)p.N0g, z0i = (int)p.z0;
    const int zs = blockIdx.z * p.zchunk;
    const int ze = min(zs + p.zchunk, n0i);
    const int64_t plane = (int64_t)p.N1 * p.N2;
    const int col = yw * p.N2 + x0w;
    const int soff = ry * PITCH + 4 + 4 * gx;
    const bool frow = ry >= 1 && ry <= TY + 2;
    const bool grow = ry >= 2 && ry <= TY + 1 && gx >= 1 && gx <= GXN - 2 && y < p.N1 && x0 < p.N2;

    const int C1 = 2 * p.R1 + 1, C2 = 2 * p.R2 + 1;
    auto cls1 = [](int i, int n, int r) -> int {
        if (i < r) return i;
        const int d = n - 1 - i;
        return d < r ? 2 * r - d : r;
    };
    const int ncls = (2 * p.R0 + 1) * C1 * C2;
    const bool tab_in_smem = ncls * 7 <= kTabSmem;
    if (tab_in_smem)
        for (int i = tid; i < ncls * 7; i += blockDim.x) tab_s[i] = p.table[i];
    const T* __restrict__ tab = tab_in_smem ? tab_s : p.table;

    // classes of the row and of the cells x0-1 .. x0+4 (packed 5 bits each); masks of non-interior cells
    const int cy = cls1(yw, p.N1, p.R1);
    const int cym = cls1(wrapi(yw - 1, p.N1), p.N1, p.R1), cyp = cls1(wrapi(yw + 1, p.N1), p.N1, p.R1);
    unsigned cxp = 0, fmask = 0, amask = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) cxp |= (unsigned)cls1(wrapi(x0w + i - 1, p.N2), p.N2, p.R2) << (5 * i);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = (cxp >> (5 * i)) & 31u, b = (cxp >> (5 * (i + 1))) & 31u, c = (cxp >> (5 * (i + 2))) & 31u;
        if (b != p.R2) fmask |= 1u << i;
        if (a != p.R2 || b != p.R2 || c != p.R2) amask |= 1u << i;
    }
    if (cy != p.R1) fmask = 0xFu;
    if (cy != p.R1 || cym != p.R1 || cyp != p.R1) amask = 0xFu;
    if (!frow) fmask = 0;
    if (!grow) amask = 0;

    auto zwrap = [&](int k) -> int {
        if (p.halo == 0) {
            while (k < 0) k += n0i;
            while (k >= n0i) k -= n0i;
        }
        return k;
    };
    auto zcls = [&](int k) -> int {
        int zg = z0i + k;
        while (zg < 0) zg += N0gi;
        while (zg >= N0gi) zg -= N0gi;
        return cls1(zg, N0gi, p.R0);
    };

    const Vec4<T> zero4{T(0), T(0), T(0), T(0)};
    Vec4<T> um = zero4, uc, up = zero4, un = zero4, cc = zero4, cn = zero4, fm = zero4, fc = zero4, fp = zero4;
    const bool zvar = HASZ || p.R0 > 0;
    const int kf0 = HASZ ? zs - 1 : zs;
    const T* __restrict__ Ucol = p.U + col;
    const T* __restrict__ Ccol = (p.c && frow) ? p.c + col : nullptr;
    if (HASZ) {
        um = ldg4<T>(Ucol + (int64_t)zwrap(kf0 - 1) * plane);
        up = ldg4<T>(Ucol + (int64_t)zwrap(kf0 + 1) * plane);
    }
    uc = ldg4<T>(Ucol + (int64_t)zwrap(kf0) * plane);
    if (Ccol) cc = ldg4<T>(Ccol + (int64_t)zwrap(kf0) * plane);
    T w[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = p.w[i];
    int czm = zvar ? zcls(kf0 - 2) : 0, cz0 = zvar ? zcls(kf0 - 1) : 0, czp = zvar ? zcls(kf0) : 0;
    // running (wrapped) plane counters for the prefetches: k1 = plane kf+1, k2 = plane kf+2, zg1 = global kf+1
    const int wrapn = p.halo == 0 ? n0i : 0x7fffffff;
    int k1 = zwrap(kf0 + 1), k2 = zwrap(kf0 + 2), zg1 = z0i + kf0 + 1;
    while (zg1 < 0) zg1 += N0gi;
    while (zg1 >= N0gi) zg1 -= N0gi;
    T* Gcol = p.G + col;
    T* Fcol = p.Fout ? p.Fout + col : nullptr;

    T accf = T(0);
    double acc2 = 0.0;
    int pboff = 0;
#pragma unroll 2
    for (int kf = kf0; kf <= ze; ++kf) {
        T* Ub = Us + pboff + soff;
        T* Fb = Fs + pboff + soff;
        pboff = PLN - pboff;
        // (a) prefetch the next plane's inputs (consumed one iteration later)
        if (kf < ze) {
            un = ldg4<T>(Ucol + (int64_t)(HASZ ? k2 : k1) * plane);
            if (Ccol) cn = ldg4<T>(Ccol + (int64_t)k1 * plane);
        }
        k1 = k1 + 1 == wrapn ? 0 : k1 + 1;
        k2 = k2 + 1 == wrapn ? 0 : k2 + 1;
        // (b) publish own U[kf] and F[kf-1]
        *reinterpret_cast<Vec4<T>*>(Ub) = uc;
        *reinterpret_cast<Vec4<T>*>(Fb) = fc;
        __syncthreads();
        // (d) F[kf]
        fp = zero4;
        if (frow && (HASZ || kf < ze)) {
            const Vec4<T> uym = *reinterpret_cast<const Vec4<T>*>(Ub - PITCH);
            const Vec4<T> uyp = *reinterpret_cast<const Vec4<T>*>(Ub + PITCH);
            const T ul = Ub[-1], ur = Ub[4];
            fp = star_fwd<T>(cc, uc, um, up, uym, uyp, ul, ur, w);
            const unsigned fmk = (czp != p.R0) ? 0xFu : fmask;  // czp == class of plane kf here
            if (fmk) star_patch_fwd<T>(fp, fmk, tab, (czp * C1 + cy) * C2, cxp, cc, uc, um, up, uym, uyp, ul, ur);
            if (grow && kf >= zs && kf < ze) {
                accf += fp.x * fp.x + fp.y * fp.y + fp.z * fp.z + fp.w * fp.w;
                if (Fcol) *reinterpret_cast<Vec4<T>*>(Fcol + (int64_t)kf * plane) = fp;
            }
        }
        // (e) g[kf-1] from F[kf-2], F[kf-1], F[kf] (own column) and the in-plane neighbours of F[kf-1]
        const int kg = kf - 1;
        if (grow && kg >= zs && kg < ze) {
            const Vec4<T> fym = *reinterpret_cast<const Vec4<T>*>(Fb - PITCH);
            const Vec4<T> fyp = *reinterpret_cast<const Vec4<T>*>(Fb + PITCH);
            const T fl = Fb[-1], fr = Fb[4];
            Vec4<T> g = star_adj<T>(fc, fp, fm, fym, fyp, fl, fr, w);
            // plane classes here: czm = class(kg-1), cz0 = class(kg), czp = class(kg+1)
            const bool zslow = cz0 != p.R0 || (HASZ && (czm != p.R0 || czp != p.R0));
            const unsigned amk = zslow ? 0xFu : amask;
            if (amk)
                star_patch_adj<T>(g, amk, tab, C1, C2, czm, cz0, czp, cym, cy, cyp, cxp, HASZ, fc, fp, fm, fym, fyp, fl,
                                  fr);
            g.x *= p.scale;
            g.y *= p.scale;
            g.z *= p.scale;
            g.w *= p.scale;
            *reinterpret_cast<Vec4<T>*>(Gcol + (int64_t)kg * plane) = g;
        }
        // (f) rotate
        fm = fc;
        fc = fp;
        if (HASZ) {
            um = uc;
            uc = up;
            up = un;
        } else {
            uc = un;
        }
        cc = cn;
        if (zvar) {
            czm = cz0;
            cz0 = czp;
            czp = cls1(zg1, N0gi, p.R0);
            zg1 = zg1 + 1 == N0gi ? 0 : zg1 + 1;
        }
        if (((kf - kf0) & 7) == 7) {  // fold the partial into the fp64 accumulator every 8 planes
            acc2 += (double)accf;
            accf = T(0);
        }
    }
    acc2 += (double)accf;
    const double sum = block_sum(acc2, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = sum;
}


// ------------------------------------------------------------------------------------------------
// Star kernel, TMA-fed (k_star_tma).  Same sweep as k_star_v3 but the U and c planes (tile + halo) are
// brought into shared-memory rings by the Tensor Memory Accelerator (cp.async.bulk.tensor.3d, zero fill
// outside the array) two planes ahead, signalled through mbarriers; F lives in a third ring.  Threads do
// no global loads and keep no planes in registers: per plane they read their column group and its
// neighbours from shared memory, write F, and (one plane later) gather g and store it with one STG.128.
// Requires a "wrap-free" plan: no coefficient multiplies a neighbour across a periodic boundary (true
// for every Dirichlet/Neumann-by-extrapolation operator; periodic problems use k_star_v3).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // a lost TMA must fail loudly, never hang the device
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

template <typename T>
struct StarTmaParams {
    T* G;
    T* Fout;
    double* partials;
    const T* table;
    int n0, N0g, z0, halo;
    int N1, N2;
    int R0, R1, R2;
    T w[7];
    T scale;
    int zchunk;
    int has_c;
};

template <typename T, int TY, int TX>
__global__ void __launch_bounds__((TX / 4 + 2) * (TY + 2), ((TX / 4 + 2) * (TY + 2) <= 352 ? 2 : 1))
    k_star_tma(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmC, StarTmaParams<T> p) {
    constexpr int GXN = TX / 4 + 2;
    constexpr int NR = TY + 4;         // rows of a staged plane: tile + 2-cell halo (U of the F ring's neighbours)
    constexpr int BX = GXN * 4;        // dense row of the TMA box
    constexpr int PLN = ((NR * BX + 31) / 32) * 32;  // elements per plane slot, 128-byte multiple (TMA destination)
    constexpr int NSU = 4, NSC = 2, NSF = 4;
    constexpr int kTabSmem = 512;
    constexpr uint32_t kBytes = NR * BX * sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [128 B pad][U ring][C ring][F ring][table][pad]
    T* Us = reinterpret_cast<T*>(smem_raw + 128);
    T* Cs = Us + NSU * PLN;
    T* Fs = Cs + NSC * PLN;
    T* tab_s = Fs + NSF * PLN;
    __shared__ __align__(8) uint64_t bar_u[NSU];
    __shared__ __align__(8) uint64_t bar_c[NSC];
    __shared__ double red[32];

    // One thread per float4 column group of the F region: rows ry = 1 .. TY+2 of the staged plane.
    const int tid = threadIdx.x;
    const int ry = tid / GXN + 1, gx = tid - (ry - 1) * GXN;
    const int lane = tid & 31;
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    const int y = ty0 - 2 + ry, x0 = tx0 - 4 + 4 * gx;
    const int zs = blockIdx.z * p.zchunk;
    const int ze = min(zs + p.zchunk, p.n0);
    const int64_t plane = (int64_t)p.N1 * p.N2;
    const int soff = ry * BX + 4 * gx;
    const bool in_dom = y >= 0 && y < p.N1 && x0 >= 0 && x0 < p.N2;
    const bool grow = ry >= 2 && ry <= TY + 1 && gx >= 1 && gx <= GXN - 2 && in_dom;
    const int col = y * p.N2 + x0;  // only used when grow
    // x-neighbours come from the adjacent lanes' registers; lanes at a warp or row edge read shared memory
    const bool shl_ok = lane > 0 && gx > 0, shr_ok = lane < 31 && gx < GXN - 1;

    const int C1 = 2 * p.R1 + 1, C2 = 2 * p.R2 + 1;
    auto cls1 = [](int i, int n, int r) -> int {
        if (i < r) return i;
        const int d = n - 1 - i;
        return d < r ? 2 * r - d : r;
    };
    const int ncls = (2 * p.R0 + 1) * C1 * C2;
    const bool tab_in_smem = ncls * 7 <= kTabSmem;
    if (tab_in_smem)
        for (int i = tid; i < ncls * 7; i += blockDim.x) tab_s[i] = p.table[i];
    const T* __restrict__ tab = tab_in_smem ? tab_s : p.table;
    const int yw = wrapi(y, p.N1), x0w = wrapi(x0, p.N2);
    const int cy = cls1(yw, p.N1, p.R1);
    const int cym = cls1(wrapi(yw - 1, p.N1), p.N1, p.R1), cyp = cls1(wrapi(yw + 1, p.N1), p.N1, p.R1);
    unsigned cxp = 0, fmask = 0, amask = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) cxp |= (unsigned)cls1(wrapi(x0w + i - 1, p.N2), p.N2, p.R2) << (5 * i);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = (cxp >> (5 * i)) & 31u, b = (cxp >> (5 * (i + 1))) & 31u, c = (cxp >> (5 * (i + 2))) & 31u;
        if (b != p.R2) fmask |= 1u << i;
        if (a != p.R2 || b != p.R2 || c != p.R2) amask |= 1u << i;
    }
    if (cy != p.R1) fmask = 0xFu;
    if (cy != p.R1 || cym != p.R1 || cyp != p.R1) amask = 0xFu;
    if (!grow) amask = 0;
    auto zcls = [&](int k) -> int {
        int zg = p.z0 + k;
        while (zg < 0) zg += p.N0g;
        while (zg >= p.N0g) zg -= p.N0g;
        return cls1(zg, p.N0g, p.R0);
    };

    const int kf0 = zs - 1;
    const int niter = ze - kf0 + 1;  // planes kf0 .. ze
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSU; ++i) mbar_init(&bar_u[i], 1);
#pragma unroll
        for (int i = 0; i < NSC; ++i) mbar_init(&bar_c[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // plane q (relative: q = plane - (kf0 - 1)) lives in U slot q & 3; c plane qc = plane - kf0 in slot qc & 1
    auto issue_u = [&](int q) {
        mbar_expect_tx(&bar_u[q & 3], kBytes);
        tma_load_3d(Us + (q & 3) * PLN, &tmU, &bar_u[q & 3], tx0 - 4, ty0 - 2, kf0 - 1 + q + p.halo);
    };
    auto issue_c = [&](int qc) {
        mbar_expect_tx(&bar_c[qc & 1], kBytes);
        tma_load_3d(Cs + (qc & 1) * PLN, &tmC, &bar_c[qc & 1], tx0 - 4, ty0 - 2, kf0 + qc + p.halo);
    };
    if (tid == 0) {
        issue_u(0);
        issue_u(1);
        issue_u(2);
        if (niter > 1) issue_u(3);
        if (p.has_c) {
            issue_c(0);
            if (niter > 1) issue_c(1);
        }
    }
    T w[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = p.w[i];
    int czm = zcls(kf0 - 2), cz0 = zcls(kf0 - 1), czp = zcls(kf0);
    int zg1 = p.z0 + kf0 + 1;
    while (zg1 < 0) zg1 += p.N0g;
    while (zg1 >= p.N0g) zg1 -= p.N0g;
    T* Gcol = p.G + col;
    T* Fcol = p.Fout ? p.Fout + col : nullptr;
    const Vec4<T> zero4{T(0), T(0), T(0), T(0)};

    // own column in registers: U[kf-1], U[kf] (U[kf+1] is read when its plane lands), F[kf-2], F[kf-1]
    mbar_wait(&bar_u[0], 0);
    mbar_wait(&bar_u[1], 0);
    Vec4<T> um = *reinterpret_cast<const Vec4<T>*>(Us + 0 * PLN + soff);
    Vec4<T> uc = *reinterpret_cast<const Vec4<T>*>(Us + 1 * PLN + soff);
    Vec4<T> fm = zero4, fc = zero4;

    T accf = T(0);
    double acc2 = 0.0;
#pragma unroll 4
    for (int it = 0; it < niter; ++it) {
        const int kf = kf0 + it;
        // plane kf lives in U slot (it+1)&3, plane kf+1 in slot (it+2)&3
        const T* Uc = Us + ((it + 1) & 3) * PLN + soff;
        T* Fw = Fs + (it & 3) * PLN + soff;
        // plane kf+1 (q = it+2): its use count of the slot is q >> 2
        mbar_wait(&bar_u[(it + 2) & 3], ((it + 2) >> 2) & 1);
        const Vec4<T> up = *reinterpret_cast<const Vec4<T>*>(Us + ((it + 2) & 3) * PLN + soff);
        Vec4<T> cc = zero4;
        if (p.has_c) {
            mbar_wait(&bar_c[it & 1], (it >> 1) & 1);
            cc = *reinterpret_cast<const Vec4<T>*>(Cs + (it & 1) * PLN + soff);
        }
        const Vec4<T> uym = *reinterpret_cast<const Vec4<T>*>(Uc - BX);
        const Vec4<T> uyp = *reinterpret_cast<const Vec4<T>*>(Uc + BX);
        T ul = __shfl_up_sync(0xffffffffu, uc.w, 1), ur = __shfl_down_sync(0xffffffffu, uc.x, 1);
        if (!shl_ok) ul = Uc[-1];
        if (!shr_ok) ur = Uc[4];
        Vec4<T> fp = star_fwd<T>(cc, uc, um, up, uym, uyp, ul, ur, w);
        const unsigned fmk = (czp != p.R0) ? 0xFu : fmask;  // czp == class of plane kf here
        if (fmk) star_patch_fwd<T>(fp, fmk, tab, (czp * C1 + cy) * C2, cxp, cc, uc, um, up, uym, uyp, ul, ur);
        *reinterpret_cast<Vec4<T>*>(Fw) = fp;
        if (grow && kf >= zs && kf < ze) {
            accf += fp.x * fp.x + fp.y * fp.y + fp.z * fp.z + fp.w * fp.w;
            if (Fcol) *reinterpret_cast<Vec4<T>*>(Fcol + (int64_t)kf * plane) = fp;
        }
        // x-neighbours of F[kf-1] (own registers of the adjacent lanes), before any divergence
        T fl = __shfl_up_sync(0xffffffffu, fc.w, 1), fr = __shfl_down_sync(0xffffffffu, fc.x, 1);
        __syncthreads();
        // refill the slots every thread has finished reading: U plane kf's... (kf-1 is only in registers now,
        // its slot was released one iteration ago; plane kf is still needed next iteration as y-neighbour? no:
        // next iteration reads planes kf+1 (neighbours) and kf+2 (own) -> slot of plane kf is free)
        if (tid == 0) {
            if (it + 4 <= niter + 1) issue_u(it + 4);  // into slot it & 3 (plane kf-1: released)
            if (p.has_c && it + 2 < niter) issue_c(it + 2);
        }
        // g[kf-1] from F[kf-2], F[kf-1], F[kf] (registers) and the in-plane neighbours of F[kf-1]
        const int kg = kf - 1;
        if (grow && kg >= zs && kg < ze) {
            const T* Fc = Fs + ((it + 3) & 3) * PLN + soff;  // plane kf-1
            const Vec4<T> fym = *reinterpret_cast<const Vec4<T>*>(Fc - BX);
            const Vec4<T> fyp = *reinterpret_cast<const Vec4<T>*>(Fc + BX);
            if (!shl_ok) fl = Fc[-1];
            if (!shr_ok) fr = Fc[4];
            Vec4<T> g = star_adj<T>(fc, fp, fm, fym, fyp, fl, fr, w);
            const bool zslow = cz0 != p.R0 || czm != p.R0 || czp != p.R0;
            const unsigned amk = zslow ? 0xFu : amask;
            if (amk)
                star_patch_adj<T>(g, amk, tab, C1, C2, czm, cz0, czp, cym, cy, cyp, cxp, true, fc, fp, fm, fym, fyp, fl,
                                  fr);
            g.x *= p.scale;
            g.y *= p.scale;
            g.z *= p.scale;
            g.w *= p.scale;
            *reinterpret_cast<Vec4<T>*>(Gcol + (int64_t)kg * plane) = g;
        }
        um = uc;
        uc = up;
        fm = fc;
        fc = fp;
        czm = cz0;
        cz0 = czp;
        czp = cls1(zg1, p.N0g, p.R0);
        zg1 = zg1 + 1 == p.N0g ? 0 : zg1 + 1;
        if ((it & 7) == 7) {
            acc2 += (double)accf;
            accf = T(0);
        }
    }
    acc2 += (double)accf;
    const double sum = block_sum(acc2, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = sum;
}

}  // namespace odil
#include "star7.cuh"
#include "star8.cuh"
#include "tile2d.cuh"
#include "tile3d.cuh"
namespace odil {

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static void fill_gen_params(const odil_b200_plan* plan, const odil_b200_slab* slab, GenParams& p) {
    memset(&p, 0, sizeof(p));
    p.ndim = plan->ndim;
    p.noff = plan->noff;
    int64_t st = 1;
    for (int a = plan->ndim - 1; a >= 0; --a) {
        p.shape[a] = plan->shape[a];
        p.stride[a] = st;
        st *= plan->shape[a];
    }
    int cs = 1;
    for (int a = plan->ndim - 1; a >= 0; --a) {
        p.R[a] = plan->R[a];
        p.cstride[a] = cs;
        cs *= 2 * plan->R[a] + 1;
    }
    p.n0 = slab->n0;
    p.z0 = slab->z0;
    p.halo = slab->halo;
    p.zero_off = -1;
    for (int o = 0; o < plan->noff; ++o) {
        bool z = true;
        for (int a = 0; a < plan->ndim; ++a) {
            p.off[o][a] = plan->off[o][a];
            z = z && plan->off[o][a] == 0;
        }
        if (z) p.zero_off = o;
    }
}

static void whole_box(const odil_b200_plan* plan, const odil_b200_slab* slab, BoxList& b) {
    memset(&b, 0, sizeof(b));
    b.nbox = 1;
    int64_t cnt = 1;
    for (int a = 0; a < plan->ndim; ++a) {
        b.lo[0][a] = 0;
        b.sz[0][a] = a == 0 ? slab->n0 : plan->shape[a];
        cnt *= b.sz[0][a];
    }
    b.start[0] = 0;
    b.start[1] = cnt;
}

// Boxes covering every owned cell within `thick[a]` of a domain face along some axis, disjoint.
static void shell_boxes(const odil_b200_plan* plan, const odil_b200_slab* slab, const int* thick, BoxList& b) {
    memset(&b, 0, sizeof(b));
    const int nd = plan->ndim;
    // global [lo, hi) ranges per axis: face-low, face-high, interior
    int64_t flo[ODIL_B200_MAX_NDIM][2], fhi[ODIL_B200_MAX_NDIM][2], inner[ODIL_B200_MAX_NDIM][2];
    bool has_hi[ODIL_B200_MAX_NDIM];
    for (int a = 0; a < nd; ++a) {
        const int64_t n = plan->shape[a];
        const int64_t t = thick[a];
        if (t <= 0) {
            flo[a][0] = flo[a][1] = 0;
            fhi[a][0] = fhi[a][1] = 0;
            has_hi[a] = false;
            inner[a][0] = 0;
            inner[a][1] = n;
        } else if (2 * t >= n) {
            flo[a][0] = 0;
            flo[a][1] = n;
            has_hi[a] = false;
            fhi[a][0] = fhi[a][1] = 0;
            inner[a][0] = inner[a][1] = 0;
        } else {
            flo[a][0] = 0;
            flo[a][1] = t;
            fhi[a][0] = n - t;
            fhi[a][1] = n;
            has_hi[a] = true;
            inner[a][0] = t;
            inner[a][1] = n - t;
        }
    }
    auto clip0 = [&](int64_t& lo, int64_t& hi) {  // global -> local along axis 0
        lo = std::max(lo, slab->z0) - slab->z0;
        hi = std::min(hi, slab->z0 + slab->n0) - slab->z0;
    };
    int nb = 0;
    int64_t total = 0;
    b.start[0] = 0;
    for (int a = 0; a < nd; ++a) {
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && !has_hi[a]) continue;
            int64_t lo[ODIL_B200_MAX_NDIM], hi[ODIL_B200_MAX_NDIM];
            bool empty = false;
            for (int c = 0; c < nd; ++c) {
                if (c == a) {
                    lo[c] = side ? fhi[a][0] : flo[a][0];
                    hi[c] = side ? fhi[a][1] : flo[a][1];
                } else if (c < a) {
                    lo[c] = inner[c][0];
                    hi[c] = inner[c][1];
                } else {
                    lo[c] = 0;
                    hi[c] = plan->shape[c];
                }
                if (c == 0) clip0(lo[c], hi[c]);
                if (hi[c] <= lo[c]) empty = true;
            }
            if (empty) continue;
            int64_t cnt = 1;
            for (int c = 0; c < nd; ++c) {
                b.lo[nb][c] = lo[c];
                b.sz[nb][c] = hi[c] - lo[c];
                cnt *= hi[c] - lo[c];
            }
            total += cnt;
            ++nb;
            b.start[nb] = total;
        }
    }
    b.nbox = nb;
}

template <typename T, int MODE>
static int launch_generic(const odil_b200_plan* plan, const GenParams& p, const BoxList& boxes, GenIO<T> io,
                          cudaStream_t st, int* nblocks_out) {
    const int64_t total = boxes.start[boxes.nbox];
    if (nblocks_out) *nblocks_out = 0;
    if (total == 0) return 0;
    const int64_t nb = (total + 255) / 256;
    ODIL_REQUIRE(nb < (1ll << 31), "generic stencil grid too large");
    if (MODE == 2) ODIL_REQUIRE(nb <= kPartialCapacity, "generic fused grid exceeds the partials workspace");
    k_generic<T, MODE><<<(unsigned)nb, 256, 0, st>>>(p, boxes, io);
    ODIL_LAUNCHED();
    if (nblocks_out) *nblocks_out = (int)nb;
    return 0;
}

template <typename T, int TY, int TX, int NT>
static int launch_star_cfg(const StarParams<T>& sp, dim3 grid, bool vec, cudaStream_t st) {
    const size_t smem = (size_t)4 * (TY + 2) * (TX + 8) * sizeof(T);
    static bool attr_set[2] = {false, false};  // per template instantiation
    if (!attr_set[vec ? 1 : 0]) {
        if (vec)
            ODIL_CUDA(cudaFuncSetAttribute(k_star3d<T, TY, TX, NT, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            ODIL_CUDA(cudaFuncSetAttribute(k_star3d<T, TY, TX, NT, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[vec ? 1 : 0] = true;
    }
    if (vec)
        k_star3d<T, TY, TX, NT, true><<<grid, NT, smem, st>>>(sp);
    else
        k_star3d<T, TY, TX, NT, false><<<grid, NT, smem, st>>>(sp);
    ODIL_LAUNCHED();
    return 0;
}

template <typename T, int TY, int TX>
static int launch_star_v3(const StarV3Params<T>& sp, dim3 grid, cudaStream_t st) {
    constexpr int NT = (TX / 4 + 2) * (TY + 4);
    const size_t smem = ((size_t)4 * (TY + 4) * ((TX / 4 + 2) * 4 + 8) + 512) * sizeof(T);
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star_v3<T, TY, TX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        ODIL_CUDA(cudaFuncSetAttribute(k_star_v3<T, TY, TX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        attr_set = true;
    }
    if (sp.has_z)
        k_star_v3<T, TY, TX, true><<<grid, NT, smem, st>>>(sp);
    else
        k_star_v3<T, TY, TX, false><<<grid, NT, smem, st>>>(sp);
    ODIL_LAUNCHED();
    return 0;
}

static void star_v3_tile(int variant, int& TY, int& TX) {
    switch (variant) {
        case 1: TY = 8; TX = 128; break;    // 408 threads
        case 2: TY = 12; TX = 128; break;   // 544 threads
        case 3: TY = 26; TX = 128; break;   // 1020 threads
        default: TY = 16; TX = 128; break;  // 680 threads
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}

// 3-D tensor map over a C-order (nplanes, N1, N2) array with a (1, BY, BX) box; zero fill outside.
template <typename T>
static int make_plane_map(CUtensorMap* map, const T* base, int64_t nplanes, int N1, int N2, int BY, int BX) {
    PFN_encodeTiled enc = get_encode_tiled();
    ODIL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t gdim[3] = {(cuuint64_t)N2, (cuuint64_t)N1, (cuuint64_t)nplanes};
    const cuuint64_t gstr[2] = {(cuuint64_t)N2 * sizeof(T), (cuuint64_t)N1 * N2 * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)BX, (cuuint32_t)BY, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    const CUresult r = enc(map, dt, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ODIL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

template <typename T, int TY, int TX>
static int launch_star_tma(const CUtensorMap& tmU, const CUtensorMap& tmC, const StarTmaParams<T>& sp, dim3 grid,
                           cudaStream_t st) {
    constexpr int NT = (TX / 4 + 2) * (TY + 2);
    constexpr int PLN = (((TY + 4) * (TX + 8) + 31) / 32) * 32;
    const size_t smem = 128 + (size_t)(4 + 2 + 4) * PLN * sizeof(T) + 512 * sizeof(T) + 128;
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star_tma<T, TY, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    k_star_tma<T, TY, TX><<<grid, NT, smem, st>>>(tmU, tmC, sp);
    ODIL_LAUNCHED();
    return 0;
}

template <typename T, int VW, int TY, bool XU>
static int launch_star7(const odil_b200_plan* plan, const odil_b200_slab* slab, const T* U, const T* c, T scale, T* G,
                        T* Fout, int* nparts, cudaStream_t st) {
    using Cfg = Star7Cfg<T, VW, TY>;
    Star7Params<T> sp;
    sp.G = G;
    sp.Fout = Fout;
    sp.partials = plan->partials;
    sp.table = (const T*)plan->star_table;
    if (plan->ndim == 3) {
        sp.n0 = (int)slab->n0;
        sp.N0g = (int)plan->shape[0];
        sp.z0 = (int)slab->z0;
        sp.halo = slab->halo;
        sp.N1 = (int)plan->shape[1];
        sp.N2 = (int)plan->shape[2];
        sp.R0 = plan->R[0];
        sp.R1 = plan->R[1];
        sp.R2 = plan->R[2];
    } else {
        sp.n0 = 1;
        sp.N0g = 1;
        sp.z0 = 0;
        sp.halo = 0;
        sp.N1 = (int)plan->shape[0];
        sp.N2 = (int)plan->shape[1];
        sp.R0 = 0;
        sp.R1 = plan->R[0];
        sp.R2 = plan->R[1];
    }
    sp.scale = scale;
    sp.has_c = c != nullptr;
    const int gx = (sp.N2 + Cfg::TX - 1) / Cfg::TX, gy = (sp.N1 + TY - 1) / TY;
    int zchunk = plan->zchunk;
    if (zchunk <= 0) {
        // fill the 148 x MINB CTA slots in whole waves while keeping the 2-plane lead-in of a chunk small:
        // cost(gz) ~ waves(gz) * (planes per chunk + lead-in)
        const int64_t slots = 148 * (XU ? Cfg::MINB_XU : Cfg::MINB), layer = (int64_t)gx * gy;
        int64_t best = -1;
        for (int gz = 1; gz <= std::max(1, sp.n0 / 8); ++gz) {
            const int zc = (sp.n0 + gz - 1) / gz;
            const int gzr = (sp.n0 + zc - 1) / zc;
            const int64_t cost = ((layer * gzr + slots - 1) / slots) * (zc + 5);
            if (best < 0 || cost < best) {
                best = cost;
                zchunk = zc;
            }
        }
    }
    if (zchunk > sp.n0) zchunk = sp.n0;
    if (zchunk < 1) zchunk = 1;
    sp.zchunk = zchunk;
    const int gz = (sp.n0 + zchunk - 1) / zchunk;
    ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity, "star grid exceeds the partials workspace");
    const int64_t planes_total = (int64_t)sp.n0 + 2 * sp.halo;
    const int64_t plane_elems = (int64_t)sp.N1 * sp.N2;
    CUtensorMap tmU, tmC;
    if (int rc = make_plane_map<T>(&tmU, U - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RU, Cfg::BX)) return rc;
    if (int rc = make_plane_map<T>(&tmC, (c ? c : U) - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RC, Cfg::BX))
        return rc;
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star7<T, VW, TY, XU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    k_star7<T, VW, TY, XU><<<dim3(gx, gy, gz), Cfg::NT, Cfg::SMEM, st>>>(tmU, tmC, sp);
    ODIL_LAUNCHED();
    *nparts = gx * gy * gz;
    return 0;
}

// Work list of k_star8: tiles of up to NR rows x TX columns times z-chunks, chosen so that the CTAs fill the
// SMs in whole waves with equal work.  cost ~ waves * (planes per chunk + lead-in) * (rows per CTA + ring rows).
static void build_work8(int NR, int TX, int slots, int n0, int N1, int N2, int zchunk, std::vector<S8Work>& out) {
    const int gx = (N2 + TX - 1) / TX;
    int64_t best = -1;
    int best_rmax = NR, best_zc = n0;
    for (int rmax = 1; rmax <= NR; ++rmax) {
        const int nblk = (N1 + rmax - 1) / rmax;
        const int gz_lo = zchunk > 0 ? std::max(1, (n0 + zchunk - 1) / zchunk) : 1;
        const int gz_hi = zchunk > 0 ? gz_lo : std::max(1, n0 / 8);
        for (int gz = gz_lo; gz <= gz_hi; ++gz) {
            const int zc = (n0 + gz - 1) / gz;
            const int gzr = (n0 + zc - 1) / zc;
            const int64_t ncta = (int64_t)gx * nblk * gzr;
            if (ncta > kPartialCapacity) continue;
            const int64_t cost = ((ncta + slots - 1) / slots) * (zc + 6) * (rmax + 3);
            if (best < 0 || cost < best) {
                best = cost;
                best_rmax = rmax;
                best_zc = zc;
            }
        }
    }
    const int nblk = (N1 + best_rmax - 1) / best_rmax;
    out.clear();
    for (int zs = 0; zs < n0; zs += best_zc)
        for (int b = 0; b < nblk; ++b) {
            const int y0 = (int)((int64_t)b * N1 / nblk), y1 = (int)((int64_t)(b + 1) * N1 / nblk);
            if (y1 <= y0) continue;
            for (int bx = 0; bx < gx; ++bx) out.push_back(S8Work{bx * TX, y0, y1 - y0, zs, std::min(n0, zs + best_zc), 0, 0, 0});
        }
}

template <typename T, int VW, int NR, bool XU>
static int launch_star8(const odil_b200_plan* plan, const odil_b200_slab* slab, const T* U, const T* c, T scale, T* G,
                        T* Fout, int* nparts, cudaStream_t st) {
    using Cfg = Star8Cfg<T, VW, NR>;
    Star8Params<T> sp;
    sp.G = G;
    sp.Fout = Fout;
    sp.partials = plan->partials;
    sp.table = (const T*)plan->star_table;
    if (plan->ndim == 3) {
        sp.n0 = (int)slab->n0;
        sp.N0g = (int)plan->shape[0];
        sp.z0 = (int)slab->z0;
        sp.halo = slab->halo;
        sp.N1 = (int)plan->shape[1];
        sp.N2 = (int)plan->shape[2];
        sp.R0 = plan->R[0];
        sp.R1 = plan->R[1];
        sp.R2 = plan->R[2];
    } else {
        sp.n0 = 1;
        sp.N0g = 1;
        sp.z0 = 0;
        sp.halo = 0;
        sp.N1 = (int)plan->shape[0];
        sp.N2 = (int)plan->shape[1];
        sp.R0 = 0;
        sp.R1 = plan->R[0];
        sp.R2 = plan->R[1];
    }
    sp.scale = scale;
    sp.has_c = c != nullptr;
    const int key[6] = {sp.n0, sp.N1, sp.N2, NR, plan->zchunk, Cfg::TX};
    if (memcmp(key, plan->work_key, sizeof(key)) != 0) {
        std::vector<S8Work> work;
        build_work8(NR, Cfg::TX, 148 * Cfg::CTAS_PER_SM, sp.n0, sp.N1, sp.N2, plan->zchunk, work);
        ODIL_REQUIRE(!work.empty() && (int64_t)work.size() <= kPartialCapacity, "star work list has %lld entries",
                     (long long)work.size());
        if ((int)work.size() > plan->work_cap) {
            if (plan->work_dev) cudaFree(plan->work_dev);
            plan->work_dev = nullptr;
            plan->work_cap = 0;
            ODIL_CUDA(cudaMalloc(&plan->work_dev, sizeof(S8Work) * work.size()));
            plan->work_cap = (int)work.size();
        }
        // pageable source: the copy is staged before the call returns, so `work` may go out of scope
        ODIL_CUDA(cudaMemcpyAsync(plan->work_dev, work.data(), sizeof(S8Work) * work.size(), cudaMemcpyHostToDevice, st));
        plan->work_n = (int)work.size();
        memcpy(plan->work_key, key, sizeof(key));
    }
    sp.work = (const S8Work*)plan->work_dev;
    const int64_t planes_total = (int64_t)sp.n0 + 2 * sp.halo;
    const int64_t plane_elems = (int64_t)sp.N1 * sp.N2;
    CUtensorMap tmU, tmC;
    if (int rc = make_plane_map<T>(&tmU, U - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RU, Cfg::BX)) return rc;
    if (int rc = make_plane_map<T>(&tmC, (c ? c : U) - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RC, Cfg::BX))
        return rc;
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star8<T, VW, NR, XU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    k_star8<T, VW, NR, XU><<<plan->work_n, Cfg::NT, Cfg::SMEM, st>>>(tmU, tmC, sp);
    {
        launch_counter()++;
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFuncAttributes fa;
            memset(&fa, 0, sizeof(fa));
            cudaFuncGetAttributes(&fa, k_star8<T, VW, NR, XU>);
            cudaGetLastError();
            return fail("k_star8<NR=%d> launch (%d CTAs x %d threads, %zu B dynamic smem) -> %s [numRegs=%d maxThreadsPerBlock=%d "
                        "static smem=%zu maxDynamic=%d local=%zu]",
                        NR, plan->work_n, Cfg::NT, (size_t)Cfg::SMEM, cudaGetErrorString(e), fa.numRegs, fa.maxThreadsPerBlock,
                        fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.localSizeBytes);
        }
    }
    *nparts = plan->work_n;
    return 0;
}

static void star_tile(int variant, int& TY, int& TX) {
    switch (variant) {
        case 1: TY = 16; TX = 64; break;
        case 2: TY = 16; TX = 128; break;
        case 3: TY = 4; TX = 128; break;
        default: TY = 8; TX = 128; break;
    }
}


// ------------------------------------------------------------------------------------------------
// 2-D tile kernel (tile2d.cuh)
// ------------------------------------------------------------------------------------------------
static bool tile2d_ok(const odil_b200_plan* plan, const odil_b200_slab* slab) {
    if (!plan->use_tile2d || plan->ndim != 2) return false;
    if (slab->halo > 0 || slab->n0 != plan->shape[0] || slab->z0 != 0) return false;
    if (plan->h2[0] > kT2MaxRadius || plan->h2[1] > kT2MaxRadius || plan->ncls > 255) return false;
    const int64_t N0 = plan->shape[0], N1 = plan->shape[1];
    if (N0 * N1 >= (1ll << 31) || N0 < 1 || N1 < 1) return false;
    const int64_t gx = (N1 + kT2X - 1) / kT2X, gy = (N0 + kT2Y - 1) / kT2Y;
    if (gy > 65535 || gx * gy > kPartialCapacity) return false;
    return (size_t)plan->ncls * plan->noff * 8 <= 64 * 1024;
}

template <typename T, int MODE>
static int launch_tile2d(const odil_b200_plan* plan, const T* A, const T* c, T scale, T* out, T* Fout, int* nparts,
                         cudaStream_t st) {
    Tile2Params<T> p;
    p.A = A;
    p.c = c;
    p.out = out;
    p.Fout = Fout;
    p.table = (const T*)plan->table_dev;
    p.partials = plan->partials;
    p.scale = scale;
    p.N0 = (int)plan->shape[0];
    p.N1 = (int)plan->shape[1];
    p.R0 = plan->R[0];
    p.R1 = plan->R[1];
    p.H0 = plan->h2[0];
    p.H1 = plan->h2[1];
    p.noff = plan->noff;
    p.ncls = plan->ncls;
    int AH, AW, FH, FW;
    t2_dims<MODE>(p.H0, p.H1, AH, AW, FH, FW);
    p.magicA = (unsigned)((1ull << 32) / (unsigned)AW + 1);
    p.magicF = (unsigned)((1ull << 32) / (unsigned)FW + 1);
    for (int o = 0; o < ODIL_B200_MAX_OFFSETS; ++o) {
        p.dy[o] = o < plan->noff ? (signed char)plan->off[o][0] : 0;
        p.dx[o] = o < plan->noff ? (signed char)plan->off[o][1] : 0;
    }
    const size_t smem = t2_smem_bytes<T, MODE>(p.H0, p.H1, p.ncls, p.noff);
    static size_t smem_set = 48 * 1024;  // per template instantiation
    if (smem > smem_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_tile2d<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    dim3 grid((p.N1 + kT2X - 1) / kT2X, (p.N0 + kT2Y - 1) / kT2Y);
    k_tile2d<T, MODE><<<grid, kT2Threads, smem, st>>>(p);
    ODIL_LAUNCHED();
    if (nparts) *nparts = (int)(grid.x * grid.y);
    return 0;
}


// ------------------------------------------------------------------------------------------------
// 3-D marching tile kernel for non-star plans (tile3d.cuh)
// ------------------------------------------------------------------------------------------------
static int tile3d_zchunk(const odil_b200_plan* plan) {
    if (plan->zchunk > 0) return (int)std::min<int64_t>(plan->zchunk, plan->shape[0]);
    // enough CTAs for a few waves of 148 SMs x 2-4 resident CTAs, chunks long enough to hide the 4*H0 lead-in planes
    const int64_t tiles = ((plan->shape[1] + kT3Y - 1) / kT3Y) * ((plan->shape[2] + kT3X - 1) / kT3X);
    int64_t zc = 64;
    while (zc > 16 && tiles * ((plan->shape[0] + zc - 1) / zc) < 148 * 8) zc /= 2;
    return (int)std::min<int64_t>(zc, plan->shape[0]);
}

static bool tile3d_ok(const odil_b200_plan* plan, const odil_b200_slab* slab) {
    if (!plan->use_tile3d || plan->ndim != 3 || plan->kind != 0) return false;
    if (slab->halo > 0 || slab->n0 != plan->shape[0] || slab->z0 != 0) return false;
    for (int a = 0; a < 3; ++a)
        if (plan->h3[a] > kT3MaxRadius || plan->shape[a] < 1 || plan->shape[a] >= (1 << 30)) return false;
    if (plan->ncls > 255 || plan->shape[1] * plan->shape[2] >= (1ll << 31)) return false;
    const int64_t gx = (plan->shape[2] + kT3X - 1) / kT3X, gy = (plan->shape[1] + kT3Y - 1) / kT3Y;
    const int zc = tile3d_zchunk(plan);
    const int64_t gz = (plan->shape[0] + zc - 1) / zc;
    return gy <= 65535 && gz <= 65535 && gx * gy * gz <= kPartialCapacity;
}

template <typename T>
static int launch_tile3d(const odil_b200_plan* plan, const T* U, const T* c, T scale, T* G, T* Fout, int* nparts,
                         cudaStream_t st) {
    Tile3Params<T> p;
    p.U = U;
    p.c = c;
    p.G = G;
    p.Fout = Fout;
    p.table = (const T*)plan->table_dev;
    p.partials = plan->partials;
    p.scale = scale;
    p.N0 = (int)plan->shape[0];
    p.N1 = (int)plan->shape[1];
    p.N2 = (int)plan->shape[2];
    p.R0 = plan->R[0];
    p.R1 = plan->R[1];
    p.R2 = plan->R[2];
    p.H0 = plan->h3[0];
    p.H1 = plan->h3[1];
    p.H2 = plan->h3[2];
    p.noff = plan->noff;
    p.ncls = plan->ncls;
    p.zchunk = tile3d_zchunk(plan);
    const Tile3Dims d = t3_dims(p.H0, p.H1, p.H2);
    p.magicA = (unsigned)((1ull << 32) / (unsigned)d.AW + 1);
    p.magicF = (unsigned)((1ull << 32) / (unsigned)d.FW + 1);
    for (int o = 0; o < ODIL_B200_MAX_OFFSETS; ++o) {
        p.dz[o] = o < plan->noff ? (signed char)plan->off[o][0] : 0;
        p.dy[o] = o < plan->noff ? (signed char)plan->off[o][1] : 0;
        p.dx[o] = o < plan->noff ? (signed char)plan->off[o][2] : 0;
    }
    const size_t smem = t3_smem_bytes<T>(p.H0, p.H1, p.H2, p.ncls, p.noff);
    ODIL_REQUIRE(smem <= 227 * 1024, "tile3d: %zu bytes of shared memory needed", smem);
    static size_t smem_set = 48 * 1024;  // per template instantiation
    if (smem > smem_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_tile3d<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    dim3 grid((p.N2 + kT3X - 1) / kT3X, (p.N1 + kT3Y - 1) / kT3Y, (p.N0 + p.zchunk - 1) / p.zchunk);
    k_tile3d<T><<<grid, kT3Threads, smem, st>>>(p);
    ODIL_LAUNCHED();
    *nparts = (int)(grid.x * grid.y * grid.z);
    return 0;
}

template <typename T>
static int run_fused(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U, const void* c,
                     double scale, void* G, void* Fout, double* sumsq, cudaStream_t st) {
    GenParams gp;
    fill_gen_params(plan, slab, gp);
    GenIO<T> io;
    io.U = (const T*)U;
    io.c = (const T*)c;
    io.out = (T*)G;
    io.Fout = (T*)Fout;
    io.table = (const T*)plan->table_dev;
    io.scale = (T)scale;
    int nparts = 0;
    if (tile2d_ok(plan, slab)) {
        if (int rc = launch_tile2d<T, 2>(plan, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)) return rc;
        k_reduce_partials<<<1, 1024, 0, st>>>(plan->partials, nparts, sumsq);
        ODIL_LAUNCHED();
        return 0;
    }
    if (tile3d_ok(plan, slab)) {
        if (int rc = launch_tile3d<T>(plan, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)) return rc;
        k_reduce_partials<<<1, 1024, 0, st>>>(plan->partials, nparts, sumsq);
        ODIL_LAUNCHED();
        return 0;
    }
    const bool slab_mode = slab->halo > 0 || slab->n0 != plan->shape[0] || slab->z0 != 0;
    bool tiled = plan->kind == 1 && !(plan->ndim == 2 && slab_mode);
    const int64_t n2 = plan->shape[plan->ndim - 1];
    const bool v3 = tiled && plan->use_v3 && (n2 % 4 == 0) && ((uintptr_t)U % 16 == 0) && ((uintptr_t)G % 16 == 0) &&
                    ((uintptr_t)c % 16 == 0) && ((uintptr_t)Fout % 16 == 0);
    bool tma = v3 && plan->use_tma && plan->wrap_free && get_encode_tiled() != nullptr;
    if (tma) {  // the three plane rings must fit the 227 KB of shared memory of one SM
        static const int tys[4] = {16, 8, 12, 26};
        const size_t need = 128 + (size_t)10 * (((tys[(plan->variant < 0 ? 1 : plan->variant) & 3] + 4) * (128 + 8) + 31) / 32 * 32) * sizeof(T) + 512 * sizeof(T) + 128;
        if (need > 227 * 1024) tma = false;
    }
    constexpr int VW7 = 16 / (int)sizeof(T);
    const bool star7 = tiled && plan->use_star7 && plan->wrap_free && get_encode_tiled() != nullptr && (n2 % VW7 == 0) &&
                       ((uintptr_t)U % 16 == 0) && ((uintptr_t)G % 16 == 0) && ((uintptr_t)c % 16 == 0) &&
                       ((uintptr_t)Fout % 16 == 0);
    if (star7 && plan->use_star8) {
        int rc;
#define ODIL_S8(NR_, XU_) launch_star8<T, VW7, NR_, XU_>(plan, slab, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)
        const bool xu = plan->star_xu && !plan->no_xu;
        switch (plan->variant) {
            case 1: rc = xu ? ODIL_S8(12, true) : ODIL_S8(12, false); break;
            case 2: rc = xu ? ODIL_S8(8, true) : ODIL_S8(8, false); break;
            default: rc = xu ? ODIL_S8(14, true) : ODIL_S8(14, false); break;
        }
#undef ODIL_S8
        if (rc) return rc;
    } else if (star7) {
        int rc;
#define ODIL_S7(TY_, XU_) launch_star7<T, VW7, TY_, XU_>(plan, slab, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)
        const bool xu = plan->star_xu && !plan->no_xu;
        switch (plan->variant) {
            case 1: rc = xu ? ODIL_S7(16, true) : ODIL_S7(16, false); break;
            case 2: rc = xu ? ODIL_S7(8, true) : ODIL_S7(8, false); break;
            default: rc = xu ? ODIL_S7(12, true) : ODIL_S7(12, false); break;
        }
#undef ODIL_S7
        if (rc) return rc;
    } else if (tma) {
        StarTmaParams<T> sp;
        sp.G = io.out;
        sp.Fout = io.Fout;
        sp.partials = plan->partials;
        sp.table = (const T*)plan->star_table;
        if (plan->ndim == 3) {
            sp.n0 = (int)slab->n0;
            sp.N0g = (int)plan->shape[0];
            sp.z0 = (int)slab->z0;
            sp.halo = slab->halo;
            sp.N1 = (int)plan->shape[1];
            sp.N2 = (int)plan->shape[2];
            sp.R0 = plan->R[0];
            sp.R1 = plan->R[1];
            sp.R2 = plan->R[2];
        } else {
            sp.n0 = 1;
            sp.N0g = 1;
            sp.z0 = 0;
            sp.halo = 0;
            sp.N1 = (int)plan->shape[0];
            sp.N2 = (int)plan->shape[1];
            sp.R0 = 0;
            sp.R1 = plan->R[0];
            sp.R2 = plan->R[1];
        }
        for (int i = 0; i < 7; ++i) sp.w[i] = (T)plan->w[i];
        sp.scale = (T)scale;
        sp.has_c = io.c != nullptr;
        int TY = 16;
        const int TX = 128;
        const int tvar = plan->variant < 0 ? 1 : plan->variant;  // auto: TY = 8
        switch (tvar) {
            case 1: TY = 8; break;
            case 2: TY = 12; break;
            case 3: TY = 26; break;
            default: TY = 16; break;
        }
        const int gx = (sp.N2 + TX - 1) / TX, gy = (sp.N1 + TY - 1) / TY;
        int zchunk = plan->zchunk;
        if (zchunk <= 0) {
            zchunk = 128;
            while (zchunk > 16 && (int64_t)gx * gy * ((sp.n0 + zchunk - 1) / zchunk) < 148 * 2 * 4) zchunk /= 2;
        }
        if (zchunk > sp.n0) zchunk = sp.n0;
        if (zchunk < 1) zchunk = 1;
        sp.zchunk = zchunk;
        const int gz = (sp.n0 + zchunk - 1) / zchunk;
        ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity, "star grid exceeds the partials workspace");
        dim3 grid(gx, gy, gz);
        const int64_t planes_total = (int64_t)sp.n0 + 2 * sp.halo;
        const int64_t plane_elems = (int64_t)sp.N1 * sp.N2;
        CUtensorMap tmU, tmC;
        if (int rc = make_plane_map<T>(&tmU, io.U - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, TY + 4, TX + 8))
            return rc;
        if (int rc = make_plane_map<T>(&tmC, (io.c ? io.c : io.U) - sp.halo * plane_elems, planes_total, sp.N1, sp.N2,
                                       TY + 4, TX + 8))
            return rc;
        int rc = 0;
        switch (tvar) {
            case 1: rc = launch_star_tma<T, 8, 128>(tmU, tmC, sp, grid, st); break;
            case 2: rc = launch_star_tma<T, 12, 128>(tmU, tmC, sp, grid, st); break;
            case 3: rc = launch_star_tma<T, 26, 128>(tmU, tmC, sp, grid, st); break;
            default: rc = launch_star_tma<T, 16, 128>(tmU, tmC, sp, grid, st); break;
        }
        if (rc) return rc;
        nparts = gx * gy * gz;
    } else if (v3) {
        StarV3Params<T> sp;
        sp.U = io.U;
        sp.c = io.c;
        sp.G = io.out;
        sp.Fout = io.Fout;
        sp.partials = plan->partials;
        sp.table = (const T*)plan->star_table;
        if (plan->ndim == 3) {
            sp.n0 = slab->n0;
            sp.N0g = plan->shape[0];
            sp.z0 = slab->z0;
            sp.halo = slab->halo;
            sp.N1 = (int)plan->shape[1];
            sp.N2 = (int)plan->shape[2];
            sp.R0 = plan->R[0];
            sp.R1 = plan->R[1];
            sp.R2 = plan->R[2];
            sp.has_z = plan->w[1] != 0.0 || plan->w[2] != 0.0 || plan->star_has_z;
        } else {
            sp.n0 = 1;
            sp.N0g = 1;
            sp.z0 = 0;
            sp.halo = 0;
            sp.N1 = (int)plan->shape[0];
            sp.N2 = (int)plan->shape[1];
            sp.R0 = 0;
            sp.R1 = plan->R[0];
            sp.R2 = plan->R[1];
            sp.has_z = 0;
        }
        for (int i = 0; i < 7; ++i) sp.w[i] = (T)plan->w[i];
        sp.scale = (T)scale;
        int TY, TX;
        const int vvar = plan->variant < 0 ? 0 : plan->variant;  // auto: TY = 16
        star_v3_tile(vvar, TY, TX);
        const int gx = (sp.N2 + TX - 1) / TX, gy = (sp.N1 + TY - 1) / TY;
        int zchunk = plan->zchunk;
        if (zchunk <= 0) {
            zchunk = 128;
            while (zchunk > 16 && (int64_t)gx * gy * ((sp.n0 + zchunk - 1) / zchunk) < 148 * 2 * 4) zchunk /= 2;
        }
        if (zchunk > sp.n0) zchunk = (int)sp.n0;
        if (zchunk < 1) zchunk = 1;
        sp.zchunk = zchunk;
        const int gz = (int)((sp.n0 + zchunk - 1) / zchunk);
        ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity, "star grid exceeds the partials workspace");
        dim3 grid(gx, gy, gz);
        int rc = 0;
        switch (vvar) {
            case 1: rc = launch_star_v3<T, 8, 128>(sp, grid, st); break;
            case 2: rc = launch_star_v3<T, 12, 128>(sp, grid, st); break;
            case 3: rc = launch_star_v3<T, 26, 128>(sp, grid, st); break;
            default: rc = launch_star_v3<T, 16, 128>(sp, grid, st); break;
        }
        if (rc) return rc;
        nparts = gx * gy * gz;
    } else if (tiled) {
        StarParams<T> sp;
        sp.U = io.U;
        sp.c = io.c;
        sp.G = io.out;
        sp.Fout = io.Fout;
        sp.partials = plan->partials;
        if (plan->ndim == 3) {
            sp.n0 = slab->n0;
            sp.N0g = plan->shape[0];
            sp.z0 = slab->z0;
            sp.halo = slab->halo;
            sp.N1 = (int)plan->shape[1];
            sp.N2 = (int)plan->shape[2];
            sp.R0 = plan->R[0];
            sp.R1 = plan->R[1];
            sp.R2 = plan->R[2];
            sp.has_z = plan->w[1] != 0.0 || plan->w[2] != 0.0;
        } else {
            sp.n0 = 1;
            sp.N0g = 1;
            sp.z0 = 0;
            sp.halo = 0;
            sp.N1 = (int)plan->shape[0];
            sp.N2 = (int)plan->shape[1];
            sp.R0 = 0;
            sp.R1 = plan->R[0];
            sp.R2 = plan->R[1];
            sp.has_z = 0;
        }
        sp.wc = (T)plan->w[0];
        sp.wzm = (T)plan->w[1];
        sp.wzp = (T)plan->w[2];
        sp.wym = (T)plan->w[3];
        sp.wyp = (T)plan->w[4];
        sp.wxm = (T)plan->w[5];
        sp.wxp = (T)plan->w[6];
        sp.scale = (T)scale;
        int TY, TX;
        const int wvar = plan->variant < 0 ? 2 : plan->variant;  // auto: 16 x 128 tiles
        star_tile(wvar, TY, TX);
        const int gx = (sp.N2 + TX - 1) / TX, gy = (sp.N1 + TY - 1) / TY;
        int zchunk = plan->zchunk;
        if (zchunk <= 0) {
            // aim for >= ~8 waves of 148 SMs x 4 CTAs, but keep the z-redundancy (2 extra planes) small
            zchunk = 64;
            while (zchunk > 16 && (int64_t)gx * gy * ((sp.n0 + zchunk - 1) / zchunk) < 148 * 4 * 6) zchunk /= 2;
        }
        if (zchunk > sp.n0) zchunk = (int)sp.n0;
        if (zchunk < 1) zchunk = 1;
        sp.zchunk = zchunk;
        const int gz = (int)((sp.n0 + zchunk - 1) / zchunk);
        ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity / 2, "star grid exceeds the partials workspace");
        dim3 grid(gx, gy, gz);
        const bool vec = (sp.N2 % 4 == 0) && ((uintptr_t)G % 16 == 0);
        int rc = 0;
        switch (wvar) {
            case 1: rc = launch_star_cfg<T, 16, 64, 256>(sp, grid, vec, st); break;
            case 2: rc = launch_star_cfg<T, 16, 128, 512>(sp, grid, vec, st); break;
            case 3: rc = launch_star_cfg<T, 4, 128, 128>(sp, grid, vec, st); break;
            default: rc = launch_star_cfg<T, 8, 128, 256>(sp, grid, vec, st); break;
        }
        if (rc) return rc;
        nparts = gx * gy * gz;
        // Boundary shell: per-cell class lookup, overwrites g within 2r of a face, adds the loss
        // of the cells whose own row is a boundary row.
        BoxList shell;
        int thick[ODIL_B200_MAX_NDIM];
        for (int a = 0; a < plan->ndim; ++a) thick[a] = 2 * plan->R[a];
        shell_boxes(plan, slab, thick, shell);
        gp.count_mode = 1;
        io.partials = plan->partials + nparts;
        int nb = 0;
        rc = launch_generic<T, 2>(plan, gp, shell, io, st, &nb);
        if (rc) return rc;
        ODIL_REQUIRE(nparts + nb <= kPartialCapacity, "partials workspace overflow");
        nparts += nb;
    } else {
        BoxList all;
        whole_box(plan, slab, all);
        gp.count_mode = 0;
        io.partials = plan->partials;
        int nb = 0;
        int rc = launch_generic<T, 2>(plan, gp, all, io, st, &nb);
        if (rc) return rc;
        nparts = nb;
    }
    k_reduce_partials<<<1, 1024, 0, st>>>(plan->partials, nparts, sumsq);
    ODIL_LAUNCHED();
    return 0;
}

template <typename T, int VW, int NR>
static int worklist_for(int n0, int N1, int N2, int zchunk, int32_t* out, int cap) {
    using Cfg = Star8Cfg<T, VW, NR>;
    std::vector<S8Work> work;
    build_work8(NR, Cfg::TX, 148 * Cfg::CTAS_PER_SM, n0, N1, N2, zchunk, work);
    for (size_t i = 0; i < work.size() && (int)i < cap; ++i) {
        out[5 * i + 0] = work[i].tx0;
        out[5 * i + 1] = work[i].ty0;
        out[5 * i + 2] = work[i].nrows;
        out[5 * i + 3] = work[i].zs;
        out[5 * i + 4] = work[i].ze;
    }
    return (int)work.size();
}

}  // namespace odil

extern "C" {

int odil_b200_version(void) { return 100; }
const char* odil_b200_last_error(void) { return last_error_ref().c_str(); }
int64_t odil_b200_launch_count(void) { return launch_counter().load(); }

int odil_b200_stencil_plan_create(int ndim, const int64_t* shape, int dtype, int noff, const int32_t* offsets,
                                  const int32_t* rwidth, const double* table, odil_b200_plan** out) {
    ODIL_REQUIRE(out != nullptr, "plan out pointer is null");
    *out = nullptr;
    ODIL_REQUIRE(ndim >= 1 && ndim <= ODIL_B200_MAX_NDIM, "ndim=%d unsupported (1..%d)", ndim, ODIL_B200_MAX_NDIM);
    ODIL_REQUIRE(noff >= 1 && noff <= ODIL_B200_MAX_OFFSETS, "noff=%d unsupported (1..%d)", noff,
                 ODIL_B200_MAX_OFFSETS);
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    odil_b200_plan* p = new odil_b200_plan();
    p->ndim = ndim;
    p->dtype = dtype;
    p->noff = noff;
    p->ncls = 1;
    p->zchunk = 0;
    p->variant = -1;  // auto: best measured tile per kernel at 512^3 fp32 on B200 (tools/bench_kernels.py)
    p->table_dev = nullptr;
    p->partials = nullptr;
    int64_t total = 1;
    for (int a = 0; a < ndim; ++a) {
        p->shape[a] = shape[a];
        p->R[a] = rwidth[a];
        if (shape[a] < 1 || rwidth[a] < 0 || 2 * (int64_t)rwidth[a] > shape[a]) {
            delete p;
            return fail("axis %d: shape=%lld rwidth=%d invalid (need shape >= 2*rwidth)", a, (long long)shape[a],
                        rwidth[a]);
        }
        p->ncls *= 2 * rwidth[a] + 1;
        total *= shape[a];
    }
    p->rmax0 = 0;
    for (int o = 0; o < noff; ++o)
        for (int a = 0; a < ndim; ++a) {
            p->off[o][a] = offsets[o * ndim + a];
            if (std::abs(p->off[o][a]) >= shape[a] && shape[a] > 1 && p->off[o][a] != 0) {
                delete p;
                return fail("offset %d axis %d = %d exceeds the grid size", o, a, p->off[o][a]);
            }
            if (a == 0) p->rmax0 = std::max(p->rmax0, std::abs(p->off[o][a]));
        }
    p->table.assign(table, table + (size_t)p->ncls * noff);
    // wrap_free: every (class, offset) whose neighbour would cross the periodic boundary has a zero coefficient.
    {
        bool wf = true;
        std::vector<int> cstr(ndim);
        int cs2 = 1;
        for (int a = ndim - 1; a >= 0; --a) {
            cstr[a] = cs2;
            cs2 *= 2 * p->R[a] + 1;
        }
        for (int cl = 0; cl < p->ncls && wf; ++cl)
            for (int o = 0; o < noff && wf; ++o) {
                if (p->table[(size_t)cl * noff + o] == 0.0) continue;
                for (int a = 0; a < ndim; ++a) {
                    const int d = p->off[o][a];
                    if (d == 0) continue;
                    const int r = p->R[a];
                    const int c = (cl / cstr[a]) % (2 * r + 1);
                    bool crosses;
                    if (c < r)
                        crosses = c + d < 0;            // low row i = c
                    else if (c > r)
                        crosses = (2 * r - c) < d;      // high row i = N-1-(2r-c)
                    else
                        crosses = std::abs(d) > r;      // interior rows next to the boundary rows
                    if (crosses) {
                        wf = false;
                        break;
                    }
                }
            }
        p->wrap_free = wf ? 1 : 0;
    }
    // Tiled eligibility: 2-D / 3-D, every offset a unit star arm, grid not degenerate.
    p->kind = 0;
    p->star_table = nullptr;
    p->star_has_z = 0;
    p->use_v3 = 1;
    p->use_tma = 1;
    p->use_star7 = 1;
    p->star_xu = 0;
    p->no_xu = 0;
    p->use_star8 = 1;
    {
        const char* e = getenv("ODIL_B200_TILE2D");
        p->use_tile2d = !(e && e[0] == '0');
        p->h2[0] = p->h2[1] = 0;
        if (ndim == 2)
            for (int o = 0; o < noff; ++o)
                for (int a = 0; a < 2; ++a) p->h2[a] = std::max(p->h2[a], std::abs(p->off[o][a]));
        const char* e3 = getenv("ODIL_B200_TILE3D");
        p->use_tile3d = !(e3 && e3[0] == '0');
        p->h3[0] = p->h3[1] = p->h3[2] = 0;
        if (ndim == 3)
            for (int o = 0; o < noff; ++o)
                for (int a = 0; a < 3; ++a) p->h3[a] = std::max(p->h3[a], std::abs(p->off[o][a]));
    }
    p->work_dev = nullptr;
    p->work_cap = 0;
    p->work_n = 0;
    for (int i = 0; i < 6; ++i) p->work_key[i] = -1;
    std::vector<double> star;
    for (int i = 0; i < 7; ++i) p->w[i] = 0.0;
    if ((ndim == 3 || ndim == 2) && total >= 512) {
        bool ok = true;
        int cint = 0, cs = 1;
        for (int a = ndim - 1; a >= 0; --a) {
            cint += p->R[a] * cs;
            cs *= 2 * p->R[a] + 1;
            if (shape[a] < 4 * p->R[a] + 1 || shape[a] < 4) ok = false;
        }
        double w[7] = {0, 0, 0, 0, 0, 0, 0};
        const int base = ndim == 3 ? 0 : 1;  // axis a of the plan maps to tiled axis base + a
        for (int o = 0; o < noff && ok; ++o) {
            int nz = 0, ax = -1, sg = 0;
            for (int a = 0; a < ndim; ++a)
                if (p->off[o][a] != 0) {
                    ++nz;
                    ax = a;
                    sg = p->off[o][a];
                }
            const double v = p->table[(size_t)cint * noff + o];
            if (nz == 0)
                w[0] += v;
            else if (nz == 1 && (sg == 1 || sg == -1))
                w[1 + 2 * (base + ax) + (sg > 0 ? 1 : 0)] += v;
            else
                ok = false;
        }
        if (ok) {
            p->kind = 1;
            for (int i = 0; i < 7; ++i) p->w[i] = w[i];
            star.assign((size_t)p->ncls * 7, 0.0);
            for (int o = 0; o < noff; ++o) {
                int slot = 0;
                for (int a = 0; a < ndim; ++a)
                    if (p->off[o][a] != 0) slot = 1 + 2 * (base + a) + (p->off[o][a] > 0 ? 1 : 0);
                for (int cl = 0; cl < p->ncls; ++cl) {
                    star[(size_t)cl * 7 + slot] += p->table[(size_t)cl * noff + o];
                    if ((slot == 1 || slot == 2) && p->table[(size_t)cl * noff + o] != 0.0) p->star_has_z = 1;
                }
            }
            // x-uniform arms: for the interior classes along every axis but the last, the four z/y arm
            // coefficients are the same for every class along the last axis
            {
                const int C2 = 2 * p->R[ndim - 1] + 1;
                const int C1 = ndim >= 2 ? 2 * p->R[ndim - 2] + 1 : 1;
                const int cz_int = ndim == 3 ? p->R[0] : 0;  // interior class of axis 0 (3-D only)
                bool xu = true;
                for (int cy = 0; cy < C1 && xu; ++cy) {
                    const int rowbase = (cz_int * C1 + cy) * C2;
                    for (int cx = 1; cx < C2 && xu; ++cx)
                        for (int sl = 1; sl <= 4; ++sl)
                            if (star[(size_t)(rowbase + cx) * 7 + sl] != star[(size_t)rowbase * 7 + sl]) xu = false;
                }
                p->star_xu = xu ? 1 : 0;
            }
        }
    }
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) {
        const size_t esz = dtype == ODIL_B200_F32 ? 4 : 8;
        e = cudaMalloc(&p->table_dev, esz * p->table.size());
        if (e == cudaSuccess) {
            if (dtype == ODIL_B200_F32) {
                std::vector<float> tf(p->table.begin(), p->table.end());
                e = cudaMemcpy(p->table_dev, tf.data(), esz * tf.size(), cudaMemcpyHostToDevice);
            } else {
                e = cudaMemcpy(p->table_dev, p->table.data(), esz * p->table.size(), cudaMemcpyHostToDevice);
            }
        }
        if (e == cudaSuccess) e = cudaMalloc((void**)&p->partials, sizeof(double) * kPartialCapacity);
        if (e == cudaSuccess && !star.empty()) {
            e = cudaMalloc(&p->star_table, esz * star.size());
            if (e == cudaSuccess) {
                if (dtype == ODIL_B200_F32) {
                    std::vector<float> tf(star.begin(), star.end());
                    e = cudaMemcpy(p->star_table, tf.data(), esz * tf.size(), cudaMemcpyHostToDevice);
                } else {
                    e = cudaMemcpy(p->star_table, star.data(), esz * star.size(), cudaMemcpyHostToDevice);
                }
            }
        }
    }
    if (e != cudaSuccess) {
        if (p->table_dev) cudaFree(p->table_dev);
        if (p->partials) cudaFree(p->partials);
        if (p->star_table) cudaFree(p->star_table);
        delete p;
        return fail("plan_create: %s", cudaGetErrorString(e));
    }
    *out = p;
    return 0;
}

int odil_b200_stencil_plan_destroy(odil_b200_plan* plan) {
    if (!plan) return 0;
    if (plan->table_dev) cudaFree(plan->table_dev);
    if (plan->partials) cudaFree(plan->partials);
    if (plan->star_table) cudaFree(plan->star_table);
    if (plan->work_dev) cudaFree(plan->work_dev);
    delete plan;
    return 0;
}

int odil_b200_stencil_plan_kind(const odil_b200_plan* plan) { return plan ? plan->kind : -1; }

int odil_b200_star_worklist(int dtype, int variant, int64_t n0, int64_t N1, int64_t N2, int zchunk, int32_t* out, int cap) {
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    ODIL_REQUIRE(n0 >= 1 && N1 >= 1 && N2 >= 1 && n0 < (1 << 30) && N1 < (1 << 30) && N2 < (1 << 30), "bad shape");
    ODIL_REQUIRE(cap >= 0 && (out != nullptr || cap == 0), "bad output buffer");
    const int v = variant < 0 ? 0 : variant % 10;
    const int a = (int)n0, b = (int)N1, c = (int)N2;
    if (dtype == ODIL_B200_F32) {
        if (v == 1) return worklist_for<float, 4, 12>(a, b, c, zchunk, out, cap);
        if (v == 2) return worklist_for<float, 4, 8>(a, b, c, zchunk, out, cap);
        return worklist_for<float, 4, 14>(a, b, c, zchunk, out, cap);
    }
    if (v == 1) return worklist_for<double, 2, 12>(a, b, c, zchunk, out, cap);
    if (v == 2) return worklist_for<double, 2, 8>(a, b, c, zchunk, out, cap);
    return worklist_for<double, 2, 14>(a, b, c, zchunk, out, cap);
}

int odil_b200_stencil_plan_tune(odil_b200_plan* plan, int zchunk, int variant) {
    ODIL_REQUIRE(plan != nullptr, "null plan");
    ODIL_REQUIRE((variant >= 0 && variant <= 3) || (variant >= 10 && variant <= 13) ||
                     (variant >= 20 && variant <= 23) || (variant >= 30 && variant <= 32) ||
                     (variant >= 40 && variant <= 42) || (variant >= 50 && variant <= 52) ||
                     (variant >= 60 && variant <= 62) || variant == 70 || variant == 71 || variant == 80 ||
                     variant == 81 || variant == -1,
                 "variant=%d unknown", variant);
    plan->zchunk = zchunk;
    // 70 / 71: 2-D tile kernel on / off with the default 3-D choice; any explicit star variant also turns it off
    // (so the star kernels stay reachable on 2-D grids)
    if (variant == 80 || variant == 81) {  // 3-D marching tile kernel for non-star plans on / off
        plan->use_tile3d = variant == 80;
        return 0;
    }
    plan->use_tile2d = variant == 70 || variant == -1;
    if (variant >= 70) variant = -1;
    // -1 / 50..52: k_star8 (default; up to 14 / 12 / 8 rows per CTA); 60..62: same with per-cell z/y-arm
    // coefficients even when the plan allows the x-uniform form; 30..32 / 40..42: k_star7 (tiles of 12 / 16 / 8
    // rows, x-uniform / per-cell arms); 0..3: previous TMA-fed kernel; 10..13: v2 tile kernel + shell pass;
    // 20..23: LDG column-group kernel
    if (variant < 0) variant = 50;
    plan->use_star8 = variant >= 50;
    plan->use_star7 = variant >= 30;
    plan->no_xu = (variant >= 40 && variant < 50) || variant >= 60;
    plan->use_tma = variant < 10;
    plan->use_v3 = variant < 10 || variant >= 20;
    plan->variant = variant % 10;
    for (int i = 0; i < 6; ++i) plan->work_key[i] = -1;
    return 0;
}

static int check_slab(const odil_b200_plan* plan, const odil_b200_slab* slab, int need_halo) {
    ODIL_REQUIRE(plan && slab, "null plan/slab");
    ODIL_REQUIRE(slab->n0 >= 1 && slab->z0 >= 0 && slab->z0 + slab->n0 <= plan->shape[0],
                 "slab [%lld,+%lld) outside axis 0 of size %lld", (long long)slab->z0, (long long)slab->n0,
                 (long long)plan->shape[0]);
    if (slab->halo == 0)
        ODIL_REQUIRE(slab->n0 == plan->shape[0], "halo=0 requires the whole axis 0 (single-GPU semantics)");
    else
        ODIL_REQUIRE(slab->halo >= need_halo, "halo=%d too small, need %d", slab->halo, need_halo);
    return 0;
}

int odil_b200_stencil_forward(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U,
                              const void* F_in, void* F_out, void* stream) {
    if (int rc = check_slab(plan, slab, plan ? plan->rmax0 : 0)) return rc;
    ODIL_REQUIRE(U && F_out, "null array");
    if (tile2d_ok(plan, slab)) {
        if (plan->dtype == ODIL_B200_F32)
            return launch_tile2d<float, 0>(plan, (const float*)U, (const float*)F_in, 1.f, (float*)F_out, nullptr,
                                           nullptr, (cudaStream_t)stream);
        return launch_tile2d<double, 0>(plan, (const double*)U, (const double*)F_in, 1.0, (double*)F_out, nullptr,
                                        nullptr, (cudaStream_t)stream);
    }
    GenParams gp;
    fill_gen_params(plan, slab, gp);
    BoxList all;
    whole_box(plan, slab, all);
    if (plan->dtype == ODIL_B200_F32) {
        GenIO<float> io{(const float*)U, (const float*)F_in, (float*)F_out, nullptr, (const float*)plan->table_dev,
                        nullptr, 1.f};
        return launch_generic<float, 0>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
    }
    GenIO<double> io{(const double*)U, (const double*)F_in, (double*)F_out, nullptr, (const double*)plan->table_dev,
                     nullptr, 1.0};
    return launch_generic<double, 0>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
}

int odil_b200_stencil_adjoint(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* F, double scale,
                              const void* G_in, void* G_out, void* stream) {
    if (int rc = check_slab(plan, slab, plan ? plan->rmax0 : 0)) return rc;
    ODIL_REQUIRE(F && G_out, "null array");
    if (tile2d_ok(plan, slab)) {
        if (plan->dtype == ODIL_B200_F32)
            return launch_tile2d<float, 1>(plan, (const float*)F, (const float*)G_in, (float)scale, (float*)G_out,
                                           nullptr, nullptr, (cudaStream_t)stream);
        return launch_tile2d<double, 1>(plan, (const double*)F, (const double*)G_in, scale, (double*)G_out, nullptr,
                                        nullptr, (cudaStream_t)stream);
    }
    GenParams gp;
    fill_gen_params(plan, slab, gp);
    BoxList all;
    whole_box(plan, slab, all);
    if (plan->dtype == ODIL_B200_F32) {
        GenIO<float> io{(const float*)F, (const float*)G_in, (float*)G_out, nullptr, (const float*)plan->table_dev,
                        nullptr, (float)scale};
        return launch_generic<float, 1>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
    }
    GenIO<double> io{(const double*)F, (const double*)G_in, (double*)G_out, nullptr, (const double*)plan->table_dev,
                     nullptr, scale};
    return launch_generic<double, 1>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
}

int odil_b200_stencil_fused(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U, const void* c,
                            double scale, void* G_out, void* F_out, double* sumsq_out, void* stream) {
    if (int rc = check_slab(plan, slab, plan ? 2 * plan->rmax0 : 0)) return rc;
    ODIL_REQUIRE(U && G_out && sumsq_out, "null array");
    if (plan->dtype == ODIL_B200_F32)
        return run_fused<float>(plan, slab, U, c, scale, G_out, F_out, sumsq_out, (cudaStream_t)stream);
    return run_fused<double>(plan, slab, U, c, scale, G_out, F_out, sumsq_out, (cudaStream_t)stream);
}

}  // extern "C"
