// Earlier generations of the fused 3-D / 2-D star sweep, kept selectable through odil_b200_stencil_plan_tune as
// measured history and for periodic plans: k_star3d (tile kernel, interior row only, + boundary shell pass),
// k_star_v3 (column groups, boundary rows in the kernel), k_star_tma (TMA-fed rings).  The current kernels are
// star8.cuh (default) and star7.cuh.  Included by stencil.cu.
#pragma once
#include "common.cuh"

namespace odil {

// ------------------------------------------------------------------------------------------------
// Tiled star kernel
// ------------------------------------------------------------------------------------------------
#ifdef ODIL_B200_LEGACY
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
#endif  // ODIL_B200_LEGACY

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
// (smem_u32, mbar_init / _expect_tx / _wait and tma_load_3d live in tma.cuh)

#ifdef ODIL_B200_LEGACY
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

#endif  // ODIL_B200_LEGACY

}  // namespace odil
