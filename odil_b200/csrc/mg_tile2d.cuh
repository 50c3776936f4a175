// 2-D cell-centred multigrid transfers through shared-memory tiles (sm_100a).
//
// k_interp_add2t:      out = ffac * term + cfac * I(coarse)           (core.py:245-263 with :606-700)
// k_interp_adjoint2t:  g_coarse = scale * I^T g_fine                   (what AD of core.py:606-700 yields)
//
// One CTA owns 16 x 64 coarse cells = 32 x 128 fine cells.  The synthesis stages the JOINTLY padded coarse tile
// (2 u[clamp q] - u[reflect q], core.py:640-643, evaluated per staged element so edges and corners are exact) and
// every thread emits 16-byte fine vectors from separable integer weights (1,3)x(1,3)/16.  The transpose stages the
// fine tile (zero outside the grid), gathers the gradient of the PADDED coarse array gp[q] (taps 2q-1 .. 2q+2,
// weights 1,3,3,1 per axis) on the tile plus two cells, and folds the pad in closed form:
//     g[c] = 2 * sum_{q in Cl(c)} gp[q] - sum_{q in Rf(c)} gp[q],
// Cl(c) / Rf(c) = products over the axes of {c} plus the pad indices that clamp / reflect onto c
// (q = -1 clamps to 0 and reflects to 1; q = n clamps to n-1 and reflects to n-2).  Interior cells reduce to gp[c].
#pragma once
#include "common.cuh"

namespace odil {

constexpr int kM2Y = 16, kM2X = 64, kM2Threads = 256;
constexpr int kM2PW = kM2X + 2;                 // padded coarse tile pitch
constexpr int kM2FH = 2 * (kM2Y + 4) + 2;       // staged fine rows of the transpose
constexpr int kM2FW = 2 * (kM2X + 4) + 2;
constexpr int kM2GW = kM2X + 4;                 // gp tile pitch

__device__ __forceinline__ int m2_clamp(int q, int n) { return q < 0 ? 0 : (q > n - 1 ? n - 1 : q); }
__device__ __forceinline__ int m2_reflect(int q, int n) { return q < 0 ? 1 : (q > n - 1 ? n - 2 : q); }

template <typename T>
__global__ void __launch_bounds__(kM2Threads) k_interp_add2t(const T* __restrict__ coarse, T cfac,
                                                             const T* __restrict__ term, T ffac, T* __restrict__ out,
                                                             int n0, int n1) {
    __shared__ T sP[(kM2Y + 2) * kM2PW];
    const int tid = threadIdx.x;
    const int cy0 = blockIdx.y * kM2Y, cx0 = blockIdx.x * kM2X;
    for (int e = tid; e < (kM2Y + 2) * kM2PW; e += kM2Threads) {
        const int r = e / kM2PW, cc = e % kM2PW;
        const int qy = min(cy0 - 1 + r, n0), qx = min(cx0 - 1 + cc, n1);
        const int sy = m2_clamp(qy, n0), sx = m2_clamp(qx, n1);
        T v = __ldg(coarse + (int64_t)sy * n1 + sx);
        if (sy != qy || sx != qx) v = T(2) * v - __ldg(coarse + (int64_t)m2_reflect(qy, n0) * n1 + m2_reflect(qx, n1));
        sP[e] = v;
    }
    __syncthreads();
    const int fn1 = 2 * n1;
    constexpr int VPR = kM2X / 2;  // 16-byte... 4-cell vectors per fine tile row
    for (int vec = tid; vec < 2 * kM2Y * VPR; vec += kM2Threads) {
        const int fr = vec / VPR, vx = vec % VPR;
        const int I = fr >> 1, a = fr & 1;
        const int fy = 2 * cy0 + fr, lc = 2 * vx, fx = 2 * (cx0 + lc);
        if (fy >= 2 * n0 || fx >= fn1) continue;
        const T* near = sP + (I + 1) * kM2PW + lc;      // padded columns lc-1 .. lc+2 of row I
        const T* far = sP + (I + 2 * a) * kM2PW + lc;   // row I-1 (a = 0) or I+1 (a = 1)
        const T h0 = T(3) * near[0] + far[0], h1 = T(3) * near[1] + far[1];
        const T h2 = T(3) * near[2] + far[2], h3 = T(3) * near[3] + far[3];
        const T k = T(0.0625);
        Vec4<T> res;
        res.x = cfac * ((T(3) * h1 + h0) * k);
        res.y = cfac * ((T(3) * h1 + h2) * k);
        res.z = cfac * ((T(3) * h2 + h1) * k);
        res.w = cfac * ((T(3) * h2 + h3) * k);
        const int64_t lin = (int64_t)fy * fn1 + fx;
        if (term) {
            const Vec4<T> t = *reinterpret_cast<const Vec4<T>*>(term + lin);
            res.x += ffac * t.x;
            res.y += ffac * t.y;
            res.z += ffac * t.z;
            res.w += ffac * t.w;
        }
        *reinterpret_cast<Vec4<T>*>(out + lin) = res;
    }
}

template <typename T>
__global__ void __launch_bounds__(kM2Threads) k_interp_adjoint2t(const T* __restrict__ gf, T scale,
                                                                 T* __restrict__ gc, int n0, int n1) {
    extern __shared__ __align__(16) unsigned char m2_smem[];
    T* sG = reinterpret_cast<T*>(m2_smem);   // kM2FH x kM2FW fine values
    T* gp = sG + kM2FH * kM2FW;              // (kM2Y + 4) x kM2GW padded-coarse gradients
    const int tid = threadIdx.x;
    const int cy0 = blockIdx.y * kM2Y, cx0 = blockIdx.x * kM2X;
    const int fy0 = 2 * (cy0 - 2) - 1, fx0 = 2 * (cx0 - 2) - 1;
    const int fn0 = 2 * n0, fn1 = 2 * n1;
    for (int e = tid; e < kM2FH * kM2FW; e += kM2Threads) {
        const int r = e / kM2FW, cc = e % kM2FW;
        const int y = fy0 + r, x = fx0 + cc;
        T v = T(0);
        if (y >= 0 && y < fn0 && x >= 0 && x < fn1) v = __ldg(gf + (int64_t)y * fn1 + x);
        sG[e] = v;
    }
    __syncthreads();
    for (int e = tid; e < (kM2Y + 4) * kM2GW; e += kM2Threads) {
        const int r = e / kM2GW, cc = e % kM2GW;
        const int qy = cy0 - 2 + r, qx = cx0 - 2 + cc;
        T acc = T(0);
        if (qy >= -1 && qy <= n0 && qx >= -1 && qx <= n1) {
            const T* s = sG + (2 * r) * kM2FW + 2 * cc;
            T col[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                col[j] = (s[j] + s[3 * kM2FW + j]) + T(3) * (s[kM2FW + j] + s[2 * kM2FW + j]);
            acc = ((col[0] + col[3]) + T(3) * (col[1] + col[2])) * T(0.0625);
        }
        gp[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < kM2Y * kM2X; e += kM2Threads) {
        const int r = e / kM2X, cc = e % kM2X;
        const int cy = cy0 + r, cx = cx0 + cc;
        if (cy >= n0 || cx >= n1) continue;
        const int b = (r + 2) * kM2GW + cc + 2;
        const bool x_lo = cx == 0, x_hi = cx == n1 - 1, x_lo1 = cx == 1, x_hi1 = cx == n1 - 2;
        auto cl = [&](int at) {  // sum over the columns that clamp onto cx
            T s = gp[at];
            if (x_lo) s += gp[at - 1];
            if (x_hi) s += gp[at + 1];
            return s;
        };
        auto rf = [&](int at) {  // sum over the columns that reflect onto cx
            T s = gp[at];
            if (x_lo1) s += gp[at - 2];
            if (x_hi1) s += gp[at + 2];
            return s;
        };
        T scl = cl(b), srf = rf(b);
        if (cy == 0) scl += cl(b - kM2GW);
        if (cy == n0 - 1) scl += cl(b + kM2GW);
        if (cy == 1) srf += rf(b - 2 * kM2GW);
        if (cy == n0 - 2) srf += rf(b + 2 * kM2GW);
        gc[(int64_t)cy * n1 + cx] = scale * (T(2) * scl - srf);
    }
}

template <typename T>
inline size_t m2_adjoint_smem() {
    return (size_t)(kM2FH * kM2FW + (kM2Y + 4) * kM2GW) * sizeof(T);
}

}  // namespace odil
