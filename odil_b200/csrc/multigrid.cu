// Multigrid transfers (sm_100a): synthesis step  out = ffac*t + cfac*I(coarse), its exact transpose,
// and restriction.  Restates reference core.py:245-263 (multigrid_to_regular), :606-700
// (interp_to_finer, method "stack"/"conv" -- both pinned to the same answers) and :703-755
// (restrict_to_coarser).  The boundary rule is the JOINT pad  P = 2*symmetric(u) - reflect(u)
// (core.py:640-643): for an out-of-range tap q the value is 2*u[clamp(q)] - u[reflect(q)] with
// clamp/reflect applied to all axes at once, so corners are not the tensor product of 1-D rules.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace odil {

enum : int { LOC_C = 0, LOC_N = 1, LOC_DOT = 2 };

struct MgGeom {
    int ndim;
    int loc[ODIL_B200_MAX_NDIM];
    int64_t cn[ODIL_B200_MAX_NDIM];       // coarse global array shape
    int64_t fn[ODIL_B200_MAX_NDIM];       // fine global array shape
    int64_t cstride[ODIL_B200_MAX_NDIM];  // element strides (local == global except the axis-0 origin)
    int64_t fstride[ODIL_B200_MAX_NDIM];
};

__device__ __forceinline__ int64_t clampi(int64_t q, int64_t n) { return q < 0 ? 0 : (q > n - 1 ? n - 1 : q); }
__device__ __forceinline__ int64_t reflecti(int64_t q, int64_t n) {
    if (n == 1) return 0;
    return q < 0 ? 1 : (q > n - 1 ? n - 2 : q);
}

// Padded coarse value at global padded coords q (components in [-1, n]).
template <typename T>
__device__ __forceinline__ T padded_value(const MgGeom& g, const T* __restrict__ coarse, int64_t coarse_z0,
                                          const int64_t* q) {
    int64_t ls = 0, lr = 0;
    bool outside = false;
#pragma unroll
    for (int a = 0; a < ODIL_B200_MAX_NDIM; ++a) {
        if (a >= g.ndim) break;
        const int64_t n = g.cn[a];
        int64_t qs = clampi(q[a], n), qr = reflecti(q[a], n);
        outside |= (qs != q[a]);
        if (a == 0) {
            qs -= coarse_z0;
            qr -= coarse_z0;
        }
        ls += qs * g.cstride[a];
        lr += qr * g.cstride[a];
    }
    if (!outside) return __ldg(coarse + ls);
    return T(2) * __ldg(coarse + ls) - __ldg(coarse + lr);
}

// Interpolated value I(coarse) at the GLOBAL fine cell f (generic: any ndim <= 4, any loc).
template <typename T>
__device__ __noinline__ T interp_cell_generic(const MgGeom& g, const T* __restrict__ coarse, int64_t coarse_z0,
                                              const int64_t* f) {
    // taps per axis: (index, integer weight); denominators multiply to `den`
    int64_t q0[ODIL_B200_MAX_NDIM], q1[ODIL_B200_MAX_NDIM];
    int w0[ODIL_B200_MAX_NDIM], w1[ODIL_B200_MAX_NDIM];
    int den = 1;
#pragma unroll
    for (int a = 0; a < ODIL_B200_MAX_NDIM; ++a) {
        if (a >= g.ndim) break;
        const int64_t i = f[a] >> 1;
        const int par = (int)(f[a] & 1);
        if (g.loc[a] == LOC_C) {
            q0[a] = par ? i + 1 : i - 1;
            w0[a] = 1;
            q1[a] = i;
            w1[a] = 3;
            den *= 4;
        } else if (g.loc[a] == LOC_N) {
            q0[a] = i;
            w0[a] = 1;
            q1[a] = par ? i + 1 : i;
            w1[a] = 1;
            den *= 2;
        } else {
            q0[a] = f[a];
            w0[a] = 1;
            q1[a] = f[a];
            w1[a] = 0;
        }
    }
    T acc = T(0);
    const int ncombo = 1 << g.ndim;
    for (int m = 0; m < ncombo; ++m) {
        int64_t q[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
        int w = 1;
        for (int a = 0; a < g.ndim; ++a) {
            const bool second = (m >> a) & 1;
            q[a] = second ? q1[a] : q0[a];
            w *= second ? w1[a] : w0[a];
        }
        if (w == 0) continue;
        acc += T(w) * padded_value<T>(g, coarse, coarse_z0, q);
    }
    return acc / T(den);
}

// One thread per fine cell (generic path).
template <typename T>
__global__ void __launch_bounds__(256) k_interp_add(MgGeom g, const T* __restrict__ coarse, T cfac,
                                                    const T* __restrict__ term, T ffac, T* __restrict__ out,
                                                    int64_t fz_begin, int64_t nfz, int64_t out_z0, int64_t coarse_z0) {
    int64_t total = nfz;
    for (int a = 1; a < g.ndim; ++a) total *= g.fn[a];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    int64_t f[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
    int64_t rem = gid;
    for (int a = g.ndim - 1; a >= 1; --a) {
        f[a] = rem % g.fn[a];
        rem /= g.fn[a];
    }
    f[0] = fz_begin + rem;
    int64_t lin = 0;
    for (int a = 0; a < g.ndim; ++a) lin += (a == 0 ? f[0] - out_z0 : f[a]) * g.fstride[a];
    T res = cfac * interp_cell_generic<T>(g, coarse, coarse_z0, f);
    if (term) res += ffac * __ldg(term + lin);
    out[lin] = res;
}

// ------------------------------------------------------------------------------------------------
// Fast path for cell-centred fields: the last two axes are 'c', axis 0 is 'c' (CZ) or has size 1.
// One thread per COARSE cell: loads its 3x3x3 neighbourhood once and produces the 2x2x2 fine cells
// with separable integer weights (1,3) -- exact same weights as the reference, scaled by 1/64 at the
// end (a power of two).  Cells on a single face use per-axis linear extrapolation (= the joint pad
// when only one axis leaves the range); cells on an edge/corner take the generic per-cell path.
// ------------------------------------------------------------------------------------------------
struct Mg3 {
    int n0, n1, n2;        // coarse global shape (n0 == 1 and CZ == false for 2-D)
    int64_t cs0, cs1;      // coarse strides (elements); stride of axis 2 is 1
    int64_t fs0, fs1;      // fine strides
};

template <typename T>
struct alignas(sizeof(T) * 2) Pair {
    T a, b;
};

template <typename T, bool CZ>
__global__ void __launch_bounds__(128) k_interp_add3(MgGeom g, Mg3 m, const T* __restrict__ coarse, T cfac,
                                                     const T* __restrict__ term, T ffac, T* __restrict__ out,
                                                     int cz_begin, int ncz, int out_z0, int coarse_z0) {
    const int K = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    const int I = cz_begin + blockIdx.z;
    if (K >= m.n2 || J >= m.n1) return;
    const bool bz = CZ && (I == 0 || I == m.n0 - 1);
    const bool by = (J == 0 || J == m.n1 - 1);
    const bool bx = (K == 0 || K == m.n2 - 1);
    constexpr int NZ = CZ ? 2 : 1;
    T res[NZ][2][2];
    if ((int)bz + (int)by + (int)bx >= 2) {
        for (int a = 0; a < NZ; ++a)
            for (int b = 0; b < 2; ++b)
                for (int c = 0; c < 2; ++c) {
                    const int64_t f3[ODIL_B200_MAX_NDIM] = {CZ ? 2 * (int64_t)I + a : (int64_t)I, 2 * (int64_t)J + b,
                                                            2 * (int64_t)K + c, 0};
                    const int64_t f2[ODIL_B200_MAX_NDIM] = {2 * (int64_t)J + b, 2 * (int64_t)K + c, 0, 0};
                    res[a][b][c] = interp_cell_generic<T>(g, coarse, coarse_z0, g.ndim == 3 ? f3 : f2);
                }
    } else {
        constexpr int DZ = CZ ? 3 : 1;
        T v[DZ][3][3];
        const int km = max(K - 1, 0), kp = min(K + 1, m.n2 - 1);
        const int jm = max(J - 1, 0), jp = min(J + 1, m.n1 - 1);
#pragma unroll
        for (int dz = 0; dz < DZ; ++dz) {
            int ii = CZ ? min(max(I - 1 + dz, 0), m.n0 - 1) : I;
            const T* pz = coarse + (int64_t)(ii - coarse_z0) * m.cs0;
            const int jj[3] = {jm, J, jp};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const T* py = pz + (int64_t)jj[dy] * m.cs1;
                v[dz][dy][0] = __ldg(py + km);
                v[dz][dy][1] = __ldg(py + K);
                v[dz][dy][2] = __ldg(py + kp);
            }
        }
        // single-face linear extrapolation of the out-of-range neighbour (2*u[clamp] - u[reflect])
        if (bx) {
#pragma unroll
            for (int dz = 0; dz < DZ; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    if (K == 0) v[dz][dy][0] = T(2) * v[dz][dy][1] - v[dz][dy][2];
                    if (K == m.n2 - 1) v[dz][dy][2] = T(2) * v[dz][dy][1] - v[dz][dy][0];
                }
        }
        if (by) {
#pragma unroll
            for (int dz = 0; dz < DZ; ++dz)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    if (J == 0) v[dz][0][dx] = T(2) * v[dz][1][dx] - v[dz][2][dx];
                    if (J == m.n1 - 1) v[dz][2][dx] = T(2) * v[dz][1][dx] - v[dz][0][dx];
                }
        }
        if (CZ && bz) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    if (I == 0) v[0][dy][dx] = T(2) * v[DZ > 1 ? 1 : 0][dy][dx] - v[DZ - 1][dy][dx];
                    if (I == m.n0 - 1) v[DZ - 1][dy][dx] = T(2) * v[DZ > 1 ? 1 : 0][dy][dx] - v[0][dy][dx];
                }
        }
        // separable accumulation with integer weights
        T ax[DZ][3][2];
#pragma unroll
        for (int dz = 0; dz < DZ; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                ax[dz][dy][0] = v[dz][dy][0] + T(3) * v[dz][dy][1];
                ax[dz][dy][1] = T(3) * v[dz][dy][1] + v[dz][dy][2];
            }
        T ay[DZ][2][2];
#pragma unroll
        for (int dz = 0; dz < DZ; ++dz)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                ay[dz][0][c] = ax[dz][0][c] + T(3) * ax[dz][1][c];
                ay[dz][1][c] = T(3) * ax[dz][1][c] + ax[dz][2][c];
            }
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (CZ) {
                    res[0][b][c] = (ay[0][b][c] + T(3) * ay[DZ > 1 ? 1 : 0][b][c]) * T(1.0 / 64.0);
                    res[NZ - 1][b][c] = (T(3) * ay[DZ > 1 ? 1 : 0][b][c] + ay[DZ - 1][b][c]) * T(1.0 / 64.0);
                } else {
                    res[0][b][c] = ay[0][b][c] * T(1.0 / 16.0);
                }
            }
    }
#pragma unroll
    for (int a = 0; a < NZ; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t fz = CZ ? 2 * (int64_t)I + a : (int64_t)I;
            const int64_t lin = (fz - out_z0) * m.fs0 + (int64_t)(2 * J + b) * m.fs1 + 2 * K;
            T r0 = cfac * res[a][b][0], r1 = cfac * res[a][b][1];
            if (term) {
                const Pair<T> t = *reinterpret_cast<const Pair<T>*>(term + lin);  // 2K is even: aligned pair
                r0 += ffac * t.a;
                r1 += ffac * t.b;
            }
            *reinterpret_cast<Pair<T>*>(out + lin) = Pair<T>{r0, r1};
        }
}

// Gather of fine gradient onto the PADDED coarse index q (separable 4-tap / 3-tap rule, clipped to
// the fine array), i.e. the plain transpose of the interpolation weights on the padded grid.
template <typename T>
__device__ __forceinline__ T gather_fine(const MgGeom& g, const T* __restrict__ gf, int64_t fine_z0,
                                         const int64_t* q) {
    // per-axis tap lists
    int64_t fi[ODIL_B200_MAX_NDIM][4];
    T fw[ODIL_B200_MAX_NDIM][4];
    int nt[ODIL_B200_MAX_NDIM];
#pragma unroll
    for (int a = 0; a < ODIL_B200_MAX_NDIM; ++a) {
        if (a >= g.ndim) {
            nt[a] = 0;
            continue;
        }
        int k = 0;
        const int64_t n = g.fn[a];
        if (g.loc[a] == LOC_C) {
            const int64_t b = 2 * q[a];
            const T ww[4] = {T(0.25), T(0.75), T(0.75), T(0.25)};
            for (int t = 0; t < 4; ++t) {
                const int64_t fidx = b - 1 + t;
                if (fidx >= 0 && fidx < n) {
                    fi[a][k] = fidx;
                    fw[a][k] = ww[t];
                    ++k;
                }
            }
        } else if (g.loc[a] == LOC_N) {
            const int64_t b = 2 * q[a];
            const T ww[3] = {T(0.5), T(1), T(0.5)};
            for (int t = 0; t < 3; ++t) {
                const int64_t fidx = b - 1 + t;
                if (fidx >= 0 && fidx < n) {
                    fi[a][k] = fidx;
                    fw[a][k] = ww[t];
                    ++k;
                }
            }
        } else {
            fi[a][0] = q[a];
            fw[a][0] = T(1);
            k = 1;
        }
        nt[a] = k;
    }
    T acc = T(0);
    const int n1 = g.ndim > 1 ? nt[1] : 1, n2 = g.ndim > 2 ? nt[2] : 1, n3 = g.ndim > 3 ? nt[3] : 1;
    for (int t0 = 0; t0 < nt[0]; ++t0) {
        const int64_t l0 = (fi[0][t0] - fine_z0) * g.fstride[0];
        const T w0 = fw[0][t0];
        for (int t1 = 0; t1 < n1; ++t1) {
            const int64_t l1 = g.ndim > 1 ? l0 + fi[1][t1] * g.fstride[1] : l0;
            const T w1 = g.ndim > 1 ? w0 * fw[1][t1] : w0;
            for (int t2 = 0; t2 < n2; ++t2) {
                const int64_t l2 = g.ndim > 2 ? l1 + fi[2][t2] * g.fstride[2] : l1;
                const T w2 = g.ndim > 2 ? w1 * fw[2][t2] : w1;
                for (int t3 = 0; t3 < n3; ++t3) {
                    const int64_t l3 = g.ndim > 3 ? l2 + fi[3][t3] * g.fstride[3] : l2;
                    const T w3 = g.ndim > 3 ? w2 * fw[3][t3] : w2;
                    acc += w3 * __ldg(gf + l3);
                }
            }
        }
    }
    return acc;
}

// (I^T g_fine)[J] for the GLOBAL coarse cell J:  sum_{q: clamp(q)=J} 2 G(q) - sum_{q: reflect(q)=J} G(q).
template <typename T>
__device__ __noinline__ T adjoint_cell_generic(const MgGeom& g, const T* __restrict__ gf, int64_t fine_z0,
                                               const int64_t* J) {
    int64_t cand[ODIL_B200_MAX_NDIM][3];
    int nc[ODIL_B200_MAX_NDIM];
    bool boundary = false;
    for (int a = 0; a < g.ndim; ++a) {
        const int64_t n = g.cn[a];
        int k = 0;
        cand[a][k++] = J[a];
        if (g.loc[a] == LOC_C) {
            if (J[a] <= 1) cand[a][k++] = -1;
            if (J[a] >= n - 2) cand[a][k++] = n;
        }
        nc[a] = k;
        boundary |= k > 1;
    }
    if (!boundary) return gather_fine<T>(g, gf, fine_z0, J);
    T acc = T(0);
    const int n1 = g.ndim > 1 ? nc[1] : 1, n2 = g.ndim > 2 ? nc[2] : 1, n3 = g.ndim > 3 ? nc[3] : 1;
    for (int c0 = 0; c0 < nc[0]; ++c0)
        for (int c1 = 0; c1 < n1; ++c1)
            for (int c2 = 0; c2 < n2; ++c2)
                for (int c3 = 0; c3 < n3; ++c3) {
                    int64_t q[ODIL_B200_MAX_NDIM] = {cand[0][c0], g.ndim > 1 ? cand[1][c1] : 0,
                                                     g.ndim > 2 ? cand[2][c2] : 0, g.ndim > 3 ? cand[3][c3] : 0};
                    bool mc = true, mr = true, outside = false;
                    for (int a = 0; a < g.ndim; ++a) {
                        mc = mc && clampi(q[a], g.cn[a]) == J[a];
                        mr = mr && reflecti(q[a], g.cn[a]) == J[a];
                        outside |= q[a] < 0 || q[a] > g.cn[a] - 1;
                    }
                    // in-range q is the plain value u[q]: coefficient 1 (=2-1) when q == J
                    T coef = outside ? T((mc ? 2 : 0) - (mr ? 1 : 0)) : T(1);
                    if (coef != T(0)) acc += coef * gather_fine<T>(g, gf, fine_z0, q);
                }
    return acc;
}

// One thread per coarse cell (generic path).
template <typename T>
__global__ void __launch_bounds__(128) k_interp_adjoint(MgGeom g, const T* __restrict__ gf, T scale,
                                                        T* __restrict__ gc, int64_t cz_begin, int64_t ncz,
                                                        int64_t out_z0, int64_t fine_z0) {
    int64_t total = ncz;
    for (int a = 1; a < g.ndim; ++a) total *= g.cn[a];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    int64_t J[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
    int64_t rem = gid;
    for (int a = g.ndim - 1; a >= 1; --a) {
        J[a] = rem % g.cn[a];
        rem /= g.cn[a];
    }
    J[0] = cz_begin + rem;
    int64_t lin = 0;
    for (int a = 0; a < g.ndim; ++a) lin += (a == 0 ? J[0] - out_z0 : J[a]) * g.cstride[a];
    gc[lin] = scale * adjoint_cell_generic<T>(g, gf, fine_z0, J);
}

// 1-D transposed-interpolation weights of coarse cell J (axis size n) on the SIX fine cells
// 2J-2 .. 2J+3: interior [0,1,3,3,1,0]/4, clipped to the fine array, plus the pad corrections
// (coarse 0: +2/4 on fine 0; coarse 1: -1/4 on fine 0; mirrored at the high end).  This is the
// separable rule; it equals the joint rule unless two or more axes are at a boundary.
template <typename T>
__device__ __forceinline__ void adjoint_weights(int J, int n, T* w) {
    const int nf = 2 * n;
    const int f0 = 2 * J - 2;
    const T base[6] = {T(0), T(0.25), T(0.75), T(0.75), T(0.25), T(0)};
#pragma unroll
    for (int t = 0; t < 6; ++t) w[t] = (f0 + t >= 0 && f0 + t < nf) ? base[t] : T(0);
    if (J == 0) w[2] += T(0.5);
    if (J == 1) w[0] -= T(0.25);
    if (J == n - 1) w[3] += T(0.5);
    if (J == n - 2) w[5] -= T(0.25);
}

// Fast transposed interpolation, cell-centred, one thread per coarse cell, 6x6(x6) window read as
// aligned pairs along x.
template <typename T, bool CZ>
__global__ void __launch_bounds__(128) k_interp_adjoint3(MgGeom g, Mg3 m, const T* __restrict__ gf, T scale,
                                                         T* __restrict__ gc, int cz_begin, int out_z0, int fine_z0) {
    const int K = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    const int I = cz_begin + blockIdx.z;
    if (K >= m.n2 || J >= m.n1) return;
    const bool bz = CZ && (I <= 1 || I >= m.n0 - 2);
    const bool by = (J <= 1 || J >= m.n1 - 2);
    const bool bx = (K <= 1 || K >= m.n2 - 2);
    T acc;
    if ((int)bz + (int)by + (int)bx >= 2) {
        const int64_t J3[ODIL_B200_MAX_NDIM] = {I, J, K, 0};
        const int64_t J2[ODIL_B200_MAX_NDIM] = {J, K, 0, 0};
        acc = adjoint_cell_generic<T>(g, gf, fine_z0, g.ndim == 3 ? J3 : J2);
    } else {
        T wx[6], wy[6], wz[6];
        adjoint_weights<T>(K, m.n2, wx);
        adjoint_weights<T>(J, m.n1, wy);
        if (CZ) adjoint_weights<T>(I, m.n0, wz);
        const int nf1 = 2 * m.n1, nf2 = 2 * m.n2;
        const int nf0 = CZ ? 2 * m.n0 : m.n0;
        // clamp the x window into the array (weights of clipped cells are zero)
        const int x0 = 2 * K - 2;
        acc = T(0);
        constexpr int TZ = CZ ? 6 : 1;
#pragma unroll
        for (int tz = 0; tz < TZ; ++tz) {
            const int fz = CZ ? 2 * I - 2 + tz : I;
            if (CZ && (wz[tz] == T(0) || fz < 0 || fz >= nf0)) continue;
            const T* pz = gf + (int64_t)(fz - fine_z0) * m.fs0;
            T accy = T(0);
#pragma unroll
            for (int ty = 0; ty < 6; ++ty) {
                const int fy = 2 * J - 2 + ty;
                if (wy[ty] == T(0) || fy < 0 || fy >= nf1) continue;
                const T* py = pz + (int64_t)fy * m.fs1;
                T accx = T(0);
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const int fx = x0 + 2 * p;
                    if (fx >= 0 && fx + 1 < nf2) {
                        const Pair<T> v = *reinterpret_cast<const Pair<T>*>(py + fx);  // fx is even: aligned pair
                        accx += wx[2 * p] * v.a + wx[2 * p + 1] * v.b;
                    }
                }
                accy += wy[ty] * accx;
            }
            acc += CZ ? wz[tz] * accy : accy;
        }
    }
    const int64_t lin = (int64_t)(I - out_z0) * m.cs0 + (int64_t)J * m.cs1 + K;
    gc[lin] = scale * acc;
}

// One thread per coarse cell; joint pad on the fine array (only 'n' axes ever leave the range).
template <typename T>
__global__ void __launch_bounds__(256) k_restrict(MgGeom g /* cn = coarse(out), fn = fine(in) */,
                                                  const T* __restrict__ fine, T* __restrict__ out) {
    int64_t total = 1;
    for (int a = 0; a < g.ndim; ++a) total *= g.cn[a];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    int64_t J[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
    int64_t rem = gid;
    for (int a = g.ndim - 1; a >= 0; --a) {
        J[a] = rem % g.cn[a];
        rem /= g.cn[a];
    }
    int64_t ti[ODIL_B200_MAX_NDIM][3];
    T tw[ODIL_B200_MAX_NDIM][3];
    int nt[ODIL_B200_MAX_NDIM] = {1, 1, 1, 1};
    for (int a = 0; a < g.ndim; ++a) {
        if (g.loc[a] == LOC_C) {
            ti[a][0] = 2 * J[a];
            ti[a][1] = 2 * J[a] + 1;
            tw[a][0] = tw[a][1] = T(0.5);
            nt[a] = 2;
        } else if (g.loc[a] == LOC_N) {
            ti[a][0] = 2 * J[a] - 1;
            ti[a][1] = 2 * J[a];
            ti[a][2] = 2 * J[a] + 1;
            tw[a][0] = tw[a][2] = T(0.25);
            tw[a][1] = T(0.5);
            nt[a] = 3;
        } else {
            ti[a][0] = J[a];
            tw[a][0] = T(1);
            nt[a] = 1;
        }
    }
    // geometry for padded_value expects "cn"/"cstride" to describe the array being read -> swap roles
    MgGeom r = g;
    for (int a = 0; a < g.ndim; ++a) {
        r.cn[a] = g.fn[a];
        r.cstride[a] = g.fstride[a];
    }
    T acc = T(0);
    for (int t0 = 0; t0 < nt[0]; ++t0)
        for (int t1 = 0; t1 < nt[1]; ++t1)
            for (int t2 = 0; t2 < nt[2]; ++t2)
                for (int t3 = 0; t3 < nt[3]; ++t3) {
                    int64_t q[ODIL_B200_MAX_NDIM] = {ti[0][t0], g.ndim > 1 ? ti[1][t1] : 0, g.ndim > 2 ? ti[2][t2] : 0,
                                                     g.ndim > 3 ? ti[3][t3] : 0};
                    T w = tw[0][t0];
                    if (g.ndim > 1) w *= tw[1][t1];
                    if (g.ndim > 2) w *= tw[2][t2];
                    if (g.ndim > 3) w *= tw[3][t3];
                    acc += w * padded_value<T>(r, fine, 0, q);
                }
    out[gid] = acc;
}

static int make_geom(int ndim, const int64_t* cshape, const char* loc, MgGeom& g) {
    ODIL_REQUIRE(ndim >= 1 && ndim <= ODIL_B200_MAX_NDIM, "ndim=%d unsupported", ndim);
    ODIL_REQUIRE(loc != nullptr, "null loc");
    g.ndim = ndim;
    for (int a = 0; a < ndim; ++a) {
        ODIL_REQUIRE(cshape[a] >= 1, "bad coarse shape");
        g.cn[a] = cshape[a];
        switch (loc[a]) {
            case 'c': g.loc[a] = LOC_C; g.fn[a] = 2 * cshape[a]; break;
            case 'n': g.loc[a] = LOC_N; g.fn[a] = 2 * (cshape[a] - 1) + 1; break;
            case '.': g.loc[a] = LOC_DOT; g.fn[a] = cshape[a]; break;
            default: return fail("loc[%d]='%c' invalid (expected c, n or .)", a, loc[a]);
        }
    }
    int64_t cs = 1, fs = 1;
    for (int a = ndim - 1; a >= 0; --a) {
        g.cstride[a] = cs;
        g.fstride[a] = fs;
        cs *= g.cn[a];
        fs *= g.fn[a];
    }
    for (int a = ndim; a < ODIL_B200_MAX_NDIM; ++a) {
        g.loc[a] = LOC_DOT;
        g.cn[a] = g.fn[a] = 1;
        g.cstride[a] = g.fstride[a] = 0;
    }
    return 0;
}

// Fast-path eligibility: cell-centred last two axes; axis 0 cell-centred (3-D), untouched ('.') or absent (2-D).
static bool fast3_geometry(const MgGeom& g, Mg3& m, bool& cz) {
    if (g.ndim == 3) {
        if (g.loc[1] != LOC_C || g.loc[2] != LOC_C) return false;
        if (g.loc[0] == LOC_C)
            cz = true;
        else if (g.loc[0] == LOC_DOT)
            cz = false;
        else
            return false;
        m.n0 = (int)g.cn[0];
        m.n1 = (int)g.cn[1];
        m.n2 = (int)g.cn[2];
    } else if (g.ndim == 2) {
        if (g.loc[0] != LOC_C || g.loc[1] != LOC_C) return false;
        cz = false;
        m.n0 = 1;
        m.n1 = (int)g.cn[0];
        m.n2 = (int)g.cn[1];
    } else {
        return false;
    }
    if (m.n1 < 4 || m.n2 < 4 || (cz && m.n0 < 4)) return false;
    if (g.cn[0] > (1 << 29) || g.cn[1] > (1 << 29) || (g.ndim > 2 && g.cn[2] > (1 << 29))) return false;
    m.cs1 = m.n2;
    m.cs0 = (int64_t)m.n1 * m.n2;
    m.fs1 = 2 * (int64_t)m.n2;
    m.fs0 = 4 * (int64_t)m.n1 * m.n2;
    return true;
}

}  // namespace odil
#include "mg_march.cuh"
#include "mg_tile2d.cuh"
#include "mg_adj_tma.cuh"

namespace odil {

// ------------------------------------------------------------------------------------------------
// k_adam_synth3 -- fusion across the optimizer seam, the other way round from k_interp_adjoint3m_adam: the Adam update
// of the FINEST multigrid term t0 (optimizer.py:311-319) also produces the regular field of the NEXT evaluation,
//     U = ffac * t0_new + cfac * I(V1)        (core.py:245-263; V1 = the already synthesised level 1),
// so the synthesis of level 0 does not read t0 back from HBM: 32.5 instead of 36.5 bytes per fine cell for the pair
// (k_adam of level 0 + k_interp_add3m of level 0).  A thread owns the fine vector (4 cells along x) of one fine row and
// marches over kAsZC coarse planes (2 kAsZC fine planes), carrying the in-plane interpolation of three coarse planes in
// registers; the interpolation uses mg_plane() and the plane combination of k_interp_add3m statement for statement, the
// update is adam_one(): x, m, v and U are bit-identical to the unfused pair.  The coarse values (1/8 of the fine data)
// come through L1 / L2.  (First version: one fine vector per thread, 131072 short-lived CTAs at 512^3: 0.78 ms = 86 %
// of the measured HBM peak, no better than the unfused pair.)
// grid (ceil(n2 / 128), ceil(2 n1 / 4), ceil(n0 / kAsZC)), block (64, 4).
// ------------------------------------------------------------------------------------------------
constexpr int kAsZC = 4;  // coarse planes (pairs of fine planes) per CTA

template <typename T, int OCC>
__global__ void __launch_bounds__(256, OCC) k_adam_synth3(Mg3 mm, const T* __restrict__ coarse, T cfac, T ffac, T* __restrict__ x,
                                                     T* __restrict__ m, T* __restrict__ v, const T* __restrict__ g,
                                                     T* __restrict__ out, T alpha_host, const double* __restrict__ alpha_dev,
                                                     T omb1, T omb2, T eps, int cz_begin, int cz_end, int fine_z0,
                                                     int coarse_z0) {
    // cz_begin .. cz_end: coarse planes whose fine planes are updated (global numbers); fine_z0 / coarse_z0: global
    // plane number of local plane 0 of the fine arrays / of `coarse` (slabs; 0 for whole arrays)
    const int k = blockIdx.x * 64 + threadIdx.x;  // fine cells 4k .. 4k+3 = coarse cells 2k, 2k+1
    const int fy = blockIdx.y * 4 + threadIdx.y;
    if (2 * k >= mm.n2 || fy >= 2 * mm.n1) return;
    const int J = fy >> 1, b = fy & 1;
    const int Ibeg = cz_begin + blockIdx.z * kAsZC, Iend = min(Ibeg + kAsZC, cz_end);
    if (Ibeg >= Iend) return;
    const T alpha = alpha_dev ? (T)__ldg(alpha_dev) : alpha_host;
    const T s = cfac * T(1.0 / 64.0);
    // in-plane interpolation of the coarse planes I-1, I, I+1 for the own fine row (row b of mg_plane's pair), carried
    // along axis 0 like k_interp_add3m does
    auto plane_row = [&](int zp, T (&p)[4]) {
        const MgP<T> P = mg_plane<T>(mm, coarse, coarse_z0, zp, J, k);
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = b == 0 ? P.v[0][c] : P.v[1][c];
    };
    T Pm[4], Pc[4], Pp[4];
    plane_row(Ibeg - 1, Pm);
    plane_row(Ibeg, Pc);
    int64_t lin = (int64_t)(2 * Ibeg - fine_z0) * mm.fs0 + (int64_t)fy * mm.fs1 + 4 * k;
    for (int I = Ibeg; I < Iend; ++I) {
        // the eight streams of the two fine planes first (independent of the coarse loads below)
        MgVec4<T> xx[2], mv[2], vv[2], gg[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            xx[a] = mg_ld4<T>(x + lin + a * mm.fs0);
            mv[a] = mg_ld4<T>(m + lin + a * mm.fs0);
            vv[a] = mg_ld4<T>(v + lin + a * mm.fs0);
            gg[a] = mg_ld4<T>(g + lin + a * mm.fs0);
        }
        plane_row(I + 1, Pp);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            T r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) r[c] = s * (a == 0 ? Pm[c] + T(3) * Pc[c] : T(3) * Pc[c] + Pp[c]);
            adam_one(xx[a].x, mv[a].x, vv[a].x, gg[a].x, alpha, omb1, omb2, eps);
            adam_one(xx[a].y, mv[a].y, vv[a].y, gg[a].y, alpha, omb1, omb2, eps);
            adam_one(xx[a].z, mv[a].z, vv[a].z, gg[a].z, alpha, omb1, omb2, eps);
            adam_one(xx[a].w, mv[a].w, vv[a].w, gg[a].w, alpha, omb1, omb2, eps);
            r[0] = fma(ffac, xx[a].x, r[0]);
            r[1] = fma(ffac, xx[a].y, r[1]);
            r[2] = fma(ffac, xx[a].z, r[2]);
            r[3] = fma(ffac, xx[a].w, r[3]);
            const int64_t l = lin + a * mm.fs0;
            mg_st4(out + l, MgVec4<T>{r[0], r[1], r[2], r[3]});
            mg_st4(x + l, xx[a]);
            mg_st4(m + l, mv[a]);
            mg_st4(v + l, vv[a]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            Pm[c] = Pc[c];
            Pc[c] = Pp[c];
        }
        lin += 2 * mm.fs0;
    }
}

}  // namespace odil
namespace odil {

// 2-D cell-centred whole-array transfers through the shared-memory tile kernels (mg_tile2d.cuh).
static bool tile2_ok(const MgGeom& g, bool whole, const void* a, const void* b, const void* c) {
    static const bool enabled = [] {
        const char* e = getenv("ODIL_B200_MG2D");
        return !(e && e[0] == '0');
    }();
    if (!enabled || !whole || g.ndim != 2 || g.loc[0] != LOC_C || g.loc[1] != LOC_C) return false;
    if (g.cn[0] < 2 || g.cn[1] < 2 || g.cn[1] % 2 != 0) return false;
    if (4 * g.cn[0] * g.cn[1] >= (1ll << 31) || (g.cn[0] + kM2Y - 1) / kM2Y > 65535) return false;
    return ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)c % 16 == 0);
}

// z-chunk of the marching transfer kernels (CTAs of 128 threads, `ctas_per_sm` resident per SM by their
// __launch_bounds__): the number of chunks is chosen so that the CTAs fill the 148 SMs in (nearly) whole waves --
// the first version used 1792 CTAs on 888 slots = 2.02 waves, i.e. a third wave for 2 % of the work -- with
// chunks of at least 8 coarse planes so that the 2-plane lead-in stays small.
static int chunk_for_waves(int64_t layer, int ncz, int ctas_per_sm);
static int march_chunk(const Mg3& m, int ncz, int ctas_per_sm) {
    return chunk_for_waves((int64_t)((m.n2 / 2 + 31) / 32) * ((m.n1 + 3) / 4), ncz, ctas_per_sm);
}
static int chunk_for_waves(int64_t layer, int ncz, int ctas_per_sm) {
    const int64_t slots = 148 * (int64_t)ctas_per_sm;
    int best_zc = ncz;
    double best = -1.0;
    for (int gz = 1; gz <= std::max(1, ncz / 8); ++gz) {
        const int zc = (ncz + gz - 1) / gz;
        const int gzr = (ncz + zc - 1) / zc;
        const int64_t ctas = layer * gzr;
        const int64_t waves = (ctas + slots - 1) / slots;
        // time ~ waves * (planes per chunk + lead-in); fewer, longer chunks win ties
        const double cost = (double)waves * (zc + 2);
        if (best < 0 || cost < best * 0.999) {
            best = cost;
            best_zc = zc;
        }
    }
    return best_zc;
}
// resident CTAs per SM k_interp_adjoint3m<float> is compiled for: 4 (128 registers), 5 (96) or 6 (80, small spills)
static int adj_occ() {
    static const int occ = [] {
        const char* e = getenv("ODIL_B200_ADJ_OCC");
        const int v = e ? atoi(e) : 4;
        return v == 5 || v == 6 ? v : 4;
    }();
    return occ;
}

// TMA-fed transposed interpolation: dense fine array whose rows are a multiple of 16 bytes (n2 even: guaranteed by
// march_ok), a driver that exports cuTensorMapEncodeTiled; ODIL_B200_ADJ_TMA=0 keeps the LDG kernel.
static bool adj_tma_ok(const Mg3& m) {
    const char* e = getenv("ODIL_B200_ADJ_TMA");  // read per call: tests flip it to compare the two kernels
    const bool off = e && atoi(e) == 0;
    return !off && m.fs1 == 2 * (int64_t)m.n2 && m.fs0 == m.fs1 * 2 * m.n1 && get_encode_tiled() != nullptr;
}

static bool march_ok(const Mg3& m, bool cz, int ndim, const void* a, const void* b, const void* c) {
    if (8 * (int64_t)m.n0 * m.n1 * m.n2 >= (1ll << 31)) return false;  // 32-bit element offsets inside
    return cz && ndim == 3 && m.n2 % 2 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) &&
           ((uintptr_t)c % 16 == 0) && getenv("ODIL_B200_MG_OLD") == nullptr;
}

}  // namespace odil

using namespace odil;

extern "C" {

int odil_b200_mg_interp_add(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* coarse,
                            double cfac, const void* fine_term, double ffac, void* out,
                            const odil_b200_mg_range* range, void* stream) {
    MgGeom g;
    if (int rc = make_geom(ndim, cshape, loc, g)) return rc;
    ODIL_REQUIRE(coarse && out, "null array");
    odil_b200_mg_range r{0, g.fn[0], 0, 0};
    if (range) r = *range;
    ODIL_REQUIRE(r.fz_begin >= 0 && r.fz_end <= g.fn[0] && r.fz_begin <= r.fz_end, "bad fine plane range");
    int64_t total = r.fz_end - r.fz_begin;
    for (int a = 1; a < ndim; ++a) total *= g.fn[a];
    if (total == 0) return 0;
    const int64_t nb = (total + 255) / 256;
    ODIL_REQUIRE(nb < (1ll << 31), "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    if (tile2_ok(g, r.fz_begin == 0 && r.fz_end == g.fn[0] && r.out_z0 == 0 && r.coarse_z0 == 0, coarse, fine_term,
                 out)) {
        const int n0 = (int)g.cn[0], n1 = (int)g.cn[1];
        dim3 grid((n1 + kM2X - 1) / kM2X, (n0 + kM2Y - 1) / kM2Y);
        if (dtype == ODIL_B200_F32)
            k_interp_add2t<float><<<grid, kM2Threads, 0, st>>>((const float*)coarse, (float)cfac,
                                                               (const float*)fine_term, (float)ffac, (float*)out, n0, n1);
        else
            k_interp_add2t<double><<<grid, kM2Threads, 0, st>>>((const double*)coarse, cfac, (const double*)fine_term,
                                                                ffac, (double*)out, n0, n1);
        ODIL_LAUNCHED();
        return 0;
    }
    {
        Mg3 m;
        bool cz = false;
        const bool pairs = (r.fz_begin % 2 == 0) && (r.fz_end % 2 == 0);
        if (fast3_geometry(g, m, cz) && march_ok(m, cz, ndim, coarse, fine_term, out) && r.fz_end > r.fz_begin) {
            const int ib = (int)(r.fz_begin >> 1), ie = (int)(((r.fz_end - 1) >> 1) + 1);
            // ODIL_B200_ADD_PF=1: the variant that loads the fine term one step ahead (4 CTAs per SM)
            const char* epf = getenv("ODIL_B200_ADD_PF");
            const bool pf = dtype == ODIL_B200_F32 && epf && atoi(epf) != 0;
            const int zc = march_chunk(m, ie - ib, pf ? 4 : 5);
            dim3 block(32, 4, 1);
            dim3 grid((m.n2 / 2 + 31) / 32, (m.n1 + 3) / 4, (ie - ib + zc - 1) / zc);
            if (grid.y <= 65535 && grid.z <= 65535) {
                if (pf) {
                    k_interp_add3m<float, true><<<grid, block, 0, st>>>(m, (const float*)coarse, (float)cfac,
                                                                        (const float*)fine_term, (float)ffac, (float*)out,
                                                                        (int)r.fz_begin, (int)r.fz_end, (int)r.out_z0,
                                                                        (int)r.coarse_z0, zc);
                } else if (dtype == ODIL_B200_F32) {
                    k_interp_add3m<float><<<grid, block, 0, st>>>(m, (const float*)coarse, (float)cfac,
                                                                  (const float*)fine_term, (float)ffac, (float*)out,
                                                                  (int)r.fz_begin, (int)r.fz_end, (int)r.out_z0,
                                                                  (int)r.coarse_z0, zc);
                } else {
                    k_interp_add3m<double><<<grid, block, 0, st>>>(m, (const double*)coarse, cfac,
                                                                   (const double*)fine_term, ffac, (double*)out,
                                                                   (int)r.fz_begin, (int)r.fz_end, (int)r.out_z0,
                                                                   (int)r.coarse_z0, zc);
                }
                ODIL_LAUNCHED();
                return 0;
            }
        }
        if (fast3_geometry(g, m, cz) && (!cz || pairs) && !(ndim == 2 && range != nullptr)) {
            const int zb = (int)(ndim == 3 ? (cz ? r.fz_begin / 2 : r.fz_begin) : 0);
            const int nz = (int)(ndim == 3 ? (cz ? (r.fz_end - r.fz_begin) / 2 : r.fz_end - r.fz_begin) : 1);
            if (nz <= 65535) {
                dim3 block(64, 2, 1);
                dim3 grid((m.n2 + 63) / 64, (m.n1 + 1) / 2, nz);
                const int oz0 = (int)(ndim == 3 ? r.out_z0 : 0), cz0 = (int)(ndim == 3 ? r.coarse_z0 : 0);
                if (grid.y <= 65535) {
                    if (dtype == ODIL_B200_F32) {
                        if (cz)
                            k_interp_add3<float, true><<<grid, block, 0, st>>>(g, m, (const float*)coarse, (float)cfac,
                                                                             (const float*)fine_term, (float)ffac,
                                                                             (float*)out, zb, nz, oz0, cz0);
                        else
                            k_interp_add3<float, false><<<grid, block, 0, st>>>(g, m, (const float*)coarse, (float)cfac,
                                                                              (const float*)fine_term, (float)ffac,
                                                                              (float*)out, zb, nz, oz0, cz0);
                    } else {
                        if (cz)
                            k_interp_add3<double, true><<<grid, block, 0, st>>>(g, m, (const double*)coarse, cfac,
                                                                              (const double*)fine_term, ffac,
                                                                              (double*)out, zb, nz, oz0, cz0);
                        else
                            k_interp_add3<double, false><<<grid, block, 0, st>>>(g, m, (const double*)coarse, cfac,
                                                                               (const double*)fine_term, ffac,
                                                                               (double*)out, zb, nz, oz0, cz0);
                    }
                    ODIL_LAUNCHED();
                    return 0;
                }
            }
        }
    }
    if (dtype == ODIL_B200_F32)
        k_interp_add<float><<<(unsigned)nb, 256, 0, st>>>(g, (const float*)coarse, (float)cfac, (const float*)fine_term,
                                                          (float)ffac, (float*)out, r.fz_begin, r.fz_end - r.fz_begin,
                                                          r.out_z0, r.coarse_z0);
    else if (dtype == ODIL_B200_F64)
        k_interp_add<double><<<(unsigned)nb, 256, 0, st>>>(g, (const double*)coarse, cfac, (const double*)fine_term,
                                                           ffac, (double*)out, r.fz_begin, r.fz_end - r.fz_begin,
                                                           r.out_z0, r.coarse_z0);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    return 0;
}

// Adam update of the finest multigrid term (x, m, v with gradient g) AND out = ffac * x_new + cfac * I(coarse) in one pass
// (k_adam_synth3).  Returns 1 -- nothing done -- when the geometry is not cell-centred 3-D with an even coarse width
// and 16-byte aligned arrays: the caller then runs odil_b200_adam_step and odil_b200_mg_interp_add separately.  (A 2-D
// version, k_interp_add2t with the update in front of it, measured slower than the pair on configs[1] -- 47.1 vs 45.5 us
// per epoch -- and was removed.)
int odil_b200_adam_synth(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* coarse, double cfac,
                         double ffac, void* x, void* m_state, void* v_state, const void* g, void* out, double alpha,
                         const double* alpha_dev, double one_minus_beta1, double one_minus_beta2, double epsilon,
                         const odil_b200_mg_range* range, void* stream) {
    MgGeom geo;
    if (int rc = make_geom(ndim, cshape, loc, geo)) return rc;
    ODIL_REQUIRE(coarse && x && m_state && v_state && g && out, "null array");
    odil_b200_mg_range r{0, geo.fn[0], 0, 0};
    if (range) r = *range;
    ODIL_REQUIRE(r.fz_begin >= 0 && r.fz_end <= geo.fn[0] && r.fz_begin <= r.fz_end, "bad fine plane range");
    if (r.fz_begin % 2 != 0 || r.fz_end % 2 != 0) return 1;
    if (r.fz_begin == r.fz_end) return 0;
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    cudaStream_t st = (cudaStream_t)stream;
    Mg3 m;
    bool cz = false;
    if (!(fast3_geometry(geo, m, cz) && march_ok(m, cz, ndim, x, g, out) && (uintptr_t)m_state % 16 == 0 &&
          (uintptr_t)v_state % 16 == 0 && (uintptr_t)coarse % 16 == 0))
        return 1;
    dim3 block(64, 4, 1);
    const int cb = (int)(r.fz_begin / 2), ce = (int)(r.fz_end / 2), fz0 = (int)r.out_z0, cz0 = (int)r.coarse_z0;
    dim3 grid((m.n2 / 2 + 63) / 64, (2 * m.n1 + 3) / 4, (unsigned)((ce - cb + kAsZC - 1) / kAsZC));
    if (grid.y > 65535 || grid.z > 65535) return 1;
    // resident CTAs per SM the fp32 kernel is compiled for: 3 (80 registers, 16 bytes of spills; default) or 2 (128)
    static const int occ = [] {
        const char* e = getenv("ODIL_B200_SYNTH_OCC");
        return e && atoi(e) == 2 ? 2 : 3;
    }();
#define ODIL_SYNTH_F32(OCC_)                                                                                              \
    k_adam_synth3<float, OCC_><<<grid, block, 0, st>>>(m, (const float*)coarse, (float)cfac, (float)ffac, (float*)x,         \
                                                       (float*)m_state, (float*)v_state, (const float*)g, (float*)out,       \
                                                       (float)alpha, alpha_dev, (float)one_minus_beta1,                      \
                                                       (float)one_minus_beta2, (float)epsilon, cb, ce, fz0, cz0)
    if (dtype == ODIL_B200_F32) {
        if (occ == 2) ODIL_SYNTH_F32(2);
        else ODIL_SYNTH_F32(3);
    } else
        k_adam_synth3<double, 1><<<grid, block, 0, st>>>(m, (const double*)coarse, cfac, ffac, (double*)x, (double*)m_state,
                                                      (double*)v_state, (const double*)g, (double*)out, alpha, alpha_dev,
                                                      one_minus_beta1, one_minus_beta2, epsilon, cb, ce, fz0, cz0);
    ODIL_LAUNCHED();
    return 0;
}

// g_coarse = scale * I^T g_fine  AND  the Adam update of the fine array (x, m, v) with gradient g_fine, in one pass
// over g_fine (whole arrays, single GPU).  Returns 1 -- nothing done -- when the geometry is not the one the marching
// kernel handles (cell-centred 3-D, even coarse width, 16-byte aligned arrays): the caller then runs
// odil_b200_mg_interp_adjoint and odil_b200_adam_step separately.
int odil_b200_mg_interp_adjoint_adam(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* g_fine,
                                     double scale, void* g_coarse, void* x, void* m_state, void* v_state, double alpha,
                                     const double* alpha_dev, double one_minus_beta1, double one_minus_beta2,
                                     double epsilon, void* stream) {
    MgGeom g;
    if (int rc = make_geom(ndim, cshape, loc, g)) return rc;
    ODIL_REQUIRE(g_fine && g_coarse && x && m_state && v_state, "null array");
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    Mg3 m;
    bool cz = false;
    if (!(fast3_geometry(g, m, cz) && march_ok(m, cz, ndim, g_fine, g_coarse, x) && (uintptr_t)m_state % 16 == 0 &&
          (uintptr_t)v_state % 16 == 0))
        return 1;
    const int n0 = (int)g.cn[0];
    const int zc = march_chunk(m, n0, 4);
    dim3 block(32, 4, 1);
    dim3 grid((m.n2 / 2 + 31) / 32, (m.n1 + 3) / 4, (unsigned)((n0 + zc - 1) / zc));
    if (grid.y > 65535 || grid.z > 65535) return 1;
    const int nfix = 16 * n0 + 16 * (m.n1 + m.n2);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ODIL_B200_F32) {
        MgAdam<float> ad{(float*)x, (float*)m_state, (float*)v_state, (float)alpha, (float)one_minus_beta1,
                         (float)one_minus_beta2, (float)epsilon, alpha_dev};
        k_interp_adjoint3m_adam<float><<<grid, block, 0, st>>>(m, (const float*)g_fine, (float)scale, (float*)g_coarse, 0, n0,
                                                               zc, ad);
        k_adjoint_joint_fix<float><<<(nfix + 127) / 128, 128, 0, st>>>(m, (const float*)g_fine, (float)scale,
                                                                       (float*)g_coarse, 0, n0, 0, 0);
    } else {
        MgAdam<double> ad{(double*)x, (double*)m_state, (double*)v_state, alpha, one_minus_beta1, one_minus_beta2, epsilon,
                          alpha_dev};
        k_interp_adjoint3m_adam<double><<<grid, block, 0, st>>>(m, (const double*)g_fine, scale, (double*)g_coarse, 0, n0, zc,
                                                                ad);
        k_adjoint_joint_fix<double><<<(nfix + 127) / 128, 128, 0, st>>>(m, (const double*)g_fine, scale, (double*)g_coarse, 0,
                                                                        n0, 0, 0);
    }
    launch_counter()++;  // two launches
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_mg_interp_adjoint(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* g_fine,
                                double scale, void* g_coarse, const odil_b200_mg_adj_range* range, void* stream) {
    MgGeom g;
    if (int rc = make_geom(ndim, cshape, loc, g)) return rc;
    ODIL_REQUIRE(g_fine && g_coarse, "null array");
    odil_b200_mg_adj_range r{0, g.cn[0], 0, 0};
    if (range) r = *range;
    ODIL_REQUIRE(r.cz_begin >= 0 && r.cz_end <= g.cn[0] && r.cz_begin <= r.cz_end, "bad coarse plane range");
    int64_t total = r.cz_end - r.cz_begin;
    for (int a = 1; a < ndim; ++a) total *= g.cn[a];
    if (total == 0) return 0;
    const int64_t nb = (total + 127) / 128;
    ODIL_REQUIRE(nb < (1ll << 31), "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    if (tile2_ok(g, r.cz_begin == 0 && r.cz_end == g.cn[0] && r.out_z0 == 0 && r.fine_z0 == 0, g_fine, g_coarse,
                 nullptr)) {
        const int n0 = (int)g.cn[0], n1 = (int)g.cn[1];
        dim3 grid((n1 + kM2X - 1) / kM2X, (n0 + kM2Y - 1) / kM2Y);
        if (dtype == ODIL_B200_F32) {
            k_interp_adjoint2t<float><<<grid, kM2Threads, m2_adjoint_smem<float>(), st>>>(
                (const float*)g_fine, (float)scale, (float*)g_coarse, n0, n1);
        } else {
            static bool attr_set = false;
            if (!attr_set) {
                ODIL_CUDA(cudaFuncSetAttribute(k_interp_adjoint2t<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)m2_adjoint_smem<double>()));
                attr_set = true;
            }
            k_interp_adjoint2t<double><<<grid, kM2Threads, m2_adjoint_smem<double>(), st>>>(
                (const double*)g_fine, scale, (double*)g_coarse, n0, n1);
        }
        ODIL_LAUNCHED();
        return 0;
    }
    {
        Mg3 m;
        bool cz = false;
        if (fast3_geometry(g, m, cz) && march_ok(m, cz, ndim, g_fine, g_coarse, nullptr)) {
            const int occ = dtype == ODIL_B200_F32 ? adj_occ() : 4;
            const int zc = march_chunk(m, (int)(r.cz_end - r.cz_begin), occ);
            dim3 block(32, 4, 1);
            dim3 grid((m.n2 / 2 + 31) / 32, (m.n1 + 3) / 4, (unsigned)((r.cz_end - r.cz_begin + zc - 1) / zc));
            const int nfix = 16 * (int)(r.cz_end - r.cz_begin) + 16 * (m.n1 + m.n2);
            if (dtype == ODIL_B200_F32 && adj_tma_ok(m)) {
                // TMA-fed sweep (mg_adj_tma.cuh).  The tensor map covers exactly the fine planes the range may touch
                // (the masks of k_interp_adjoint3m); everything outside reads as zero.
                constexpr int NS = 4;
                using Cfg = A3Cfg<float, NS>;
                const int cb = (int)r.cz_begin, ce = (int)r.cz_end, nf0 = 2 * m.n0;
                const int lo = cb == 1 ? 0 : std::max(2 * cb - 1, 0);
                const int hi = ce - 1 == m.n0 - 2 ? nf0 - 1 : std::min(2 * ce, nf0 - 1);
                CUtensorMap tm;
                if (int rc = make_plane_map<float>(&tm, (const float*)g_fine + (int64_t)(lo - r.fine_z0) * m.fs0,
                                                   hi - lo + 1, 2 * m.n1, 2 * m.n2, kA3BY, kA3BX, 2))
                    return rc;
                const int64_t layer = (int64_t)((m.n2 / 2 + 31) / 32) * ((m.n1 + kA3RJ - 1) / kA3RJ);
                const int zt = chunk_for_waves(layer, ce - cb, 2);
                dim3 gt((m.n2 / 2 + 31) / 32, (m.n1 + kA3RJ - 1) / kA3RJ, (unsigned)((ce - cb + zt - 1) / zt));
                if (gt.y <= 65535 && gt.z <= 65535) {
                    static bool attr_set = false;
                    if (!attr_set) {
                        ODIL_CUDA(cudaFuncSetAttribute(k_interp_adjoint3t<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)Cfg::SMEM));
                        attr_set = true;
                    }
                    k_interp_adjoint3t<NS><<<gt, kA3Threads, Cfg::SMEM, st>>>(tm, m, (float)scale, (float*)g_coarse, cb, ce,
                                                                           (int)r.out_z0, lo, zt);
                    k_adjoint_joint_fix<float><<<(nfix + 127) / 128, 128, 0, st>>>(
                        m, (const float*)g_fine, (float)scale, (float*)g_coarse, cb, ce, (int)r.out_z0, (int)r.fine_z0);
                    launch_counter()++;  // two launches
                    ODIL_LAUNCHED();
                    return 0;
                }
            }
            if (grid.y <= 65535 && grid.z <= 65535) {
#define ODIL_ADJ(T_, OCC_)                                                                                             \
    k_interp_adjoint3m<T_, OCC_><<<grid, block, 0, st>>>(m, (const T_*)g_fine, (T_)scale, (T_*)g_coarse, (int)r.cz_begin, \
                                                         (int)r.cz_end, (int)r.out_z0, (int)r.fine_z0, zc)
                if (dtype == ODIL_B200_F32) {
                    if (occ == 5) ODIL_ADJ(float, 5);
                    else if (occ == 6) ODIL_ADJ(float, 6);
                    else ODIL_ADJ(float, 4);
                    k_adjoint_joint_fix<float><<<(nfix + 127) / 128, 128, 0, st>>>(
                        m, (const float*)g_fine, (float)scale, (float*)g_coarse, (int)r.cz_begin, (int)r.cz_end,
                        (int)r.out_z0, (int)r.fine_z0);
                } else {
                    ODIL_ADJ(double, 4);
                    k_adjoint_joint_fix<double><<<(nfix + 127) / 128, 128, 0, st>>>(
                        m, (const double*)g_fine, scale, (double*)g_coarse, (int)r.cz_begin, (int)r.cz_end,
                        (int)r.out_z0, (int)r.fine_z0);
                }
#undef ODIL_ADJ
                launch_counter()++;  // two launches
                ODIL_LAUNCHED();
                return 0;
            }
        }
        if (fast3_geometry(g, m, cz) && !(ndim == 2 && range != nullptr)) {
            const int zb = (int)(ndim == 3 ? r.cz_begin : 0);
            const int nz = (int)(ndim == 3 ? r.cz_end - r.cz_begin : 1);
            dim3 block(64, 2, 1);
            dim3 grid((m.n2 + 63) / 64, (m.n1 + 1) / 2, nz);
            const int oz0 = (int)(ndim == 3 ? r.out_z0 : 0), fz0 = (int)(ndim == 3 ? r.fine_z0 : 0);
            if (nz <= 65535 && grid.y <= 65535) {
                if (dtype == ODIL_B200_F32) {
                    if (cz)
                        k_interp_adjoint3<float, true><<<grid, block, 0, st>>>(g, m, (const float*)g_fine, (float)scale,
                                                                             (float*)g_coarse, zb, oz0, fz0);
                    else
                        k_interp_adjoint3<float, false><<<grid, block, 0, st>>>(g, m, (const float*)g_fine,
                                                                              (float)scale, (float*)g_coarse, zb, oz0,
                                                                              fz0);
                } else {
                    if (cz)
                        k_interp_adjoint3<double, true><<<grid, block, 0, st>>>(g, m, (const double*)g_fine, scale,
                                                                              (double*)g_coarse, zb, oz0, fz0);
                    else
                        k_interp_adjoint3<double, false><<<grid, block, 0, st>>>(g, m, (const double*)g_fine, scale,
                                                                               (double*)g_coarse, zb, oz0, fz0);
                }
                ODIL_LAUNCHED();
                return 0;
            }
        }
    }
    if (dtype == ODIL_B200_F32)
        k_interp_adjoint<float><<<(unsigned)nb, 128, 0, st>>>(g, (const float*)g_fine, (float)scale, (float*)g_coarse,
                                                              r.cz_begin, r.cz_end - r.cz_begin, r.out_z0, r.fine_z0);
    else if (dtype == ODIL_B200_F64)
        k_interp_adjoint<double><<<(unsigned)nb, 128, 0, st>>>(g, (const double*)g_fine, scale, (double*)g_coarse,
                                                               r.cz_begin, r.cz_end - r.cz_begin, r.out_z0, r.fine_z0);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_mg_restrict(int ndim, const int64_t* fshape, const char* loc, int dtype, const void* in, void* out,
                          void* stream) {
    ODIL_REQUIRE(ndim >= 1 && ndim <= ODIL_B200_MAX_NDIM && loc && fshape, "bad arguments");
    int64_t cshape[ODIL_B200_MAX_NDIM];
    for (int a = 0; a < ndim; ++a) {
        if (loc[a] == 'c')
            cshape[a] = fshape[a] / 2;
        else if (loc[a] == 'n')
            cshape[a] = (fshape[a] - 1) / 2 + 1;
        else
            cshape[a] = fshape[a];
    }
    MgGeom g;
    if (int rc = make_geom(ndim, cshape, loc, g)) return rc;
    // the fine array may be one larger than 2*coarse along odd-sized 'c' axes; take the given shape
    int64_t fs = 1;
    for (int a = ndim - 1; a >= 0; --a) {
        g.fn[a] = fshape[a];
        g.fstride[a] = fs;
        fs *= fshape[a];
    }
    ODIL_REQUIRE(in && out, "null array");
    int64_t total = 1;
    for (int a = 0; a < ndim; ++a) total *= g.cn[a];
    if (total == 0) return 0;
    const int64_t nb = (total + 255) / 256;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == ODIL_B200_F32)
        k_restrict<float><<<(unsigned)nb, 256, 0, st>>>(g, (const float*)in, (float*)out);
    else if (dtype == ODIL_B200_F64)
        k_restrict<double><<<(unsigned)nb, 256, 0, st>>>(g, (const double*)in, (double*)out);
    else
        return fail("dtype=%d unsupported", dtype);
    ODIL_LAUNCHED();
    return 0;
}

}  // extern "C"
