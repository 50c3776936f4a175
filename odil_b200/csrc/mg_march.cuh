// Marching versions of the cell-centred 3-D multigrid transfers (the two kernels that dominated the epoch
// after the stencil sweep: 0.72 ms + 1.19 ms of a 2.98 ms epoch at 512^3, at 33 % / 8 % of the HBM roofline).
//
//   k_interp_add3m      out = ffac * term + cfac * I(coarse)          (core.py:245-263, :606-700)
//   k_interp_adjoint3m  g_coarse = scale * I^T g_fine                 (what AD of core.py:606-700 produces)
//
// Both stream the FINE array exactly once with 16-byte accesses: a thread owns two coarse cells along x
// (= four fine cells, one vector) of one coarse row and marches along axis 0, carrying the in-plane partial
// results of the neighbouring planes in registers, so every coarse value / fine vector is loaded once per
// thread instead of 27 / 48 scalar loads per coarse cell.
// Boundary rule: the joint pad P = 2*symmetric(u) - reflect(u) (core.py:640-643).  With one axis out of range
// it is the linear extrapolation along that axis (done in registers); with two or more it is not separable and
// those few edge / corner values go through the generic per-value routines of multigrid.cu.
#pragma once

namespace odil {

template <typename T>
struct alignas(16) MgVec4 {
    T x, y, z, w;
};
template <typename T>
__device__ __forceinline__ MgVec4<T> mg_ld4(const T* p);
template <>
__device__ __forceinline__ MgVec4<float> mg_ld4<float>(const float* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return MgVec4<float>{v.x, v.y, v.z, v.w};
}
template <>
__device__ __forceinline__ MgVec4<double> mg_ld4<double>(const double* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return MgVec4<double>{a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ void mg_st4(float* p, const MgVec4<float>& v) {
    *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void mg_st4(double* p, const MgVec4<double>& v) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2*>(p)[1] = make_double2(v.z, v.w);
}
template <typename T>
__device__ __forceinline__ Pair<T> mg_ld2(const T* p) {
    return *reinterpret_cast<const Pair<T>*>(p);
}

// ------------------------------------------------------------------------------------------------
// Interpolation.  In-plane result of one padded coarse plane z' in [-1, n0]: P.v[b][c] = sum over the 3 x 4
// coarse neighbourhood with the integer weights (1,3) x (1,3) (scaled by 16), for the fine rows 2J+b and
// the fine cells 4k+c.  Everything is passed and returned BY VALUE so that the planes stay in registers.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct MgP {
    T v[2][4];
};
template <typename T>
struct MgNb {
    T v[3][4];
};

template <typename T>
__device__ __forceinline__ MgP<T> mg_reduce(const MgNb<T>& n) {
    T ax[3][4];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        ax[dy][0] = n.v[dy][0] + T(3) * n.v[dy][1];
        ax[dy][1] = T(3) * n.v[dy][1] + n.v[dy][2];
        ax[dy][2] = n.v[dy][1] + T(3) * n.v[dy][2];
        ax[dy][3] = T(3) * n.v[dy][2] + n.v[dy][3];
    }
    MgP<T> P;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        P.v[0][c] = ax[0][c] + T(3) * ax[1][c];
        P.v[1][c] = T(3) * ax[1][c] + ax[2][c];
    }
    return P;
}

// plane inside the array; out-of-range y / x neighbours by linear extrapolation along their axis (exact unless
// two axes leave the range at once -- those fine cells are rewritten by k_interp_fix_edges)
template <typename T>
__device__ __forceinline__ MgP<T> mg_plane_inrange(const T* __restrict__ coarse, int coarse_z0, int64_t cs0, int64_t cs1,
                                                   int n1, int n2, int zp, int J, int k) {
    MgNb<T> n;
    const T* pz = coarse + (int64_t)(zp - coarse_z0) * cs0;
    const int xl = max(2 * k - 1, 0), xr = min(2 * k + 2, n2 - 1);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int jj = min(max(J - 1 + dy, 0), n1 - 1);
        const T* py = pz + (int64_t)jj * cs1;
        const Pair<T> mid = mg_ld2<T>(py + 2 * k);
        n.v[dy][0] = __ldg(py + xl);
        n.v[dy][1] = mid.a;
        n.v[dy][2] = mid.b;
        n.v[dy][3] = __ldg(py + xr);
    }
    // single-axis linear extrapolation of the out-of-range neighbours (2 u[clamp] - u[reflect])
    const bool x_lo = k == 0, x_hi = 2 * k + 2 > n2 - 1, y_lo = J == 0, y_hi = J == n1 - 1;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        n.v[dy][0] = x_lo ? T(2) * n.v[dy][1] - n.v[dy][2] : n.v[dy][0];
        n.v[dy][3] = x_hi ? T(2) * n.v[dy][2] - n.v[dy][1] : n.v[dy][3];
    }
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) {
        n.v[0][dx] = y_lo ? T(2) * n.v[1][dx] - n.v[2][dx] : n.v[0][dx];
        n.v[2][dx] = y_hi ? T(2) * n.v[1][dx] - n.v[0][dx] : n.v[2][dx];
    }
    return mg_reduce<T>(n);
}

template <typename T>
__device__ __forceinline__ MgP<T> mg_plane(const Mg3& m, const T* __restrict__ coarse, int coarse_z0, int zp, int J, int k) {
    if (zp >= 0 && zp <= m.n0 - 1) return mg_plane_inrange<T>(coarse, coarse_z0, m.cs0, m.cs1, m.n1, m.n2, zp, J, k);
    // padded plane -1 / n0: linear extrapolation along axis 0 of the in-plane results
    const MgP<T> Pa = mg_plane_inrange<T>(coarse, coarse_z0, m.cs0, m.cs1, m.n1, m.n2, zp < 0 ? 0 : m.n0 - 1, J, k);
    const MgP<T> Pb = mg_plane_inrange<T>(coarse, coarse_z0, m.cs0, m.cs1, m.n1, m.n2, zp < 0 ? 1 : m.n0 - 2, J, k);
    MgP<T> P;
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) P.v[b][c] = T(2) * Pa.v[b][c] - Pb.v[b][c];
    return P;
}

// Fine cells on the EDGES of the fine box (two or more axes at index 0 / last): the joint pad is not the product
// of the per-axis extrapolations there; recompute them with the generic per-cell routine.  One thread per cell:
// 4 per fine plane (z-edges) + the y- and x-edges of the first / last fine plane.
template <typename T>
__global__ void __launch_bounds__(128) k_interp_fix_edges(MgGeom g, Mg3 m, const T* __restrict__ coarse, T cfac,
                                                          const T* __restrict__ term, T ffac, T* __restrict__ out,
                                                          int fz_begin, int fz_end, int out_z0, int coarse_z0) {
    const int nf0 = 2 * m.n0, nf1 = 2 * m.n1, nf2 = 2 * m.n2;
    const int nz = fz_end - fz_begin;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int fz, fy, fx;
    if (t < 4 * nz) {
        fz = fz_begin + (t >> 2);
        fy = (t & 1) ? nf1 - 1 : 0;
        fx = (t & 2) ? nf2 - 1 : 0;
    } else if ((t -= 4 * nz) < 4 * nf1) {
        fz = (t & 1) ? nf0 - 1 : 0;
        fx = (t & 2) ? nf2 - 1 : 0;
        fy = t >> 2;
    } else if ((t -= 4 * nf1) < 4 * nf2) {
        fz = (t & 1) ? nf0 - 1 : 0;
        fy = (t & 2) ? nf1 - 1 : 0;
        fx = t >> 2;
    } else {
        return;
    }
    if (fz < fz_begin || fz >= fz_end) return;
    const int64_t f3[ODIL_B200_MAX_NDIM] = {fz, fy, fx, 0};
    const int64_t lin = (int64_t)(fz - out_z0) * m.fs0 + (int64_t)fy * m.fs1 + fx;
    T r = cfac * interp_cell_generic<T>(g, coarse, coarse_z0, f3);
    if (term) r += ffac * __ldg(term + lin);
    out[lin] = r;
}

// grid (ceil(n2 / 64), ceil(n1 / 4), z-chunks), block (32, 4): thread = coarse cells 2k, 2k+1 of coarse row J
template <typename T>
__global__ void __launch_bounds__(128) k_interp_add3m(Mg3 m, const T* __restrict__ coarse, T cfac,
                                                      const T* __restrict__ term, T ffac, T* __restrict__ out,
                                                      int fz_begin, int fz_end, int out_z0, int coarse_z0, int zc) {
    const int k = blockIdx.x * 32 + threadIdx.x;
    const int J = blockIdx.y * 4 + threadIdx.y;
    if (2 * k >= m.n2 || J >= m.n1) return;
    const int Ibeg = (fz_begin >> 1) + blockIdx.z * zc;
    const int Iend = min(Ibeg + zc, ((fz_end - 1) >> 1) + 1);
    if (Ibeg >= Iend) return;
    MgP<T> Pm = mg_plane<T>(m, coarse, coarse_z0, Ibeg - 1, J, k);
    MgP<T> Pc = mg_plane<T>(m, coarse, coarse_z0, Ibeg, J, k);
    const T s = cfac * T(1.0 / 64.0);
    const int64_t col = (int64_t)(2 * J) * m.fs1 + 4 * k;
    for (int I = Ibeg; I < Iend; ++I) {
        // the four fine vectors of this step first (independent of the coarse loads below)
        MgVec4<T> t[2][2];
        bool on[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int fz = 2 * I + a;
            on[a] = fz >= fz_begin && fz < fz_end;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int64_t lin = (int64_t)(fz - out_z0) * m.fs0 + col + (int64_t)b * m.fs1;
                t[a][b] = (term && on[a]) ? mg_ld4<T>(term + lin) : MgVec4<T>{T(0), T(0), T(0), T(0)};
            }
        }
        const MgP<T> Pp = mg_plane<T>(m, coarse, coarse_z0, I + 1, J, k);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int fz = 2 * I + a;
            if (on[a]) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int64_t lin = (int64_t)(fz - out_z0) * m.fs0 + col + (int64_t)b * m.fs1;
                    T r[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        r[c] = s * (a == 0 ? Pm.v[b][c] + T(3) * Pc.v[b][c] : T(3) * Pc.v[b][c] + Pp.v[b][c]);
                    r[0] = fma(ffac, t[a][b].x, r[0]);
                    r[1] = fma(ffac, t[a][b].y, r[1]);
                    r[2] = fma(ffac, t[a][b].z, r[2]);
                    r[3] = fma(ffac, t[a][b].w, r[3]);
                    mg_st4(out + lin, MgVec4<T>{r[0], r[1], r[2], r[3]});
                }
            }
        }
        Pm = Pc;
        Pc = Pp;
    }
}

// ------------------------------------------------------------------------------------------------
// Transposed interpolation.  Q(fz) = in-plane gather of the fine plane fz onto the coarse cells (J, 2k) and
// (J, 2k+1) with the separable 6-tap weights (interior [1,3,3,1]/4 + pad corrections); the coarse plane I is
// sum_t wz[t] * Q(2I-2+t).  A 6-plane window of Q slides along axis 0.
// grid (ceil(n2 / 64), ceil(n1 / 4), z-chunks), block (32, 4).
// ------------------------------------------------------------------------------------------------
template <typename T>
struct MgW6 {
    T w[6];
};
template <typename T>
__device__ __forceinline__ MgW6<T> mg_adjw(int J, int n) {
    MgW6<T> r;
    adjoint_weights<T>(J, n, r.w);
    return r;
}
template <typename T>
struct MgQ {
    T a, b;
};

// One warp-uniform variant per row class: BY = the coarse row is within 2 of a y face (6 fine rows with pad
// corrections, clipped rows carry zero weight) or interior (4 fine rows 2J-1 .. 2J+2, no clipping).
template <typename T, bool BY>
__device__ __forceinline__ void mg_adj_march(const Mg3& m, const T* __restrict__ gf, T scale, T* __restrict__ gc,
                                             int Ibeg, int Iend, int out_z0, int fine_z0, int J, int k, int lane) {
    constexpr int NROW = BY ? 6 : 4, R0 = BY ? 0 : 1;
    const int nf0 = 2 * m.n0, nf1 = 2 * m.n1, nf2 = 2 * m.n2;
    const bool valid = 2 * k < m.n2;
    const bool lane0 = lane == 0, lane31 = lane == 31;
    // neighbours outside the warp's span: lane 0 reads (4k-2, 4k-1), lane 31 reads (4k+4, 4k+5)
    const int ex = lane0 ? 4 * k - 2 : 4 * k + 4;
    const bool edge = valid && (lane0 || lane31) && ex >= 0 && ex + 1 < nf2;
    const MgW6<T> wx0 = mg_adjw<T>(valid ? 2 * k : 2, m.n2), wx1 = mg_adjw<T>(valid ? 2 * k + 1 : 3, m.n2);
    const MgW6<T> wy = mg_adjw<T>(J, m.n1);
    int roff[NROW];  // element offsets of the own vector in the rows 2J-2+R0 .. (clipped rows: zero weight)
#pragma unroll
    for (int i = 0; i < NROW; ++i) roff[i] = min(max(2 * J - 2 + R0 + i, 0), nf1 - 1) * (int)m.fs1;
    const int vcol = valid ? 4 * k : 0, ecol = edge ? ex : 0;
    // fine planes this chunk may touch (the outermost taps carry weight only next to the domain faces: fine plane 0
    // for I == 1, the last fine plane for I == n0-2; elsewhere they belong to the neighbouring chunk / slab)
    const int fz_lo = Ibeg == 1 ? 0 : max(2 * Ibeg - 1, 0);
    const int fz_hi = Iend - 1 == m.n0 - 2 ? nf0 - 1 : min(2 * Iend, nf0 - 1);

    MgQ<T> Q0{T(0), T(0)}, Q1 = Q0, Q2 = Q0, Q3 = Q0;
    for (int I = Ibeg - 2; I < Iend; ++I) {
        // in-plane gather of the fine planes 2I+2, 2I+3: ALL loads first (2 * NROW vectors in flight per thread)
        const int fa = 2 * I + 2, fb = 2 * I + 3;
        const bool va = fa >= fz_lo && fa <= fz_hi, vb = fb >= fz_lo && fb <= fz_hi;
        const T* pa = gf + (int64_t)(min(max(fa, fz_lo), fz_hi) - fine_z0) * m.fs0;
        const T* pb = gf + (int64_t)(min(max(fb, fz_lo), fz_hi) - fine_z0) * m.fs0;
        MgVec4<T> v[2][NROW];
        Pair<T> ev[2][NROW];
#pragma unroll
        for (int i = 0; i < NROW; ++i) {
            v[0][i] = mg_ld4<T>(pa + roff[i] + vcol);
            v[1][i] = mg_ld4<T>(pb + roff[i] + vcol);
        }
        if (lane0 || lane31) {
#pragma unroll
            for (int i = 0; i < NROW; ++i) {
                ev[0][i] = mg_ld2<T>(pa + roff[i] + ecol);
                ev[1][i] = mg_ld2<T>(pb + roff[i] + ecol);
            }
        }
        MgQ<T> Qn[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0), e0 = T(0), e1 = T(0);
#pragma unroll
            for (int i = 0; i < NROW; ++i) {
                const T w = wy.w[R0 + i];
                s0 = fma(w, v[p][i].x, s0);
                s1 = fma(w, v[p][i].y, s1);
                s2 = fma(w, v[p][i].z, s2);
                s3 = fma(w, v[p][i].w, s3);
            }
            if (lane0 || lane31) {
#pragma unroll
                for (int i = 0; i < NROW; ++i) {
                    const T w = wy.w[R0 + i];
                    e0 = fma(w, ev[p][i].a, e0);
                    e1 = fma(w, ev[p][i].b, e1);
                }
            }
            if (!valid) s0 = s1 = s2 = s3 = T(0);
            if (!edge) e0 = e1 = T(0);
            T lz = __shfl_up_sync(0xffffffffu, s2, 1), lw = __shfl_up_sync(0xffffffffu, s3, 1);
            T rx = __shfl_down_sync(0xffffffffu, s0, 1), ry = __shfl_down_sync(0xffffffffu, s1, 1);
            lz = lane0 ? e0 : lz;
            lw = lane0 ? e1 : lw;
            rx = lane31 ? e0 : rx;
            ry = lane31 ? e1 : ry;
            const T f[8] = {lz, lw, s0, s1, s2, s3, rx, ry};  // fine cells 4k-2 .. 4k+5
            T qa = T(0), qb = T(0);
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                qa = fma(wx0.w[t], f[t], qa);
                qb = fma(wx1.w[t], f[t + 2], qb);
            }
            const bool vp = p == 0 ? va : vb;
            Qn[p] = MgQ<T>{vp ? qa : T(0), vp ? qb : T(0)};
        }
        if (I >= Ibeg) {
            const MgW6<T> wz = mg_adjw<T>(I, m.n0);
            T a0 = wz.w[0] * Q0.a, a1 = wz.w[0] * Q0.b;
            a0 = fma(wz.w[1], Q1.a, a0);
            a1 = fma(wz.w[1], Q1.b, a1);
            a0 = fma(wz.w[2], Q2.a, a0);
            a1 = fma(wz.w[2], Q2.b, a1);
            a0 = fma(wz.w[3], Q3.a, a0);
            a1 = fma(wz.w[3], Q3.b, a1);
            a0 = fma(wz.w[4], Qn[0].a, a0);
            a1 = fma(wz.w[4], Qn[0].b, a1);
            a0 = fma(wz.w[5], Qn[1].a, a0);
            a1 = fma(wz.w[5], Qn[1].b, a1);
            // (cells with two or more axes within 2 of a face are rewritten by k_adjoint_fix_edges)
            if (valid) {
                const int64_t lin = (int64_t)(I - out_z0) * m.cs0 + (int64_t)J * m.cs1 + 2 * k;
                *reinterpret_cast<Pair<T>*>(gc + lin) = Pair<T>{scale * a0, scale * a1};
            }
        }
        Q0 = Q2;
        Q1 = Q3;
        Q2 = Qn[0];
        Q3 = Qn[1];
    }
}

template <typename T>
__global__ void __launch_bounds__(128, 4) k_interp_adjoint3m(Mg3 m, const T* __restrict__ gf, T scale,
                                                             T* __restrict__ gc, int cz_begin, int cz_end, int out_z0,
                                                             int fine_z0, int zc) {
    const int lane = threadIdx.x;
    const int k = blockIdx.x * 32 + lane;
    const int J = blockIdx.y * 4 + threadIdx.y;
    if (J >= m.n1) return;  // whole warps only: the lanes of a warp share J
    const int Ibeg = cz_begin + blockIdx.z * zc;
    const int Iend = min(Ibeg + zc, cz_end);
    if (Ibeg >= Iend) return;
    if (J <= 1 || J >= m.n1 - 2)
        mg_adj_march<T, true>(m, gf, scale, gc, Ibeg, Iend, out_z0, fine_z0, J, k, lane);
    else
        mg_adj_march<T, false>(m, gf, scale, gc, Ibeg, Iend, out_z0, fine_z0, J, k, lane);
}

// Coarse cells with two or more axes within 2 of a face: the joint pad is not separable there; recompute them
// with the generic transpose  (I^T g)[J] = sum_{q: clamp(q)=J} 2 G(q) - sum_{q: reflect(q)=J} G(q),  G(q) = the
// 4x4x4-tap gather of the fine gradient onto the padded coarse index q.  ONE WARP PER CELL: the (candidate q,
// tap) pairs are dealt to the lanes and summed in a fixed order (a single thread walking the up to 27 x 64
// taps took ~60 us, which was the whole cost of the coarse levels).  16 cells per coarse plane (z-edges) + the
// y- and x-edges of the four boundary planes.
template <typename T>
__global__ void __launch_bounds__(128) k_adjoint_fix_edges(Mg3 m, const T* __restrict__ gf, T scale,
                                                           T* __restrict__ gc, int cz_begin, int cz_end, int out_z0,
                                                           int fine_z0) {
    const int nz = cz_end - cz_begin;
    const int lane = threadIdx.x & 31;
    auto bnd = [](int i, int n) { return i < 2 ? i : n - 4 + i; };  // 0, 1, n-2, n-1
    int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;          // cell index = warp index
    int c[3];
    if (t < 16 * nz) {
        c[0] = cz_begin + (t >> 4);
        c[1] = bnd(t & 3, m.n1);
        c[2] = bnd((t >> 2) & 3, m.n2);
    } else if ((t -= 16 * nz) < 16 * m.n1) {
        c[0] = bnd(t & 3, m.n0);
        c[2] = bnd((t >> 2) & 3, m.n2);
        c[1] = t >> 4;
    } else if ((t -= 16 * m.n1) < 16 * m.n2) {
        c[0] = bnd(t & 3, m.n0);
        c[1] = bnd((t >> 2) & 3, m.n1);
        c[2] = t >> 4;
    } else {
        return;
    }
    if (c[0] < cz_begin || c[0] >= cz_end) return;  // warp-uniform
    const int n[3] = {m.n0, m.n1, m.n2};
    const int64_t fs[3] = {m.fs0, m.fs1, 1};
    // candidate padded indices per axis: the cell itself, -1 if it is one of the two lowest, n if one of the two highest
    int cand[3][3], nc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int kk = 0;
        cand[a][kk++] = c[a];
        if (c[a] <= 1) cand[a][kk++] = -1;
        if (c[a] >= n[a] - 2) cand[a][kk++] = n[a];
        nc[a] = kk;
    }
    const int ncombo = nc[0] * nc[1] * nc[2];
    T acc = T(0);
    for (int item = lane; item < ncombo * 64; item += 32) {
        const int combo = item >> 6, tap = item & 63;
        int q[3];
        int r = combo;
        q[2] = cand[2][r % nc[2]];
        r /= nc[2];
        q[1] = cand[1][r % nc[1]];
        r /= nc[1];
        q[0] = cand[0][r];
        bool mc = true, mr = true, outside = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int qc = q[a] < 0 ? 0 : (q[a] > n[a] - 1 ? n[a] - 1 : q[a]);
            const int qr = q[a] < 0 ? 1 : (q[a] > n[a] - 1 ? n[a] - 2 : q[a]);
            mc = mc && qc == c[a];
            mr = mr && qr == c[a];
            outside = outside || q[a] < 0 || q[a] > n[a] - 1;
        }
        // an in-range q is the plain value u[q]: coefficient 1 when q == J
        const T coef = outside ? T((mc ? 2 : 0) - (mr ? 1 : 0)) : T(1);
        if (coef == T(0)) continue;
        const int tt[3] = {tap >> 4, (tap >> 2) & 3, tap & 3};
        T w = coef;
        int64_t lin = 0;
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int fi = 2 * q[a] - 1 + tt[a];  // fine cells 2q-1 .. 2q+2, weights 1/4 3/4 3/4 1/4
            ok = ok && fi >= 0 && fi < 2 * n[a];
            w *= (tt[a] == 0 || tt[a] == 3) ? T(0.25) : T(0.75);
            lin += (int64_t)(a == 0 ? fi - fine_z0 : fi) * fs[a];
        }
        if (ok) acc = fma(w, __ldg(gf + lin), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) gc[(int64_t)(c[0] - out_z0) * m.cs0 + (int64_t)c[1] * m.cs1 + c[2]] = scale * acc;
}

}  // namespace odil
