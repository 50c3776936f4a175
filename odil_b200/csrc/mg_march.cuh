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
// Boundary rule: the joint pad P = 2*symmetric(u) - reflect(u) (core.py:640-643), exact in both kernels: the
// interpolation loads u[clamp(q)] and, predicated, u[reflect(q)]; its transpose uses the separable 6-tap weights
// (exact with one axis out of range) plus a closed-form correction for the cells near an edge or a corner.
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

// In-plane result of the padded coarse plane zp in [-1, n0] with the joint pad applied literally: a neighbour q with
// any coordinate out of range is 2*u[clamp(q)] - u[reflect(q)] (clamp / reflect on ALL axes at once, core.py:640-643).
// Exact for every thread and plane; used where two or more axes can leave the range at once.
template <typename T>
__device__ __noinline__ MgP<T> mg_plane_joint(Mg3 m, const T* __restrict__ coarse, int coarse_z0, int zp, int J, int k) {
    const int zc = min(max(zp, 0), m.n0 - 1), zr = zp < 0 ? 1 : (zp > m.n0 - 1 ? m.n0 - 2 : zp);
    const bool oz = zp != zc;
    const T* pzc = coarse + (int64_t)(zc - coarse_z0) * m.cs0;
    const T* pzr = coarse + (int64_t)(zr - coarse_z0) * m.cs0;
    const int xl = 2 * k - 1, xr = 2 * k + 2;
    const int xlc = max(xl, 0), xlr = xl < 0 ? 1 : xl;
    const int xrc = min(xr, m.n2 - 1), xrr = xr > m.n2 - 1 ? m.n2 - 2 : xr;
    const bool oxl = xl < 0, oxr = xr > m.n2 - 1;
    MgNb<T> n;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yq = J - 1 + dy;
        const int yc = min(max(yq, 0), m.n1 - 1), yr = yq < 0 ? 1 : (yq > m.n1 - 1 ? m.n1 - 2 : yq);
        const bool ozy = oz || yq != yc;
        const T* pa = pzc + (int64_t)yc * m.cs1;
        const T* pb = pzr + (int64_t)yr * m.cs1;
        const Pair<T> mid = mg_ld2<T>(pa + 2 * k);
        T v0 = __ldg(pa + xlc), v3 = __ldg(pa + xrc);
        T v1 = mid.a, v2 = mid.b;
        if (ozy) {  // the whole row of neighbours is out of range along z and / or y
            const Pair<T> midb = mg_ld2<T>(pb + 2 * k);
            v1 = T(2) * v1 - midb.a;
            v2 = T(2) * v2 - midb.b;
        }
        if (ozy || oxl) v0 = T(2) * v0 - __ldg(pb + xlr);
        if (ozy || oxr) v3 = T(2) * v3 - __ldg(pb + xrr);
        n.v[dy][0] = v0;
        n.v[dy][1] = v1;
        n.v[dy][2] = v2;
        n.v[dy][3] = v3;
    }
    return mg_reduce<T>(n);
}

// Plane inside the array with at most ONE axis of each neighbour out of range: the joint pad is then the linear
// extrapolation along that axis, done with selects on values that are loaded anyway (no second load, no branch).
template <typename T>
__device__ __forceinline__ MgP<T> mg_plane_inrange(const Mg3& m, const T* __restrict__ coarse, int coarse_z0, int zp, int J,
                                                   int k) {
    MgNb<T> n;
    const T* pz = coarse + (int64_t)(zp - coarse_z0) * m.cs0;
    const int xl = max(2 * k - 1, 0), xr = min(2 * k + 2, m.n2 - 1);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int jj = min(max(J - 1 + dy, 0), m.n1 - 1);
        const T* py = pz + (int64_t)jj * m.cs1;
        const Pair<T> mid = mg_ld2<T>(py + 2 * k);
        n.v[dy][0] = __ldg(py + xl);
        n.v[dy][1] = mid.a;
        n.v[dy][2] = mid.b;
        n.v[dy][3] = __ldg(py + xr);
    }
    const bool x_lo = k == 0, x_hi = 2 * k + 2 > m.n2 - 1, y_lo = J == 0, y_hi = J == m.n1 - 1;
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

// Dispatcher: the out-of-line joint loader only where two or more axes can leave the range at once (the four corner
// columns of every plane, the boundary rows / columns of the two padded planes -1 and n0).
template <typename T>
__device__ __forceinline__ MgP<T> mg_plane(const Mg3& m, const T* __restrict__ coarse, int coarse_z0, int zp, int J, int k) {
    const bool bx = k == 0 || 2 * k + 2 > m.n2 - 1, by = J == 0 || J == m.n1 - 1;
    const bool zout = zp < 0 || zp > m.n0 - 1;
    if (zout ? (bx || by) : (bx && by)) return mg_plane_joint<T>(m, coarse, coarse_z0, zp, J, k);
    if (!zout) return mg_plane_inrange<T>(m, coarse, coarse_z0, zp, J, k);
    const MgP<T> Pa = mg_plane_inrange<T>(m, coarse, coarse_z0, zp < 0 ? 0 : m.n0 - 1, J, k);
    const MgP<T> Pb = mg_plane_inrange<T>(m, coarse, coarse_z0, zp < 0 ? 1 : m.n0 - 2, J, k);
    MgP<T> P;
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) P.v[b][c] = T(2) * Pa.v[b][c] - Pb.v[b][c];
    return P;
}

// grid (ceil(n2 / 64), ceil(n1 / 4), z-chunks), block (32, 4): thread = coarse cells 2k, 2k+1 of coarse row J
template <typename T, bool PF = false>
__global__ void __launch_bounds__(128, PF ? 4 : 5) k_interp_add3m(Mg3 m, const T* __restrict__ coarse, T cfac,
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
    // the four fine vectors of a step are loaded one step AHEAD (PF: twice the bytes in flight per warp; the kernel
    // waits on memory -- ncu: 10 long-scoreboard stall cycles per issued instruction, 35 % of the issue slots used)
    auto load_terms = [&](int I, MgVec4<T> (&t)[2][2]) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int fz = 2 * I + a;
            const bool on = I < Iend && fz >= fz_begin && fz < fz_end;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int64_t lin = (int64_t)(fz - out_z0) * m.fs0 + col + (int64_t)b * m.fs1;
                t[a][b] = (term && on) ? mg_ld4<T>(term + lin) : MgVec4<T>{T(0), T(0), T(0), T(0)};
            }
        }
    };
    MgVec4<T> tn[2][2];
    if (PF) load_terms(Ibeg, tn);
    for (int I = Ibeg; I < Iend; ++I) {
        MgVec4<T> t[2][2];
        bool on[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int fz = 2 * I + a;
            on[a] = fz >= fz_begin && fz < fz_end;
        }
        if (PF) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) t[a][b] = tn[a][b];
            load_terms(I + 1, tn);
        } else {
            load_terms(I, t);
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

// Joint-pad correction of the transposed interpolation.  The separable 6-tap weights treat the pad as successive
// per-axis linear extrapolations, prod_a (2 C_a - R_a) (C = clamp, R = reflect); the reference pads jointly,
// 2 prod_a C_a - prod_a R_a (core.py:640-643).  They agree when one axis leaves the range; for a padded index q
// with the out-of-range axis set S, |S| >= 2, the difference transposes to
//     gc[J] += coef(pattern of J) * G(q),   G(q) = gather of the fine gradient onto q,
// with coef = -2, +2, +2, -2 for the patterns CC, CR, RC, RR (|S| = 2) and -6 (CCC), +4 (one R), -2 (two R), 0 (RRR)
// (|S| = 3).  Along an out-of-range axis the gather has the single tap 1/4 on the first / last fine cell, along the
// others the 4 taps (1,3,3,1)/4 on the fine cells 2q-1 .. 2q+2.  Only coarse cells with two or more coordinates
// within 2 of a face are affected; the cost is at most 13 loads for such a cell.
template <typename T>
__device__ __forceinline__ T mg_adj_joint_correction(const Mg3& m, const T* __restrict__ gf, int fine_z0, int I, int J, int K) {
    const int c[3] = {I, J, K};
    const int n[3] = {m.n0, m.n1, m.n2};
    const int64_t fs[3] = {m.fs0, m.fs1, 1};
    int side[3], pat[3];  // side: -1 low, +1 high, 0 not near a face; pat: 0 = clamp target, 1 = reflect target
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        side[a] = c[a] <= 1 ? -1 : (c[a] >= n[a] - 2 ? 1 : 0);
        pat[a] = (c[a] == 0 || c[a] == n[a] - 1) ? 0 : 1;
    }
    T acc = T(0);
    // subsets S of the axes near a face, |S| >= 2: masks 3, 5, 6, 7
#pragma unroll
    for (int mask = 3; mask <= 7; ++mask) {
        if (mask == 4) continue;
        bool ok = true;
        int nr = 0, ns = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            if (mask >> a & 1) {
                ok = ok && side[a] != 0;
                nr += pat[a];
                ++ns;
            }
        if (!ok) continue;
        const T coef = ns == 2 ? (nr == 1 ? T(2) : T(-2)) : (nr == 0 ? T(-6) : (nr == 1 ? T(4) : (nr == 2 ? T(-2) : T(0))));
        if (coef == T(0)) continue;
        // G(q): q_a = -1 / n for a in S (single tap), q_a = c_a otherwise (4 taps)
        int64_t base = 0;
        T w0 = coef;
        int free_axis = -1;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (mask >> a & 1) {
                const int fi = side[a] < 0 ? 0 : 2 * n[a] - 1;
                base += (int64_t)(a == 0 ? fi - fine_z0 : fi) * fs[a];
                w0 *= T(0.25);
            } else {
                free_axis = a;
            }
        }
        if (free_axis < 0) {
            acc = fma(w0, __ldg(gf + base), acc);
        } else {
            const int a = free_axis;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int fi = 2 * c[a] - 1 + t;
                if (fi >= 0 && fi < 2 * n[a]) {
                    const T w = (t == 0 || t == 3) ? T(0.25) : T(0.75);
                    acc = fma(w0 * w, __ldg(gf + base + (int64_t)(a == 0 ? fi - fine_z0 : fi) * fs[a]), acc);
                }
            }
        }
    }
    return acc;
}

// Adam state of the FINE array whose gradient the transposed interpolation streams (the finest multigrid term: its
// gradient IS g_fine).  With ADAM = true the thread that owns a fine vector -- rows 2J, 2J+1 of the planes 2I, 2I+1 of
// its coarse cells -- also applies the Adam update to it, so the gradient is read from HBM once instead of twice
// (once here, once by k_adam): 28.5 instead of 32.5 bytes per fine cell for the pair of kernels.
template <typename T>
struct MgAdam {
    T* x;
    T* m;
    T* v;
    T alpha, omb1, omb2, eps;
    const double* alpha_dev;
};

template <typename T>
__device__ __forceinline__ void mg_adam4(const MgAdam<T>& ad, T alpha, unsigned idx, const MgVec4<T>& g) {
    MgVec4<T> xx = mg_ld4<T>(ad.x + idx), mm = mg_ld4<T>(ad.m + idx), vv = mg_ld4<T>(ad.v + idx);
    adam_one(xx.x, mm.x, vv.x, g.x, alpha, ad.omb1, ad.omb2, ad.eps);
    adam_one(xx.y, mm.y, vv.y, g.y, alpha, ad.omb1, ad.omb2, ad.eps);
    adam_one(xx.z, mm.z, vv.z, g.z, alpha, ad.omb1, ad.omb2, ad.eps);
    adam_one(xx.w, mm.w, vv.w, g.w, alpha, ad.omb1, ad.omb2, ad.eps);
    mg_st4(ad.x + idx, xx);
    mg_st4(ad.m + idx, mm);
    mg_st4(ad.v + idx, vv);
}

// One warp-uniform variant per row class: BY = the coarse row is within 2 of a y face (6 fine rows with pad
// corrections, clipped rows carry zero weight) or interior (4 fine rows 2J-1 .. 2J+2, no clipping).
template <typename T, bool BY, bool ADAM = false>
__device__ __forceinline__ void mg_adj_march(const Mg3& m, const T* __restrict__ gf, T scale, T* __restrict__ gc,
                                             int Ibeg, int Iend, int out_z0, int fine_z0, int J, int k, int lane,
                                             const MgAdam<T>* ad = nullptr) {
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
        // 32-bit element offsets (the host routes arrays of 2^31 or more elements to the older kernels): one
        // IMAD.WIDE per load instead of a 64-bit multiply-add chain
        const unsigned oa = (unsigned)(min(max(fa, fz_lo), fz_hi) - fine_z0) * (unsigned)m.fs0;
        const unsigned ob = (unsigned)(min(max(fb, fz_lo), fz_hi) - fine_z0) * (unsigned)m.fs0;
        MgVec4<T> v[2][NROW];
        Pair<T> ev[2][NROW];
#pragma unroll
        for (int i = 0; i < NROW; ++i) {
            v[0][i] = mg_ld4<T>(gf + (oa + (unsigned)(roff[i] + vcol)));
            v[1][i] = mg_ld4<T>(gf + (ob + (unsigned)(roff[i] + vcol)));
        }
        if (lane0 || lane31) {
#pragma unroll
            for (int i = 0; i < NROW; ++i) {
                ev[0][i] = mg_ld2<T>(gf + (oa + (unsigned)(roff[i] + ecol)));
                ev[1][i] = mg_ld2<T>(gf + (ob + (unsigned)(roff[i] + ecol)));
            }
        }
        if constexpr (ADAM) {
            // planes 2I+2, 2I+3 are the own planes of coarse plane I+1; rows 2J, 2J+1 are the own rows
            if (valid && I + 1 >= Ibeg && I + 1 < Iend) {
                constexpr int OWN = 2 - R0;
                const T alpha = ad->alpha_dev ? (T)__ldg(ad->alpha_dev) : ad->alpha;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    mg_adam4<T>(*ad, alpha, oa + (unsigned)(roff[OWN + r] + vcol), v[0][OWN + r]);
                    mg_adam4<T>(*ad, alpha, ob + (unsigned)(roff[OWN + r] + vcol), v[1][OWN + r]);
                }
            }
        }
        MgQ<T> Qn[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0), e0 = T(0), e1 = T(0);
#pragma unroll
            for (int i = 0; i < NROW; ++i) {
                // interior rows: the four taps are the literal constants 1/4 3/4 3/4 1/4 (immediate-operand FFMA)
                const T w = BY ? wy.w[R0 + i] : ((i == 0 || i == 3) ? T(0.25) : T(0.75));
                s0 = fma(w, v[p][i].x, s0);
                s1 = fma(w, v[p][i].y, s1);
                s2 = fma(w, v[p][i].z, s2);
                s3 = fma(w, v[p][i].w, s3);
            }
            if (lane0 || lane31) {
#pragma unroll
                for (int i = 0; i < NROW; ++i) {
                    const T w = BY ? wy.w[R0 + i] : ((i == 0 || i == 3) ? T(0.25) : T(0.75));
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
            T a0, a1;
            if (I >= 2 && I <= m.n0 - 3) {  // interior coarse plane: taps 1/4 3/4 3/4 1/4 on the planes 2I-1 .. 2I+2
                a0 = T(0.25) * (Q1.a + Qn[0].a) + T(0.75) * (Q2.a + Q3.a);
                a1 = T(0.25) * (Q1.b + Qn[0].b) + T(0.75) * (Q2.b + Q3.b);
            } else {
                const MgW6<T> wz = mg_adjw<T>(I, m.n0);
                a0 = wz.w[0] * Q0.a;
                a1 = wz.w[0] * Q0.b;
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
            }
            // (cells with two or more coordinates within 2 of a face get the joint-pad correction from
            //  k_adjoint_joint_fix, launched right after this kernel)
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

template <typename T, int OCC = 4>
__global__ void __launch_bounds__(128, OCC) k_interp_adjoint3m(Mg3 m, const T* __restrict__ gf, T scale,
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

// Same sweep + the Adam update of the fine array (whole array, single GPU: fine_z0 = out_z0 = 0).
template <typename T>
__global__ void __launch_bounds__(128, 3) k_interp_adjoint3m_adam(Mg3 m, const T* __restrict__ gf, T scale,
                                                                  T* __restrict__ gc, int cz_begin, int cz_end, int zc,
                                                                  const MgAdam<T> ad) {
    const int lane = threadIdx.x;
    const int k = blockIdx.x * 32 + lane;
    const int J = blockIdx.y * 4 + threadIdx.y;
    if (J >= m.n1) return;
    const int Ibeg = cz_begin + blockIdx.z * zc;
    const int Iend = min(Ibeg + zc, cz_end);
    if (Ibeg >= Iend) return;
    if (J <= 1 || J >= m.n1 - 2)
        mg_adj_march<T, true, true>(m, gf, scale, gc, Ibeg, Iend, 0, 0, J, k, lane, &ad);
    else
        mg_adj_march<T, false, true>(m, gf, scale, gc, Ibeg, Iend, 0, 0, J, k, lane, &ad);
}

// gc += scale * (joint-pad correction) on the coarse cells with two or more coordinates within 2 of a face: 16 per
// coarse plane (z-edges) + the y- and x-edges of the four boundary planes; at most 13 loads per cell.  Cells that
// belong to two families (the corners) are handled by the first family only.
template <typename T>
__global__ void __launch_bounds__(128) k_adjoint_joint_fix(Mg3 m, const T* __restrict__ gf, T scale, T* __restrict__ gc,
                                                           int cz_begin, int cz_end, int out_z0, int fine_z0) {
    const int nz = cz_end - cz_begin;
    auto bnd = [](int i, int n) { return i < 2 ? i : n - 4 + i; };  // 0, 1, n-2, n-1
    auto near = [](int i, int n) { return i <= 1 || i >= n - 2; };
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int I, J, K;
    if (t < 16 * nz) {  // y and x near a face, every plane
        I = cz_begin + (t >> 4);
        J = bnd(t & 3, m.n1);
        K = bnd((t >> 2) & 3, m.n2);
    } else if ((t -= 16 * nz) < 16 * m.n1) {  // z and x near a face, rows not near a y face
        I = bnd(t & 3, m.n0);
        K = bnd((t >> 2) & 3, m.n2);
        J = t >> 4;
        if (near(J, m.n1)) return;
    } else if ((t -= 16 * m.n1) < 16 * m.n2) {  // z and y near a face, columns not near an x face
        I = bnd(t & 3, m.n0);
        J = bnd((t >> 2) & 3, m.n1);
        K = t >> 4;
        if (near(K, m.n2)) return;
    } else {
        return;
    }
    if (I < cz_begin || I >= cz_end) return;
    const int64_t lin = (int64_t)(I - out_z0) * m.cs0 + (int64_t)J * m.cs1 + K;
    gc[lin] += scale * mg_adj_joint_correction<T>(m, gf, fine_z0, I, J, K);
}

}  // namespace odil
