// k_star7 -- the hot kernel of the library: fused residual + loss + adjoint gradient of a 3-D (or 2-D)
// star stencil in ONE sweep, written for a low instruction count per cell (the previous TMA kernel
// issued 116 thread-instructions per cell and was issue-bound at 30 % of the HBM roofline).
//
// Replaces on the reference side (one launch): ctx.field()=roll (core.py:910-975), the operator's
// arithmetic incl. where(index-mask) boundary rows (examples/poisson/poisson.py:57-68,100-113), the loss
// reduction (core.py:1093) and the reverse-mode gradient (core.py:1100-1101).
//
// Work decomposition (CTA tile = TY rows x TX = 32*VW columns, marching along axis 0 over a z-chunk)
//   strip warp w  tile rows 2w, 2w+1; a lane owns VW consecutive cells (one 16-byte vector) of both rows.
//                 Own columns of U[k-1..k+1] and F[k-2..k] stay in registers (rotated by a 3-way unrolled
//                 loop, no moves); x-neighbours come from lane shuffles, y-neighbours of the strip from
//                 shared memory.  Computes F, the loss partial and g for its rows.
//   y-ring warp   F (only) of the two rows just outside the tile, y = ty0-1 and y = ty0+TY.
//   x-ring warp   F (only) of the two columns x = tx0-1 and x = tx0+TX, one lane per cell; lane 0 is
//                 also the TMA producer.
//   TMA           U planes (TY+4 rows, TX+2VW columns, zero fill outside the array) into a 4-slot ring,
//                 c planes (TY+2 rows) into a 3-slot ring, up to three stages in flight, one mbarrier
//                 per stage.  F[k] is handed to the neighbouring warps through a 3-slot ring and an
//                 mbarrier with one arrival per warp and plane: a warp arrives right after storing F[k]
//                 and waits only just before storing F[k+1], so nobody idles at a CTA-wide barrier.
// Coefficients
//   Every lane keeps the 7 coefficients of each of ITS OWN cells in registers (class along x resolved
//   per cell at start-up, interior class along y and z), so x-boundary columns cost nothing.  Planes and
//   rows whose y or z class is not interior (2r of N) take the `step` path with out-of-line per-cell
//   table lookups; everything else runs `lean`, which has no flags, no calls and no address arithmetic
//   beyond the ring slots.
//   F is forced to exactly 0 on every cell outside the global domain, so out-of-domain sources never
//   contribute to the adjoint whatever coefficient they meet.  Requires a wrap-free plan.
#pragma once

namespace odil {

template <typename T, int VW>
struct alignas(16) Pack {
    T v[VW];
};

template <typename T, int VW, int TY>
struct Star7Cfg {
    static_assert(TY % 2 == 0 && TY <= 16, "TY must be even and <= 16 (one x-ring warp)");
    static_assert(sizeof(T) * VW == 16, "a lane owns one 16-byte vector");
    static constexpr int TX = 32 * VW;
    static constexpr int BX = TX + 2 * VW;
    static constexpr int NWS = TY / 2;         // strip warps
    static constexpr int NT = 32 * (NWS + 2);  // + y-ring warp + x-ring warp
    static constexpr int RU = TY + 4, RC = TY + 2;
    static constexpr int BXB = BX * (int)sizeof(T);                // bytes per staged row
    static constexpr int SLOT_U = ((RU * BXB + 127) / 128) * 128;  // bytes
    static constexpr int SLOT_C = ((RC * BXB + 127) / 128) * 128;
    static constexpr int SLOT_F = RC * BXB;
    static constexpr int NSU = 4, NSC = 3, NSF = 3;
    static constexpr int TAB = 1024;  // coefficient-table elements cached in shared memory
    static constexpr int OFF_U = 128, OFF_C = OFF_U + NSU * SLOT_U, OFF_F = OFF_C + NSC * SLOT_C;
    static constexpr int OFF_TAB = OFF_F + NSF * SLOT_F;
    static constexpr uint32_t BYTES_U = RU * BXB;
    static constexpr uint32_t BYTES_C = RC * BXB;
    static constexpr size_t SMEM = OFF_TAB + sizeof(T) * TAB + 64;
    // CTAs per SM the register budget is sized for (general / x-uniform coefficient registers)
    static constexpr int MINB = NT <= 256 ? 2 : 1;
    static constexpr int MINB_XU = NT <= 192 ? 3 : 2;
    static constexpr int regs_for(int ctas) { return ((65536 / (NT * ctas)) / 8) * 8 > 255 ? 255 : ((65536 / (NT * ctas)) / 8) * 8; }
    static constexpr int MAXREG = regs_for(MINB), MAXREG_XU = regs_for(MINB_XU);
};

template <typename T>
struct Star7Params {
    T* G;
    T* Fout;
    double* partials;
    const T* table;  // [ncls][7]: c, zm, zp, ym, yp, xm, xp
    int n0, N0g, z0, halo;
    int N1, N2;
    int R0, R1, R2;
    T scale;
    int zchunk;
    int has_c;
};

__device__ __forceinline__ int s7_cls(int i, int n, int r) {  // class of in-range index i
    if (i < r) return i;
    const int d = n - 1 - i;
    return d < r ? 2 * r - d : r;
}
__device__ __forceinline__ int s7_clamp(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

__device__ __forceinline__ void s7_bar_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// Shared-memory access through 32-bit shared addresses (no generic-pointer conversion in the hot loop).
__device__ __forceinline__ Pack<float, 4> s7_lds(uint32_t a, Pack<float, 4>*) {
    Pack<float, 4> r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(a));
    return r;
}
__device__ __forceinline__ Pack<double, 2> s7_lds(uint32_t a, Pack<double, 2>*) {
    Pack<double, 2> r;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "r"(a));
    return r;
}
__device__ __forceinline__ void s7_sts(uint32_t a, const Pack<float, 4>& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.v[0]), "f"(v.v[1]), "f"(v.v[2]), "f"(v.v[3])
                 : "memory");
}
__device__ __forceinline__ void s7_sts(uint32_t a, const Pack<double, 2>& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v.v[0]), "d"(v.v[1]) : "memory");
}
__device__ __forceinline__ float s7_lds1(uint32_t a, float*) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ double s7_lds1(uint32_t a, double*) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void s7_sts1(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void s7_sts1(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// predicated scalar load: only the lanes with `pred` touch shared memory (one wavefront), others get 0
__device__ __forceinline__ float s7_lds1_if(uint32_t a, bool pred, float*) {
    float r;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@p ld.shared.f32 %0, [%1];\n\t}"
        : "=f"(r)
        : "r"(a), "r"((int)pred));
    return r;
}
__device__ __forceinline__ double s7_lds1_if(uint32_t a, bool pred, double*) {
    double r;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.shared.f64 %0, [%1];\n\t}"
        : "=d"(r)
        : "r"(a), "r"((int)pred));
    return r;
}
__device__ __forceinline__ void s7_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // a lost TMA must fail loudly, never hang the device
    } while (!ok);
}

__device__ __forceinline__ void s7_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Coefficients of the own cells of a lane (interior y/z class).  The centre and x-arm coefficients are kept
// per cell (the class along x differs from cell to cell at the domain faces); the z- and y-arm coefficients
// are per cell too in general (XU = false) or, when the plan says they do not depend on the x class
// (XU = true: Poisson and every operator whose boundary rows only touch their own axis), one value each.
template <typename T, int VW, bool XU>
struct S7W;
template <typename T, int VW>
struct S7W<T, VW, false> {
    Pack<T, VW> c, xm, xp, a[4];  // a: zm, zp, ym, yp
    __device__ __forceinline__ T arm(int o, int j) const { return a[o].v[j]; }
    __device__ __forceinline__ void set_arm(int o, int j, T v) { a[o].v[j] = v; }
};
template <typename T, int VW>
struct S7W<T, VW, true> {
    Pack<T, VW> c, xm, xp;
    T a[4];
    __device__ __forceinline__ T arm(int o, int) const { return a[o]; }
    __device__ __forceinline__ void set_arm(int o, int j, T v) {
        if (j == 0) a[o] = v;
    }
};

// Packed fp32 arithmetic (sm_100: FFMA2 / FMUL2, two IEEE fp32 operations per instruction and lane).  The sweep is
// bound by instruction issue (ncu: 55 % issue slots at 4 warps per scheduler, FMA pipe 23 %): pairing the four cells of
// a lane halves the issue slots of the multiply-adds.  Every cell keeps its own chain of operations in the same order,
// so the results are bit-identical to the scalar form.
#ifndef ODIL_B200_NO_FFMA2
#define ODIL_B200_FFMA2 1
#else
#define ODIL_B200_FFMA2 0
#endif

__device__ __forceinline__ float2 s7_arm2(const S7W<float, 4, true>& w, int o, int) { return make_float2(w.a[o], w.a[o]); }
__device__ __forceinline__ float2 s7_arm2(const S7W<float, 4, false>& w, int o, int h) {
    return make_float2(w.a[o].v[2 * h], w.a[o].v[2 * h + 1]);
}
__device__ __forceinline__ float2 s7_pair(const Pack<float, 4>& a, int h) { return make_float2(a.v[2 * h], a.v[2 * h + 1]); }

template <bool XU>
__device__ __forceinline__ Pack<float, 4> s7_fwd_pk(const S7W<float, 4, XU>& w, const Pack<float, 4>& cc,
                                                    const Pack<float, 4>& uc, const Pack<float, 4>& um,
                                                    const Pack<float, 4>& up, const Pack<float, 4>& uym,
                                                    const Pack<float, 4>& uyp, float ul, float ur) {
    // aligned pairs (cells 0,1 and 2,3) for the centre and the z / y arms; the x arms read the neighbouring cell, whose
    // pairs would have to be assembled with moves, so they stay scalar (same position in each cell's chain)
    Pack<float, 4> f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float2 a = s7_pair(cc, h);
        a = __ffma2_rn(s7_pair(w.c, h), s7_pair(uc, h), a);
        a = __ffma2_rn(s7_arm2(w, 0, h), s7_pair(um, h), a);
        a = __ffma2_rn(s7_arm2(w, 1, h), s7_pair(up, h), a);
        a = __ffma2_rn(s7_arm2(w, 2, h), s7_pair(uym, h), a);
        a = __ffma2_rn(s7_arm2(w, 3, h), s7_pair(uyp, h), a);
        f.v[2 * h] = a.x;
        f.v[2 * h + 1] = a.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float xl = j > 0 ? uc.v[j > 0 ? j - 1 : 0] : ul;
        const float xr = j < 3 ? uc.v[j < 3 ? j + 1 : 0] : ur;
        f.v[j] = fmaf(w.xm.v[j], xl, f.v[j]);
        f.v[j] = fmaf(w.xp.v[j], xr, f.v[j]);
    }
    return f;
}

// F of VW consecutive cells of one row.
template <typename T, int VW, bool XU>
__device__ __forceinline__ Pack<T, VW> s7_fwd(const S7W<T, VW, XU>& w, const Pack<T, VW>& cc, const Pack<T, VW>& uc,
                                              const Pack<T, VW>& um, const Pack<T, VW>& up, const Pack<T, VW>& uym,
                                              const Pack<T, VW>& uyp, T ul, T ur) {
#if ODIL_B200_FFMA2
    if constexpr (sizeof(T) == 4 && VW == 4) return s7_fwd_pk<XU>(w, cc, uc, um, up, uym, uyp, ul, ur);
#endif
    Pack<T, VW> f;
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const T xl = j > 0 ? uc.v[j > 0 ? j - 1 : 0] : ul;
        const T xr = j < VW - 1 ? uc.v[j < VW - 1 ? j + 1 : 0] : ur;
        T a = cc.v[j];
        a = fma(w.c.v[j], uc.v[j], a);
        a = fma(w.arm(0, j), um.v[j], a);
        a = fma(w.arm(1, j), up.v[j], a);
        a = fma(w.arm(2, j), uym.v[j], a);
        a = fma(w.arm(3, j), uyp.v[j], a);
        a = fma(w.xm.v[j], xl, a);
        a = fma(w.xp.v[j], xr, a);
        f.v[j] = a;
    }
    return f;
}

// g of VW consecutive cells from the coefficients of the SOURCE cells: for the z / y arms the source has
// the same x (own coefficient registers, interior y/z class); for the x arms it is the neighbouring cell
// (wxpL / wxmR for the cells outside the lane's vector).
template <typename T, int VW, bool XU>
__device__ __forceinline__ Pack<T, VW> s7_adj(const S7W<T, VW, XU>& w, T wxpL, T wxmR, T scale, const Pack<T, VW>& fc,
                                              const Pack<T, VW>& fm, const Pack<T, VW>& fp, const Pack<T, VW>& fym,
                                              const Pack<T, VW>& fyp, T fl, T fr) {
    Pack<T, VW> g;
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const T xl = j > 0 ? fc.v[j > 0 ? j - 1 : 0] : fl;
        const T xr = j < VW - 1 ? fc.v[j < VW - 1 ? j + 1 : 0] : fr;
        const T axm = j < VW - 1 ? w.xm.v[j < VW - 1 ? j + 1 : 0] : wxmR;  // xm row of the cell at x+1
        const T axp = j > 0 ? w.xp.v[j > 0 ? j - 1 : 0] : wxpL;            // xp row of the cell at x-1
        T s = w.c.v[j] * fc.v[j];
        s = fma(w.arm(0, j), fp.v[j], s);   // zm row of the cell in plane k+1
        s = fma(w.arm(1, j), fm.v[j], s);   // zp row of the cell in plane k-1
        s = fma(w.arm(2, j), fyp.v[j], s);  // ym row of the cell at y+1
        s = fma(w.arm(3, j), fym.v[j], s);  // yp row of the cell at y-1
        s = fma(axm, xr, s);
        s = fma(axp, xl, s);
        g.v[j] = s * scale;
    }
    return g;
}

// Loads the coefficients of the cells x0 .. x0+VW-1 for the class (cz, cy) from the table; cells outside
// the domain get zeros (their F is then exactly 0: c and U are zero-filled there).
template <typename T, int VW, bool XU>
__device__ __forceinline__ void s7_load_w(S7W<T, VW, XU>& w, const T* __restrict__ tab, int rowbase, int x0, int N2, int R2) {
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const int x = x0 + j;
        const bool in = x < N2;
        const T* row = tab + (rowbase + s7_cls(s7_clamp(x, N2), N2, R2)) * 7;
        w.c.v[j] = in ? row[0] : T(0);
        w.xm.v[j] = in ? row[5] : T(0);
        w.xp.v[j] = in ? row[6] : T(0);
#pragma unroll
        for (int o = 0; o < 4; ++o) w.set_arm(o, j, (in || XU) ? row[1 + o] : T(0));
    }
}

// Out-of-line boundary paths (rows / planes whose y or z class is not interior).
struct S7Geom {
    int N0g, N1, N2, R0, R1, R2;
};

template <typename T, int VW>
__device__ __noinline__ Pack<T, VW> s7_slow_fwd(const T* __restrict__ tab, S7Geom gm, int zg, int y, int x0,
                                                Pack<T, VW> cc, Pack<T, VW> uc, Pack<T, VW> um, Pack<T, VW> up,
                                                Pack<T, VW> uym, Pack<T, VW> uyp, T ul, T ur) {
    const int C1 = 2 * gm.R1 + 1, C2 = 2 * gm.R2 + 1;
    const int base = (s7_cls(zg, gm.N0g, gm.R0) * C1 + s7_cls(y, gm.N1, gm.R1)) * C2;
    Pack<T, VW> f;
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const int x = x0 + j;
        T a = T(0);
        if (x < gm.N2) {
            const T* w = tab + (base + s7_cls(x, gm.N2, gm.R2)) * 7;
            const T xl = j > 0 ? uc.v[j > 0 ? j - 1 : 0] : ul;
            const T xr = j < VW - 1 ? uc.v[j < VW - 1 ? j + 1 : 0] : ur;
            a = cc.v[j] + w[0] * uc.v[j] + w[1] * um.v[j] + w[2] * up.v[j] + w[3] * uym.v[j] + w[4] * uyp.v[j] +
                w[5] * xl + w[6] * xr;
        }
        f.v[j] = a;
    }
    return f;
}

template <typename T, int VW>
__device__ __noinline__ Pack<T, VW> s7_slow_adj(const T* __restrict__ tab, S7Geom gm, T scale, int zg, int y, int x0,
                                                Pack<T, VW> fc, Pack<T, VW> fm, Pack<T, VW> fp, Pack<T, VW> fym,
                                                Pack<T, VW> fyp, T fl, T fr) {
    const int C1 = 2 * gm.R1 + 1, C2 = 2 * gm.R2 + 1;
    // classes of the source planes / rows (clamped: a source outside the domain has F == 0)
    const int cz0 = s7_cls(zg, gm.N0g, gm.R0);
    const int czm = s7_cls(s7_clamp(zg - 1, gm.N0g), gm.N0g, gm.R0), czp = s7_cls(s7_clamp(zg + 1, gm.N0g), gm.N0g, gm.R0);
    const int cy0 = s7_cls(y, gm.N1, gm.R1);
    const int cym = s7_cls(s7_clamp(y - 1, gm.N1), gm.N1, gm.R1), cyp = s7_cls(s7_clamp(y + 1, gm.N1), gm.N1, gm.R1);
    Pack<T, VW> g;
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const int x = x0 + j;
        T s = T(0);
        if (x < gm.N2) {
            const int cx0 = s7_cls(x, gm.N2, gm.R2);
            const int cxm = s7_cls(s7_clamp(x - 1, gm.N2), gm.N2, gm.R2), cxp = s7_cls(s7_clamp(x + 1, gm.N2), gm.N2, gm.R2);
            const T xl = j > 0 ? fc.v[j > 0 ? j - 1 : 0] : fl;
            const T xr = j < VW - 1 ? fc.v[j < VW - 1 ? j + 1 : 0] : fr;
            s = tab[((cz0 * C1 + cy0) * C2 + cx0) * 7 + 0] * fc.v[j];
            s += tab[((czp * C1 + cy0) * C2 + cx0) * 7 + 1] * fp.v[j];
            s += tab[((czm * C1 + cy0) * C2 + cx0) * 7 + 2] * fm.v[j];
            s += tab[((cz0 * C1 + cyp) * C2 + cx0) * 7 + 3] * fyp.v[j];
            s += tab[((cz0 * C1 + cym) * C2 + cx0) * 7 + 4] * fym.v[j];
            s += tab[((cz0 * C1 + cy0) * C2 + cxp) * 7 + 5] * xr;
            s += tab[((cz0 * C1 + cy0) * C2 + cxm) * 7 + 6] * xl;
            s *= scale;
        }
        g.v[j] = s;
    }
    return g;
}

#ifdef ODIL_B200_LEGACY
// State of one strip-warp thread over the sweep (all members resolve to registers: every index is a
// compile-time constant once step<PH> / lean<PH> are inlined).
template <typename T, int VW, int TY, bool XU>
struct S7Strip {
    using Cfg = Star7Cfg<T, VW, TY>;
    using PackT = Pack<T, VW>;
    static constexpr int BXB = Cfg::BXB, SLOT_U = Cfg::SLOT_U, SLOT_C = Cfg::SLOT_C, SLOT_F = Cfg::SLOT_F;
    static constexpr int NT = Cfg::NT;

    PackT U[3][2];  // own columns of three consecutive U planes, rows A and B
    PackT F[3][2];  // own columns of three consecutive F planes
    S7W<T, VW, XU> wf;  // coefficients of the own cells (x class per cell, interior y/z class)
    T wxpL, wxmR;   // xp coefficient of cell x0-1, xm coefficient of cell x0+VW
    T accf;
    T scale;
    uint32_t su, sc, sf;  // shared byte addresses of the own cells of row A in U slot 0 / c slot 0 / F slot 0
    uint32_t sbar;        // shared address of bar_stage[0]
    uint32_t sbarf;       // shared address of bar_f (F-ring hand-over, one arrival per warp and plane)
    int eoffB;            // byte offset from the own cells to the x-neighbour outside the warp (edge lanes)
    bool edge, lane0, xin;

    static __device__ __forceinline__ PackT lds(uint32_t a) { return s7_lds(a, (PackT*)nullptr); }
    static __device__ __forceinline__ PackT zero() {
        PackT z;
#pragma unroll
        for (int j = 0; j < VW; ++j) z.v[j] = T(0);
        return z;
    }
    // x-neighbours outside the lane's vector: lane shuffles; the two edge lanes of the warp read the
    // staged plane instead (`base` = shared address of the own cells of row A in that plane)
    __device__ __forceinline__ void xnb(const PackT& a, const PackT& b, uint32_t base, T& lA, T& rA, T& lB, T& rB) const {
        lA = __shfl_up_sync(0xffffffffu, a.v[VW - 1], 1);
        rA = __shfl_down_sync(0xffffffffu, a.v[0], 1);
        lB = __shfl_up_sync(0xffffffffu, b.v[VW - 1], 1);
        rB = __shfl_down_sync(0xffffffffu, b.v[0], 1);
        const T eA = s7_lds1_if(base + eoffB, edge, (T*)nullptr), eB = s7_lds1_if(base + BXB + eoffB, edge, (T*)nullptr);
        if (edge) {
            lA = lane0 ? eA : lA;
            lB = lane0 ? eB : lB;
            rA = lane0 ? rA : eA;
            rB = lane0 ? rB : eB;
        }
    }

    // ---- steady state: plane kf and its two predecessors are interior planes owned by the chunk, the rows
    //      of the strip and their y-neighbours are interior rows, c is present, F is not stored.
    //      gptr points at g[kf-1] of row A and is advanced by one plane.
    template <int PH>
    __device__ __forceinline__ void lean(const int it, const uint32_t par, T*& gptr, const int rowB, const int64_t plane) {
        constexpr int IM = PH, IC = (PH + 1) % 3, IP = (PH + 2) % 3;  // U planes kf-1, kf, kf+1
        constexpr int JP = PH, JC = (PH + 2) % 3, JM = (PH + 1) % 3;  // F planes kf, kf-1, kf-2
        const uint32_t ucur = su + ((it + 1) & 3) * SLOT_U;           // plane kf
        const uint32_t unxt = su + ((it + 2) & 3) * SLOT_U;           // plane kf+1
        s7_mbar_wait(sbar + 8 * PH, par);
        U[IP][0] = lds(unxt);
        U[IP][1] = lds(unxt + BXB);
        const PackT uyt = lds(ucur - BXB);
        const PackT uyb = lds(ucur + 2 * BXB);
        const PackT cA = lds(sc + PH * SLOT_C);
        const PackT cB = lds(sc + PH * SLOT_C + BXB);
        T ulA, urA, ulB, urB;
        xnb(U[IC][0], U[IC][1], ucur, ulA, urA, ulB, urB);
        F[JP][0] = s7_fwd<T, VW, XU>(wf, cA, U[IC][0], U[IM][0], U[IP][0], uyt, U[IC][1], ulA, urA);
        F[JP][1] = s7_fwd<T, VW, XU>(wf, cB, U[IC][1], U[IM][1], U[IP][1], U[IC][0], uyb, ulB, urB);
        // every warp has published F[kf-1] (and is done reading F[kf-3], whose slot F[kf] takes)
        s7_mbar_wait(sbarf, (uint32_t)((it - 1) & 1));
        s7_sts(sf + PH * SLOT_F, F[JP][0]);
        s7_sts(sf + PH * SLOT_F + BXB, F[JP][1]);
        __syncwarp();
        if (lane0) s7_mbar_arrive(sbarf);
#pragma unroll
        for (int j = 0; j < VW; ++j) accf = fma(F[JP][0].v[j], F[JP][0].v[j], accf);
#pragma unroll
        for (int j = 0; j < VW; ++j) accf = fma(F[JP][1].v[j], F[JP][1].v[j], accf);
        // neighbours of F[kf-1]
        const uint32_t fprev = sf + JC * SLOT_F;
        const PackT fyt = lds(fprev - BXB);
        const PackT fyb = lds(fprev + 2 * BXB);
        T flA, frA, flB, frB;
        xnb(F[JC][0], F[JC][1], fprev, flA, frA, flB, frB);
        const PackT gA = s7_adj<T, VW, XU>(wf, wxpL, wxmR, scale, F[JC][0], F[JM][0], F[JP][0], fyt, F[JC][1], flA, frA);
        const PackT gB = s7_adj<T, VW, XU>(wf, wxpL, wxmR, scale, F[JC][1], F[JM][1], F[JP][1], F[JC][0], fyb, flB, frB);
        if (xin) {
            *reinterpret_cast<PackT*>(gptr) = gA;
            *reinterpret_cast<PackT*>(gptr + rowB) = gB;
        }
        gptr += plane;
    }

    // ---- general plane / row: flags for everything, boundary classes through the table.
    struct Flags {
        const T* tab;
        T* Gcol;
        T* Fcol;
        S7Geom gm;
        int64_t plane;
        int rowB;
        int x0, yA, yB, kf0, zs, ze, z0;
        bool domA, domB, slowFA, slowFB, slowGA, slowGB, has_c;
    };
    static __device__ __forceinline__ bool zint(const S7Geom& gm, int z) {
        return z < 0 || z >= gm.N0g || s7_cls(z, gm.N0g, gm.R0) == gm.R0;
    }

    template <int PH>
    __device__ __forceinline__ void step(const int it, const uint32_t par, const Flags& fl) {
        constexpr int IM = PH, IC = (PH + 1) % 3, IP = (PH + 2) % 3;
        constexpr int JP = PH, JC = (PH + 2) % 3, JM = (PH + 1) % 3;
        const int kf = fl.kf0 + it;
        const int zg = fl.z0 + kf;
        const uint32_t ucur = su + ((it + 1) & 3) * SLOT_U;
        const uint32_t unxt = su + ((it + 2) & 3) * SLOT_U;
        s7_mbar_wait(sbar + 8 * PH, par);
        U[IP][0] = lds(unxt);
        U[IP][1] = lds(unxt + BXB);
        const PackT uyt = lds(ucur - BXB);
        const PackT uyb = lds(ucur + 2 * BXB);
        PackT cA = zero(), cB = zero();
        if (fl.has_c) {
            cA = lds(sc + PH * SLOT_C);
            cB = lds(sc + PH * SLOT_C + BXB);
        }
        T ulA, urA, ulB, urB;
        xnb(U[IC][0], U[IC][1], ucur, ulA, urA, ulB, urB);
        const bool zin = zg >= 0 && zg < fl.gm.N0g;
        const bool zslow = zin && s7_cls(zg, fl.gm.N0g, fl.gm.R0) != fl.gm.R0;
        if (!zin || !fl.domA)
            F[JP][0] = zero();
        else if (zslow || fl.slowFA)
            F[JP][0] = s7_slow_fwd<T, VW>(fl.tab, fl.gm, zg, fl.yA, fl.x0, cA, U[IC][0], U[IM][0], U[IP][0], uyt, U[IC][1],
                                          ulA, urA);
        else
            F[JP][0] = s7_fwd<T, VW, XU>(wf, cA, U[IC][0], U[IM][0], U[IP][0], uyt, U[IC][1], ulA, urA);
        if (!zin || !fl.domB)
            F[JP][1] = zero();
        else if (zslow || fl.slowFB)
            F[JP][1] = s7_slow_fwd<T, VW>(fl.tab, fl.gm, zg, fl.yB, fl.x0, cB, U[IC][1], U[IM][1], U[IP][1], U[IC][0], uyb,
                                          ulB, urB);
        else
            F[JP][1] = s7_fwd<T, VW, XU>(wf, cB, U[IC][1], U[IM][1], U[IP][1], U[IC][0], uyb, ulB, urB);
        if (it > 0) s7_mbar_wait(sbarf, (uint32_t)((it - 1) & 1));
        s7_sts(sf + PH * SLOT_F, F[JP][0]);
        s7_sts(sf + PH * SLOT_F + BXB, F[JP][1]);
        __syncwarp();
        if (lane0) s7_mbar_arrive(sbarf);
        if (kf >= fl.zs && kf < fl.ze) {
            if (fl.domA) {
#pragma unroll
                for (int j = 0; j < VW; ++j) accf = fma(F[JP][0].v[j], F[JP][0].v[j], accf);
                if (fl.Fcol && xin) *reinterpret_cast<PackT*>(fl.Fcol + (int64_t)kf * fl.plane) = F[JP][0];
            }
            if (fl.domB) {
#pragma unroll
                for (int j = 0; j < VW; ++j) accf = fma(F[JP][1].v[j], F[JP][1].v[j], accf);
                if (fl.Fcol && xin) *reinterpret_cast<PackT*>(fl.Fcol + (int64_t)kf * fl.plane + fl.rowB) = F[JP][1];
            }
        }
        const uint32_t fprev = sf + JC * SLOT_F;
        const PackT fyt = lds(fprev - BXB);
        const PackT fyb = lds(fprev + 2 * BXB);
        T flA, frA, flB, frB;
        xnb(F[JC][0], F[JC][1], fprev, flA, frA, flB, frB);
        const int kg = kf - 1;
        if (kg >= fl.zs && kg < fl.ze) {
            const int zgg = zg - 1;
            const bool zslowG = !(zint(fl.gm, zgg - 1) && zint(fl.gm, zgg) && zint(fl.gm, zgg + 1));
            if (fl.domA) {
                PackT g;
                if (zslowG || fl.slowGA)
                    g = s7_slow_adj<T, VW>(fl.tab, fl.gm, scale, zgg, fl.yA, fl.x0, F[JC][0], F[JM][0], F[JP][0], fyt,
                                           F[JC][1], flA, frA);
                else
                    g = s7_adj<T, VW, XU>(wf, wxpL, wxmR, scale, F[JC][0], F[JM][0], F[JP][0], fyt, F[JC][1], flA, frA);
                if (xin) *reinterpret_cast<PackT*>(fl.Gcol + (int64_t)kg * fl.plane) = g;
            }
            if (fl.domB) {
                PackT g;
                if (zslowG || fl.slowGB)
                    g = s7_slow_adj<T, VW>(fl.tab, fl.gm, scale, zgg, fl.yB, fl.x0, F[JC][1], F[JM][1], F[JP][1], F[JC][0],
                                           fyb, flB, frB);
                else
                    g = s7_adj<T, VW, XU>(wf, wxpL, wxmR, scale, F[JC][1], F[JM][1], F[JP][1], F[JC][0], fyb, flB, frB);
                if (xin) *reinterpret_cast<PackT*>(fl.Gcol + (int64_t)kg * fl.plane + fl.rowB) = g;
            }
        }
    }
};

template <typename T, int VW, int TY, bool XU>
__global__ void __launch_bounds__(Star7Cfg<T, VW, TY>::NT)
    __maxnreg__((XU ? Star7Cfg<T, VW, TY>::MAXREG_XU : Star7Cfg<T, VW, TY>::MAXREG)) k_star7(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmC, Star7Params<T> p) {
    using Cfg = Star7Cfg<T, VW, TY>;
    using PackT = Pack<T, VW>;
    constexpr int TX = Cfg::TX, BX = Cfg::BX, NWS = Cfg::NWS, NT = Cfg::NT, BXB = Cfg::BXB;
    constexpr int SLOT_U = Cfg::SLOT_U, SLOT_C = Cfg::SLOT_C, SLOT_F = Cfg::SLOT_F, SZ = (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_stage[3];
    __shared__ __align__(8) uint64_t bar_pro;
    __shared__ __align__(8) uint64_t bar_f;
    __shared__ double red[32];
    const uint32_t sm0 = smem_u32(smem_raw);
    const uint32_t sU = sm0 + Cfg::OFF_U, sC = sm0 + Cfg::OFF_C, sF = sm0 + Cfg::OFF_F;
    T* tab_s = reinterpret_cast<T*>(smem_raw + Cfg::OFF_TAB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx0 = blockIdx.x * TX, ty0 = blockIdx.y * TY;
    const int zs = blockIdx.z * p.zchunk;
    const int ze = min(zs + p.zchunk, p.n0);
    const int kf0 = zs - 1;
    const int niter = ze - zs + 2;  // F planes zs-1 .. ze
    const int C1 = 2 * p.R1 + 1, C2 = 2 * p.R2 + 1;
    const int ncls = (2 * p.R0 + 1) * C1 * C2;
    const bool tab_in_smem = ncls * 7 <= Cfg::TAB;
    if (tab_in_smem)
        for (int i = tid; i < ncls * 7; i += NT) tab_s[i] = p.table[i];
    const T* __restrict__ tab = tab_in_smem ? tab_s : p.table;
    const S7Geom gm{p.N0g, p.N1, p.N2, p.R0, p.R1, p.R2};
    const bool is_producer = tid == (NWS + 1) * 32;
    const uint32_t sbar = smem_u32(&bar_stage[0]), sbarp = smem_u32(&bar_pro), sbarf = smem_u32(&bar_f);

    if (is_producer) {
        mbar_init(&bar_pro, 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) mbar_init(&bar_stage[i], 1);
        mbar_init(&bar_f, NWS + 2);  // one arrival per warp and plane
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // stage j = { U plane kf0 + 1 + j -> U slot (j + 2) & 3,  c plane kf0 + j -> c slot j % 3 }, barrier j % 3
    auto issue_stage = [&](int j) {
        const int b = j % 3;
        mbar_expect_tx(&bar_stage[b], Cfg::BYTES_U + (p.has_c ? Cfg::BYTES_C : 0u));
        tma_load_3d(smem_raw + Cfg::OFF_U + ((j + 2) & 3) * SLOT_U, &tmU, &bar_stage[b], tx0 - VW, ty0 - 2,
                    kf0 + 1 + j + p.halo);
        if (p.has_c)
            tma_load_3d(smem_raw + Cfg::OFF_C + b * SLOT_C, &tmC, &bar_stage[b], tx0 - VW, ty0 - 1, kf0 + j + p.halo);
    };
    if (is_producer) {
        mbar_expect_tx(&bar_pro, 2 * Cfg::BYTES_U);
        tma_load_3d(smem_raw + Cfg::OFF_U, &tmU, &bar_pro, tx0 - VW, ty0 - 2, kf0 - 1 + p.halo);
        tma_load_3d(smem_raw + Cfg::OFF_U + SLOT_U, &tmU, &bar_pro, tx0 - VW, ty0 - 2, kf0 + p.halo);
        issue_stage(0);
        if (niter > 1) issue_stage(1);
    }
    auto yint = [&](int y) { return y < 0 || y >= p.N1 || s7_cls(y, p.N1, p.R1) == p.R1; };

    double acc2 = 0.0;

    if (warp < NWS) {
        // ------------------------------------------------------------------ strip warps
        using Strip = S7Strip<T, VW, TY, XU>;
        Strip m;
        typename Strip::Flags fl;
        m.scale = p.scale;
        m.sbar = sbar;
        m.sbarf = sbarf;
        fl.tab = tab;
        fl.gm = gm;
        fl.kf0 = kf0;
        fl.zs = zs;
        fl.ze = ze;
        fl.z0 = p.z0;
        fl.has_c = p.has_c != 0;
        fl.x0 = tx0 + VW * lane;
        m.xin = fl.x0 < p.N2;
        const int f0 = 2 * warp + 1;  // F row of strip row A; F row f <-> y = ty0 - 1 + f
        fl.yA = ty0 - 1 + f0;
        fl.yB = fl.yA + 1;
        fl.domA = fl.yA < p.N1;
        fl.domB = fl.yB < p.N1;
        fl.slowFA = fl.domA && !yint(fl.yA);
        fl.slowFB = fl.domB && !yint(fl.yB);
        fl.slowGA = !(yint(fl.yA - 1) && yint(fl.yA) && yint(fl.yA + 1));
        fl.slowGB = !(yint(fl.yB - 1) && yint(fl.yB) && yint(fl.yB + 1));
        m.su = sU + ((f0 + 1) * BX + VW * (lane + 1)) * SZ;
        m.sc = sC + (f0 * BX + VW * (lane + 1)) * SZ;
        m.sf = sF + (f0 * BX + VW * (lane + 1)) * SZ;
        m.edge = lane == 0 || lane == 31;
        m.lane0 = lane == 0;
        m.eoffB = (lane == 0 ? -1 : VW) * SZ;
        {
            const int ibase = (p.R0 * C1 + p.R1) * C2;
            s7_load_w<T, VW, XU>(m.wf, tab, ibase, fl.x0, p.N2, p.R2);
            const int xl = fl.x0 - 1, xr = fl.x0 + VW;
            m.wxpL = xl >= 0 && xl < p.N2 ? tab[(ibase + s7_cls(xl, p.N2, p.R2)) * 7 + 6] : T(0);
            m.wxmR = xr < p.N2 ? tab[(ibase + s7_cls(xr, p.N2, p.R2)) * 7 + 5] : T(0);
        }
        {
            const int yAc = s7_clamp(fl.yA, p.N1);
            const int64_t rowA = (int64_t)yAc * p.N2 + (m.xin ? fl.x0 : 0);
            fl.Gcol = p.G + rowA;
            fl.Fcol = p.Fout ? p.Fout + rowA : nullptr;
            fl.plane = (int64_t)p.N1 * p.N2;
            fl.rowB = (fl.yB - yAc) * p.N2;
        }
        // rows regular: both in the domain, they and their y-neighbours of interior class
        const bool rows_regular = fl.domA && fl.domB && !fl.slowFA && !fl.slowFB && !fl.slowGA && !fl.slowGB &&
                                  fl.has_c && fl.Fcol == nullptr;
        // steady iterations [it_lo, it_hi]: global planes zg-2 .. zg (zg = z0 + kf0 + it) inside the domain with
        // interior class, and both kf = kf0 + it and kf - 1 owned by the chunk (kf in [zs+1, ze-1] <=> it in [2, ze-zs])
        int it_lo = max(p.R0 + 2 - (p.z0 + kf0), 2);
        int it_hi = min(p.N0g - 1 - p.R0 - (p.z0 + kf0), ze - zs);
        if (!rows_regular) it_hi = -1;
        s7_mbar_wait(sbarp, 0);
        m.U[0][0] = m.lds(m.su);
        m.U[0][1] = m.lds(m.su + BXB);
        m.U[1][0] = m.lds(m.su + SLOT_U);
        m.U[1][1] = m.lds(m.su + SLOT_U + BXB);
        m.U[2][0] = m.U[2][1] = m.zero();
#pragma unroll
        for (int q = 0; q < 3; ++q) m.F[q][0] = m.F[q][1] = m.zero();
        m.accf = T(0);
        uint32_t par = 0;
        for (int it = 0; it < niter; it += 3, par ^= 1u) {
            if (it >= it_lo && it + 2 <= it_hi) {
                T* gptr = fl.Gcol + (int64_t)(kf0 + it - 1) * fl.plane;
                m.template lean<0>(it, par, gptr, fl.rowB, fl.plane);
                m.template lean<1>(it + 1, par, gptr, fl.rowB, fl.plane);
                m.template lean<2>(it + 2, par, gptr, fl.rowB, fl.plane);
            } else {
                m.template step<0>(it, par, fl);
                if (it + 1 < niter) m.template step<1>(it + 1, par, fl);
                if (it + 2 < niter) m.template step<2>(it + 2, par, fl);
            }
            acc2 += (double)m.accf;
            m.accf = T(0);
        }
    } else if (warp == NWS) {
        // ------------------------------------------------------------------ y-ring warp: F rows 0 and TY+1
        const int x0 = tx0 + VW * lane;
        const int yT = ty0 - 1, yD = ty0 + TY;
        const bool domT = yT >= 0, domD = yD < p.N1;
        const bool slowT = domT && !yint(yT), slowD = domD && !yint(yD);
        const uint32_t suT = sU + (1 * BX + VW * (lane + 1)) * SZ, suD = sU + ((TY + 2) * BX + VW * (lane + 1)) * SZ;
        const uint32_t scT = sC + (VW * (lane + 1)) * SZ, scD = sC + ((TY + 1) * BX + VW * (lane + 1)) * SZ;
        const uint32_t sfT = sF + (VW * (lane + 1)) * SZ, sfD = sF + ((TY + 1) * BX + VW * (lane + 1)) * SZ;
        const bool edge = lane == 0 || lane == 31, lane0 = lane == 0;
        const int eoffB = (lane == 0 ? -1 : VW) * SZ;
        S7W<T, VW, XU> wf;
        s7_load_w<T, VW, XU>(wf, tab, (p.R0 * C1 + p.R1) * C2, x0, p.N2, p.R2);
        auto lds = [](uint32_t a) { return s7_lds(a, (PackT*)nullptr); };
        PackT zero;
#pragma unroll
        for (int j = 0; j < VW; ++j) zero.v[j] = T(0);
        s7_mbar_wait(sbarp, 0);
        PackT umT = lds(suT), umD = lds(suD), ucT = lds(suT + SLOT_U), ucD = lds(suD + SLOT_U);
        int ph = 0;
        uint32_t par = 0;
        for (int it = 0; it < niter; ++it) {
            const int zg = p.z0 + kf0 + it;
            const uint32_t ocur = ((it + 1) & 3) * SLOT_U, onxt = ((it + 2) & 3) * SLOT_U;
            s7_mbar_wait(sbar + 8 * ph, par);
            const PackT upT = lds(suT + onxt), upD = lds(suD + onxt);
            const bool zin = zg >= 0 && zg < p.N0g;
            const bool zslow = zin && s7_cls(zg, p.N0g, p.R0) != p.R0;
            PackT fT = zero, fD = zero;
            if (zin && domT) {
                const PackT uym = lds(suT + ocur - BXB), uyp = lds(suT + ocur + BXB);
                const PackT cc = p.has_c ? lds(scT + ph * SLOT_C) : zero;
                T ul = __shfl_up_sync(0xffffffffu, ucT.v[VW - 1], 1), ur = __shfl_down_sync(0xffffffffu, ucT.v[0], 1);
                const T e = s7_lds1_if(suT + ocur + eoffB, edge, (T*)nullptr);
                if (edge) {
                    ul = lane0 ? e : ul;
                    ur = lane0 ? ur : e;
                }
                if (zslow || slowT)
                    fT = s7_slow_fwd<T, VW>(tab, gm, zg, yT, x0, cc, ucT, umT, upT, uym, uyp, ul, ur);
                else
                    fT = s7_fwd<T, VW, XU>(wf, cc, ucT, umT, upT, uym, uyp, ul, ur);
            }
            if (zin && domD) {
                const PackT uym = lds(suD + ocur - BXB), uyp = lds(suD + ocur + BXB);
                const PackT cc = p.has_c ? lds(scD + ph * SLOT_C) : zero;
                T ul = __shfl_up_sync(0xffffffffu, ucD.v[VW - 1], 1), ur = __shfl_down_sync(0xffffffffu, ucD.v[0], 1);
                const T e = s7_lds1_if(suD + ocur + eoffB, edge, (T*)nullptr);
                if (edge) {
                    ul = lane0 ? e : ul;
                    ur = lane0 ? ur : e;
                }
                if (zslow || slowD)
                    fD = s7_slow_fwd<T, VW>(tab, gm, zg, yD, x0, cc, ucD, umD, upD, uym, uyp, ul, ur);
                else
                    fD = s7_fwd<T, VW, XU>(wf, cc, ucD, umD, upD, uym, uyp, ul, ur);
            }
            if (it > 0) s7_mbar_wait(sbarf, (uint32_t)((it - 1) & 1));
            s7_sts(sfT + ph * SLOT_F, fT);
            s7_sts(sfD + ph * SLOT_F, fD);
            __syncwarp();
            if (lane0) s7_mbar_arrive(sbarf);
            umT = ucT;
            ucT = upT;
            umD = ucD;
            ucD = upD;
            if (++ph == 3) {
                ph = 0;
                par ^= 1u;
            }
        }
    } else {
        // ------------------------------------------------------------------ x-ring warp (+ TMA producer)
        const int hl = lane;
        const bool active = hl < 2 * TY;
        const int side = hl & 1, f = active ? (hl >> 1) + 1 : 1;  // F rows 1 .. TY
        const int y = ty0 - 1 + f, x = side ? tx0 + TX : tx0 - 1;
        const bool dom = active && y < p.N1 && x >= 0 && x < p.N2;
        const int col = side ? TX + VW : VW - 1;
        const uint32_t uo = sU + ((f + 1) * BX + col) * SZ, co = sC + (f * BX + col) * SZ, fo = sF + (f * BX + col) * SZ;
        T w7[7];
        {
            const int cls = dom ? (p.R0 * C1 + s7_cls(y, p.N1, p.R1)) * C2 + s7_cls(x, p.N2, p.R2) : 0;
#pragma unroll
            for (int o = 0; o < 7; ++o) w7[o] = dom ? tab[cls * 7 + o] : T(0);
        }
        s7_mbar_wait(sbarp, 0);
        T um = s7_lds1(uo, (T*)nullptr), uc = s7_lds1(uo + SLOT_U, (T*)nullptr);
        int ph = 0;
        uint32_t par = 0;
        for (int it = 0; it < niter; ++it) {
            const int zg = p.z0 + kf0 + it;
            const uint32_t ucur = uo + ((it + 1) & 3) * SLOT_U, unxt = uo + ((it + 2) & 3) * SLOT_U;
            s7_mbar_wait(sbar + 8 * ph, par);
            const T up = s7_lds1(unxt, (T*)nullptr);
            T fv = T(0);
            if (dom && zg >= 0 && zg < p.N0g) {
                const T cc = p.has_c ? s7_lds1(co + ph * SLOT_C, (T*)nullptr) : T(0);
                const T uym = s7_lds1(ucur - BXB, (T*)nullptr), uyp = s7_lds1(ucur + BXB, (T*)nullptr);
                const T ul = s7_lds1(ucur - SZ, (T*)nullptr), ur = s7_lds1(ucur + SZ, (T*)nullptr);
                const int cz = s7_cls(zg, p.N0g, p.R0);
                if (cz == p.R0) {
                    fv = cc + w7[0] * uc + w7[1] * um + w7[2] * up + w7[3] * uym + w7[4] * uyp + w7[5] * ul + w7[6] * ur;
                } else {
                    const T* w = tab + ((cz * C1 + s7_cls(y, p.N1, p.R1)) * C2 + s7_cls(x, p.N2, p.R2)) * 7;
                    fv = cc + w[0] * uc + w[1] * um + w[2] * up + w[3] * uym + w[4] * uyp + w[5] * ul + w[6] * ur;
                }
            }
            if (it > 0) s7_mbar_wait(sbarf, (uint32_t)((it - 1) & 1));
            if (active) s7_sts1(fo + ph * SLOT_F, fv);
            __syncwarp();
            if (lane == 0) s7_mbar_arrive(sbarf);
            um = uc;
            uc = up;
            if (is_producer) {
                // every warp has arrived for plane kf: the U plane kf and the c plane kf are consumed
                s7_mbar_wait(sbarf, (uint32_t)(it & 1));
                if (it == 0 && 2 < niter) issue_stage(2);
                if (it + 3 < niter) issue_stage(it + 3);
            }
            if (++ph == 3) {
                ph = 0;
                par ^= 1u;
            }
        }
    }
    const double sum = block_sum(acc2, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = sum;
}

#endif  // ODIL_B200_LEGACY

}  // namespace odil
