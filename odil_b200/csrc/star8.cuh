// k_star8 -- fused residual + loss + adjoint gradient of a 3-D (or 2-D) star stencil in ONE sweep.
// Same pipeline as k_star7 (star7.cuh: TMA plane rings, per-cell coefficient registers, F handed over
// through a shared-memory ring and an mbarrier) with two changes that ncu asked for:
//   * ONE ROW PER WARP instead of two: half the registers per thread and twice the warps per SM for the
//     same work, because k_star7 was latency-bound (3 warps per scheduler, issue slots 40 % used);
//   * the CTAs come from a host-built WORK LIST (tile origin, number of rows, z-range) instead of a regular
//     grid, so that the list fills the 148 SMs in whole waves with equal work: 512^3 becomes 148 CTAs of
//     13-14 rows x 128 columns that sweep all 512 planes in lock-step (k_star7 lost 20-45 % to the tail of
//     its last wave, and short z-chunks paid a 2-plane lead-in each).
//
// Replaces on the reference side (one launch): ctx.field()=roll (core.py:910-975), the operator's
// arithmetic incl. where(index-mask) boundary rows (examples/poisson/poisson.py:57-68,100-113), the loss
// reduction (core.py:1093) and the reverse-mode gradient (core.py:1100-1101).
//
// Warps of a CTA (tile = nrows <= NR rows x TX = 32*VW columns, marching along axis 0 over [zs, ze)):
//   row warp w    row y = ty0 + w: F, the loss partial and g; a lane owns VW consecutive cells.  Own cells of
//                 U[k-1..k+1] and F[k-2..k] stay in registers (3-way unrolled rotation), x-neighbours come
//                 from lane shuffles, y-neighbours from the staged planes / the F ring.
//   y-ring warp   F (only) of the rows y = ty0-1 and y = ty0+nrows.
//   x-ring warp   F (only) of the columns x = tx0-1 and x = tx0+TX, one lane per cell; after publishing its
//                 plane it waits until every warp has consumed the oldest staged planes and re-arms that
//                 TMA stage (lane 0).
// Hand-over of F between warps: a 3-slot ring per row and one "published plane" word per warp in shared
// memory (st.release / ld.acquire).  A warp only waits for the warps whose F it reads (row above, row below,
// x-ring) -- there is no CTA-wide barrier in the sweep, so a slow warp delays its neighbours, not the CTA.
// Coefficients: every row warp keeps the coefficients of ITS row class and ITS cells' x classes in registers
// (forward row + the y-arm coefficients of the rows above / below for the adjoint), so boundary rows and
// columns run the same instruction stream as interior ones; only the few planes whose z class is not
// interior go through the out-of-line table path.
#pragma once

namespace odil {

struct S8Work {
    int tx0, ty0, nrows, zs, ze, pad0, pad1, pad2;
};

template <typename T, int VW, int NR>
struct Star8Cfg {
    static_assert(NR >= 2 && NR <= 16, "2..16 rows per CTA (one x-ring warp)");
    static_assert(sizeof(T) * VW == 16, "a lane owns one 16-byte vector");
    static constexpr int TX = 32 * VW;
    static constexpr int BX = TX + 2 * VW;
    static constexpr int NT = 32 * (NR + 2);  // row warps + y-ring + x-ring (lane 0 = TMA producer)
    static constexpr int RU = NR + 4, RC = NR + 2;
    static constexpr int BXB = BX * (int)sizeof(T);
    static constexpr int SLOT_U = ((RU * BXB + 127) / 128) * 128;
    static constexpr int SLOT_C = ((RC * BXB + 127) / 128) * 128;
    static constexpr int SLOT_F = RC * BXB;
    // Rings of 6 slots (U planes, c planes, stage barriers) with 5 TMA stages in flight (one U plane + one c plane
    // each): the period 6 = 2 x 3 lines up with the 3-way unrolled plane loop, so a row warp derives every ring
    // position from the trip parity.
    static constexpr int NST = 6, NFL = 5;
    // F ring: 4 slots, so that a warp may store F[k] as soon as its readers have published plane k-2 (they are
    // then done with F[k-4]) and needs plane k-1 of its neighbours only when it gathers g[k-1]
    static constexpr int NSU = 6, NSC = 6, NSF = 4;
    static constexpr int TAB = 1024;
    static constexpr int OFF_U = 128, OFF_C = OFF_U + NSU * SLOT_U, OFF_F = OFF_C + NSC * SLOT_C;
    static constexpr int OFF_TAB = OFF_F + NSF * SLOT_F;
    static constexpr uint32_t BYTES_U = RU * BXB;
    static constexpr uint32_t BYTES_C = RC * BXB;
    // stage barriers, prologue barrier and the published-plane words live in the dynamic block too: their shared
    // addresses are then `base + constant` (taking the address of a static __shared__ variable made ptxas
    // re-derive it with S2R SR_CgaCtaId three times per plane)
    static constexpr int OFF_BAR = ((OFF_TAB + (int)sizeof(T) * TAB + 15) / 16) * 16;
    static constexpr int OFF_PUB = OFF_BAR + 8 * (NST + 1);
    static constexpr size_t SMEM = OFF_PUB + 4 * 32 + 64;
    static constexpr int CTAS_PER_SM = (NT <= 320 && 2 * SMEM <= 225 * 1024) ? 2 : 1;
    // register budget: the register file is 4 x 16384 (one bank per scheduler) and warps are dealt round-robin, so
    // a CTA of W warps needs ceil(W / 4) warps' worth of registers in one bank (measured: 544 threads x 120
    // registers is refused with "too many resources", 512 x 128 launches)
    static constexpr int WARPS_PER_BANK = ((NT / 32) * CTAS_PER_SM + 3) / 4;
    static constexpr int RAW = 512 / WARPS_PER_BANK;
    static constexpr int MAXREG = (RAW / 8) * 8 > 255 ? 255 : (RAW / 8) * 8;
};

template <typename T>
struct Star8Params {
    T* G;
    T* Fout;
    double* partials;
    const T* table;  // [ncls][7]: c, zm, zp, ym, yp, xm, xp
    const S8Work* work;
    int n0, N0g, z0, halo;
    int N1, N2;
    int R0, R1, R2;
    T scale;
    int has_c;
    int xr_async;  // x-ring warp re-arms the TMA stages without blocking on the other warps (ODIL_B200_S8_ASYNC)
};

// "published plane" words (one per working warp), release / acquire at CTA scope
__device__ __forceinline__ void s8_publish(uint32_t addr, int it) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(addr), "r"(it) : "memory");
}
__device__ __forceinline__ int s8_peek(uint32_t addr) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// non-blocking test of a TMA stage barrier, issued early so that its latency overlaps with arithmetic
__device__ __forceinline__ uint32_t s8_mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok;
}

// Running ring positions of one warp (warp-uniform): U slot of plane kf / kf+1, c slot and stage barrier of plane kf.
template <int SLOT_U, int SLOT_C, int NSU, int NST>
struct S8Ring {
    uint32_t ucur, unxt, c, bar, par;
    __device__ __forceinline__ void init() {
        ucur = SLOT_U;       // plane kf0 (q = 1)
        unxt = 2 * SLOT_U;   // plane kf0 + 1 (q = 2)
        c = 0;
        bar = 0;
        par = 0;
    }
    __device__ __forceinline__ void advance() {
        ucur = unxt;
        unxt += SLOT_U;
        if (unxt == NSU * SLOT_U) unxt = 0;
        c += SLOT_C;
        bar += 8;
        if (bar == 8 * NST) {
            bar = 0;
            c = 0;
            par ^= 1u;
        }
    }
    // barrier offset / parity of the NEXT plane's stage
    __device__ __forceinline__ uint32_t next_bar() const { return bar + 8 == 8 * NST ? 0u : bar + 8; }
    __device__ __forceinline__ uint32_t next_par() const { return bar + 8 == 8 * NST ? par ^ 1u : par; }
};

// Packed fp32 form of s8_adj (FFMA2 / FMUL2; same chain of operations per cell, bit-identical results).
template <bool XU>
__device__ __forceinline__ Pack<float, 4> s8_adj_pk(const S7W<float, 4, XU>& w, const S7W<float, 4, XU>& wa, float wxpL,
                                                    float wxmR, float scale, const Pack<float, 4>& fc,
                                                    const Pack<float, 4>& fm, const Pack<float, 4>& fp,
                                                    const Pack<float, 4>& fym, const Pack<float, 4>& fyp, float fl,
                                                    float fr) {
    Pack<float, 4> g;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float2 s = __fmul2_rn(s7_pair(w.c, h), s7_pair(fc, h));
        s = __ffma2_rn(s7_arm2(w, 0, h), s7_pair(fp, h), s);
        s = __ffma2_rn(s7_arm2(w, 1, h), s7_pair(fm, h), s);
        s = __ffma2_rn(s7_arm2(wa, 2, h), s7_pair(fyp, h), s);
        s = __ffma2_rn(s7_arm2(wa, 3, h), s7_pair(fym, h), s);
        g.v[2 * h] = s.x;
        g.v[2 * h + 1] = s.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // x arms: the neighbouring cell's row (scalar, see s7_fwd_pk)
        const float xl = j > 0 ? fc.v[j > 0 ? j - 1 : 0] : fl;
        const float xr = j < 3 ? fc.v[j < 3 ? j + 1 : 0] : fr;
        const float axm = j < 3 ? w.xm.v[j < 3 ? j + 1 : 0] : wxmR;
        const float axp = j > 0 ? w.xp.v[j > 0 ? j - 1 : 0] : wxpL;
        g.v[j] = fmaf(axm, xr, g.v[j]);
        g.v[j] = fmaf(axp, xl, g.v[j]);
    }
    const float2 sc2 = make_float2(scale, scale);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float2 s = __fmul2_rn(s7_pair(g, h), sc2);
        g.v[2 * h] = s.x;
        g.v[2 * h + 1] = s.y;
    }
    return g;
}

// g of VW consecutive cells; the y-arm coefficients come from the rows above / below (ayp: yp coefficient of
// the cell at y-1, aym: ym coefficient of the cell at y+1), everything else from the own row.
template <typename T, int VW, bool XU>
__device__ __forceinline__ Pack<T, VW> s8_adj(const S7W<T, VW, XU>& w, const S7W<T, VW, XU>& wa, T wxpL, T wxmR, T scale,
                                              const Pack<T, VW>& fc, const Pack<T, VW>& fm, const Pack<T, VW>& fp,
                                              const Pack<T, VW>& fym, const Pack<T, VW>& fyp, T fl, T fr) {
#if ODIL_B200_FFMA2
    if constexpr (sizeof(T) == 4 && VW == 4) return s8_adj_pk<XU>(w, wa, wxpL, wxmR, scale, fc, fm, fp, fym, fyp, fl, fr);
#endif
    Pack<T, VW> g;
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const T xl = j > 0 ? fc.v[j > 0 ? j - 1 : 0] : fl;
        const T xr = j < VW - 1 ? fc.v[j < VW - 1 ? j + 1 : 0] : fr;
        const T axm = j < VW - 1 ? w.xm.v[j < VW - 1 ? j + 1 : 0] : wxmR;
        const T axp = j > 0 ? w.xp.v[j > 0 ? j - 1 : 0] : wxpL;
        T s = w.c.v[j] * fc.v[j];
        s = fma(w.arm(0, j), fp.v[j], s);    // zm row of the cell in plane k+1 (same y, x)
        s = fma(w.arm(1, j), fm.v[j], s);    // zp row of the cell in plane k-1
        s = fma(wa.arm(2, j), fyp.v[j], s);  // ym row of the cell at y+1
        s = fma(wa.arm(3, j), fym.v[j], s);  // yp row of the cell at y-1
        s = fma(axm, xr, s);
        s = fma(axp, xl, s);
        g.v[j] = s * scale;
    }
    return g;
}

// State of one row-warp thread over the sweep (registers once step<PH> / lean<PH> are inlined).
template <typename T, int VW, int NR, bool XU>
struct S8Row {
    using Cfg = Star8Cfg<T, VW, NR>;
    using PackT = Pack<T, VW>;
    static constexpr int BXB = Cfg::BXB, SLOT_U = Cfg::SLOT_U, SLOT_C = Cfg::SLOT_C, SLOT_F = Cfg::SLOT_F;
    // ring positions of the current trip (3 planes): stage j = it sits in slot j % 6 = 3 h + PH (h = trip parity);
    // uA / uB = byte offset of U slot 3 h / 3 (1 - h), cb = 3 h, par = (it / 6) & 1
    uint32_t uA, uB, cb, par;
    __device__ __forceinline__ void ring_init() {
        uA = 0;
        uB = 3 * SLOT_U;
        cb = 0;
        par = 0;
    }
    __device__ __forceinline__ void ring_next_trip() {
        const uint32_t t = uA;
        uA = uB;
        uB = t;
        cb ^= 3u;
        par ^= (cb == 0u);
    }
    template <int PH>
    __device__ __forceinline__ uint32_t u_cur() const {  // U plane kf: slot (it + 1) % 6
        return PH == 0 ? uA + SLOT_U : (PH == 1 ? uA + 2 * SLOT_U : uB);
    }
    template <int PH>
    __device__ __forceinline__ uint32_t u_nxt() const {  // U plane kf + 1: slot (it + 2) % 6
        return PH == 0 ? uA + 2 * SLOT_U : (PH == 1 ? uB : uB + SLOT_U);
    }
    PackT U[3], F[3];
    S7W<T, VW, XU> wf;  // forward row of the own cells (own y class, interior z class)
    S7W<T, VW, XU> wa;  // only arms 2, 3 are used: ym coefficient of the cell at y+1, yp coefficient of the cell at y-1
    T wxpL, wxmR;
    T accf, scale;
    uint32_t su, sc, sf;  // shared byte addresses of the own cells in U slot 0 / c slot 0 / F slot 0
    uint32_t sbar;        // bar_stage[0]
    uint32_t pub_me, pub_up, pub_dn, pub_x;  // published-plane words: own, row above, row below, x-ring
    int seen;                                // smallest plane seen published by those three
    int eoffB;
    bool edge, lane0, xin;

    static __device__ __forceinline__ PackT lds(uint32_t a) { return s7_lds(a, (PackT*)nullptr); }
    static __device__ __forceinline__ PackT zero() {
        PackT z;
#pragma unroll
        for (int j = 0; j < VW; ++j) z.v[j] = T(0);
        return z;
    }
    __device__ __forceinline__ void xnb(const PackT& a, uint32_t base, T& l, T& r) const {
        l = __shfl_up_sync(0xffffffffu, a.v[VW - 1], 1);
        r = __shfl_down_sync(0xffffffffu, a.v[0], 1);
        const T e = s7_lds1_if(base + eoffB, edge, (T*)nullptr);
        if (edge) {
            l = lane0 ? e : l;
            r = lane0 ? r : e;
        }
    }
    // the three warps this row exchanges F with (row above, row below, x-ring) have published plane `need`:
    // need = it - 2 before storing F[kf] (they are done reading F[kf-4], whose ring slot it takes),
    // need = it - 1 before gathering g[kf-1] (their F[kf-1] is in the ring)
    // `seen` caches the smallest published plane of the three: the poll before the store usually already shows
    // plane it-1, and the wait before the gather then costs nothing
    __device__ __forceinline__ void wait_neighbours(int need, int& seen) const {
        uint32_t spins = 0;
        while (seen < need) {
            const int a = s8_peek(pub_up), b = s8_peek(pub_dn), c = s8_peek(pub_x);
            seen = min(a, min(b, c));
            if (++spins > (1u << 24)) __trap();
        }
    }
    __device__ __forceinline__ void wait_stage(uint32_t bar, uint32_t par, uint32_t ready) const {
        if (!ready) s7_mbar_wait(bar, par);
    }

    // steady state: planes kf-2 .. kf interior and owned, c present, F not stored
    template <int PH>
    __device__ __forceinline__ void lean(const int it, uint32_t& ready, T*& gptr, const int64_t plane) {
        constexpr int IM = PH, IC = (PH + 1) % 3, IP = (PH + 2) % 3;  // U planes kf-1, kf, kf+1
        constexpr int JP = PH, JC = (PH + 2) % 3, JM = (PH + 1) % 3;  // F planes kf, kf-1, kf-2
        const uint32_t ucur = su + u_cur<PH>();
        const uint32_t unxt = su + u_nxt<PH>();
        wait_stage(sbar + 8 * (cb + PH), par, ready);
        U[IP] = lds(unxt);
        const PackT uyt = lds(ucur - BXB);
        const PackT uyb = lds(ucur + BXB);
        const PackT cc = lds(sc + (cb + PH) * SLOT_C);
        T ul, ur;
        xnb(U[IC], ucur, ul, ur);
        F[JP] = s7_fwd<T, VW, XU>(wf, cc, U[IC], U[IM], U[IP], uyt, uyb, ul, ur);
        wait_neighbours(it - 2, seen);
        s7_sts(sf + (it & 3) * SLOT_F, F[JP]);
        __syncwarp();
        if (lane0) s8_publish(pub_me, it);
        // test the next stage now: the answer is there by the time the next plane starts
        ready = PH < 2 ? s8_mbar_test(sbar + 8 * (cb + PH + 1), par)
                       : s8_mbar_test(sbar + 8 * (cb ^ 3u), par ^ (cb == 3u));
#pragma unroll
        for (int j = 0; j < VW; ++j) accf = fma(F[JP].v[j], F[JP].v[j], accf);
        wait_neighbours(it - 1, seen);
        const uint32_t fprev = sf + ((it + 3) & 3) * SLOT_F;
        const PackT fyt = lds(fprev - BXB);
        const PackT fyb = lds(fprev + BXB);
        T fl, fr;
        xnb(F[JC], fprev, fl, fr);
        const PackT g = s8_adj<T, VW, XU>(wf, wa, wxpL, wxmR, scale, F[JC], F[JM], F[JP], fyt, fyb, fl, fr);
        if (xin) *reinterpret_cast<PackT*>(gptr) = g;
        gptr += plane;
    }

    struct Flags {
        const T* tab;
        T* Gcol;
        T* Fcol;
        S7Geom gm;
        int64_t plane;
        int x0, y, kf0, zs, ze, z0;
        bool has_c;
    };
    static __device__ __forceinline__ bool zint(const S7Geom& gm, int z) {
        return z < 0 || z >= gm.N0g || s7_cls(z, gm.N0g, gm.R0) == gm.R0;
    }

    // general plane: z flags for everything, planes of a boundary z class through the table
    template <int PH>
    __device__ __forceinline__ void step(const int it, const Flags& fl) {
        constexpr int IM = PH, IC = (PH + 1) % 3, IP = (PH + 2) % 3;
        constexpr int JP = PH, JC = (PH + 2) % 3, JM = (PH + 1) % 3;
        const int kf = fl.kf0 + it;
        const int zg = fl.z0 + kf;
        const uint32_t ucur = su + u_cur<PH>();
        const uint32_t unxt = su + u_nxt<PH>();
        s7_mbar_wait(sbar + 8 * (cb + PH), par);
        U[IP] = lds(unxt);
        const PackT uyt = lds(ucur - BXB);
        const PackT uyb = lds(ucur + BXB);
        PackT cc = zero();
        if (fl.has_c) cc = lds(sc + (cb + PH) * SLOT_C);
        T ul, ur;
        xnb(U[IC], ucur, ul, ur);
        const bool zin = zg >= 0 && zg < fl.gm.N0g;
        const bool zslow = zin && s7_cls(zg, fl.gm.N0g, fl.gm.R0) != fl.gm.R0;
        if (!zin)
            F[JP] = zero();
        else if (zslow)
            F[JP] = s7_slow_fwd<T, VW>(fl.tab, fl.gm, zg, fl.y, fl.x0, cc, U[IC], U[IM], U[IP], uyt, uyb, ul, ur);
        else
            F[JP] = s7_fwd<T, VW, XU>(wf, cc, U[IC], U[IM], U[IP], uyt, uyb, ul, ur);
        wait_neighbours(it - 2, seen);
        s7_sts(sf + (it & 3) * SLOT_F, F[JP]);
        __syncwarp();
        if (lane0) s8_publish(pub_me, it);
        if (kf >= fl.zs && kf < fl.ze) {
#pragma unroll
            for (int j = 0; j < VW; ++j) accf = fma(F[JP].v[j], F[JP].v[j], accf);
            if (fl.Fcol && xin) *reinterpret_cast<PackT*>(fl.Fcol + (int64_t)kf * fl.plane) = F[JP];
        }
        wait_neighbours(it - 1, seen);
        const uint32_t fprev = sf + ((it + 3) & 3) * SLOT_F;
        const PackT fyt = lds(fprev - BXB);
        const PackT fyb = lds(fprev + BXB);
        T fll, frr;
        xnb(F[JC], fprev, fll, frr);
        const int kg = kf - 1;
        if (kg >= fl.zs && kg < fl.ze) {
            const int zgg = zg - 1;
            const bool zslowG = !(zint(fl.gm, zgg - 1) && zint(fl.gm, zgg) && zint(fl.gm, zgg + 1));
            PackT g;
            if (zslowG)
                g = s7_slow_adj<T, VW>(fl.tab, fl.gm, scale, zgg, fl.y, fl.x0, F[JC], F[JM], F[JP], fyt, fyb, fll, frr);
            else
                g = s8_adj<T, VW, XU>(wf, wa, wxpL, wxmR, scale, F[JC], F[JM], F[JP], fyt, fyb, fll, frr);
            if (xin) *reinterpret_cast<PackT*>(fl.Gcol + (int64_t)kg * fl.plane) = g;
        }
    }
};

template <typename T, int VW, int NR, bool XU>
__global__ void __launch_bounds__(Star8Cfg<T, VW, NR>::NT) __maxnreg__((Star8Cfg<T, VW, NR>::MAXREG))
    k_star8(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmC, Star8Params<T> p) {
    using Cfg = Star8Cfg<T, VW, NR>;
    using PackT = Pack<T, VW>;
    constexpr int TX = Cfg::TX, BX = Cfg::BX, NT = Cfg::NT, BXB = Cfg::BXB;
    constexpr int SLOT_U = Cfg::SLOT_U, SLOT_C = Cfg::SLOT_C, SLOT_F = Cfg::SLOT_F, SZ = (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar_stage = reinterpret_cast<uint64_t*>(smem_raw + Cfg::OFF_BAR);
    uint64_t* bar_pro_p = bar_stage + Cfg::NST;
    int* pub = reinterpret_cast<int*>(smem_raw + Cfg::OFF_PUB);  // published plane per working warp: rows 0..NR-1,
                                                                 // y-ring NR, x-ring NR+1
    __shared__ double red[32];
    uint32_t sm0 = smem_u32(smem_raw);
    // opaque to ptxas: otherwise it re-derives the shared window base (S2R SR_CgaCtaId + LEA) at every use
    asm volatile("" : "+r"(sm0));
    const uint32_t sU = sm0 + Cfg::OFF_U, sC = sm0 + Cfg::OFF_C, sF = sm0 + Cfg::OFF_F;
    T* tab_s = reinterpret_cast<T*>(smem_raw + Cfg::OFF_TAB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const S8Work wk = p.work[blockIdx.x];
    const int tx0 = wk.tx0, ty0 = wk.ty0, nrows = wk.nrows;
    const int zs = wk.zs, ze = wk.ze;
    const int kf0 = zs - 1;
    const int niter = ze - zs + 2;  // F planes zs-1 .. ze
    const int C1 = 2 * p.R1 + 1, C2 = 2 * p.R2 + 1;
    const int ncls = (2 * p.R0 + 1) * C1 * C2;
    const bool tab_in_smem = ncls * 7 <= Cfg::TAB;
    if (tab_in_smem)
        for (int i = tid; i < ncls * 7; i += NT) tab_s[i] = p.table[i];
    const T* __restrict__ tab = tab_in_smem ? tab_s : p.table;
    const S7Geom gm{p.N0g, p.N1, p.N2, p.R0, p.R1, p.R2};
    const uint32_t sbar = sm0 + Cfg::OFF_BAR, sbarp = sbar + 8 * Cfg::NST, spub = sm0 + Cfg::OFF_PUB;
    if (tid < 32) pub[tid] = -1;

    if (tid == 0) {
        mbar_init(bar_pro_p, 1);
#pragma unroll
        for (int i = 0; i < Cfg::NST; ++i) mbar_init(&bar_stage[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // class of a row with the z class interior; rows outside the domain never contribute (their F is 0)
    auto rowbase = [&](int y) { return (p.R0 * C1 + s7_cls(s7_clamp(y, p.N1), p.N1, p.R1)) * C2; };

    double acc2 = 0.0;

    if (warp < NR) {
        if (warp < nrows) {
            // -------------------------------------------------------------- row warps
            using Row = S8Row<T, VW, NR, XU>;
            Row m;
            typename Row::Flags fl;
            m.scale = p.scale;
            m.sbar = sbar;
            m.pub_me = spub + 4 * warp;
            m.pub_up = spub + 4 * (warp > 0 ? warp - 1 : NR);
            m.pub_dn = spub + 4 * (warp < nrows - 1 ? warp + 1 : NR);
            m.pub_x = spub + 4 * (NR + 1);
            fl.tab = tab;
            fl.gm = gm;
            fl.kf0 = kf0;
            fl.zs = zs;
            fl.ze = ze;
            fl.z0 = p.z0;
            fl.has_c = p.has_c != 0;
            fl.x0 = tx0 + VW * lane;
            m.xin = fl.x0 < p.N2;
            const int f0 = warp + 1;  // F row index; F row f <-> y = ty0 - 1 + f
            fl.y = ty0 + warp;
            m.su = sU + ((f0 + 1) * BX + VW * (lane + 1)) * SZ;
            m.sc = sC + (f0 * BX + VW * (lane + 1)) * SZ;
            m.sf = sF + (f0 * BX + VW * (lane + 1)) * SZ;
            m.edge = lane == 0 || lane == 31;
            m.lane0 = lane == 0;
            m.eoffB = (lane == 0 ? -1 : VW) * SZ;
            {
                const int ibase = rowbase(fl.y);
                s7_load_w<T, VW, XU>(m.wf, tab, ibase, fl.x0, p.N2, p.R2);
                const int xl = fl.x0 - 1, xr = fl.x0 + VW;
                m.wxpL = xl >= 0 && xl < p.N2 ? tab[(ibase + s7_cls(xl, p.N2, p.R2)) * 7 + 6] : T(0);
                m.wxmR = xr < p.N2 ? tab[(ibase + s7_cls(xr, p.N2, p.R2)) * 7 + 5] : T(0);
                // adjoint y arms: the ym coefficient of the row below, the yp coefficient of the row above
                S7W<T, VW, XU> wup, wdn;
                s7_load_w<T, VW, XU>(wup, tab, rowbase(fl.y - 1), fl.x0, p.N2, p.R2);
                s7_load_w<T, VW, XU>(wdn, tab, rowbase(fl.y + 1), fl.x0, p.N2, p.R2);
#pragma unroll
                for (int j = 0; j < VW; ++j) {
                    m.wa.set_arm(2, j, wdn.arm(2, j));
                    m.wa.set_arm(3, j, wup.arm(3, j));
                }
            }
            {
                const int64_t row = (int64_t)fl.y * p.N2 + (m.xin ? fl.x0 : 0);
                fl.Gcol = p.G + row;
                fl.Fcol = p.Fout ? p.Fout + row : nullptr;
                fl.plane = (int64_t)p.N1 * p.N2;
            }
            const bool can_lean = fl.has_c && fl.Fcol == nullptr;
            // steady iterations [it_lo, it_hi]: global planes zg-2 .. zg (zg = z0 + kf0 + it) inside the domain with
            // interior class, kf = kf0 + it and kf - 1 owned by the chunk (it in [2, ze - zs])
            const int it_lo = max(p.R0 + 2 - (p.z0 + kf0), 2);
            const int it_hi = can_lean ? min(p.N0g - 1 - p.R0 - (p.z0 + kf0), ze - zs) : -1;
            s7_mbar_wait(sbarp, 0);
            m.U[0] = m.lds(m.su);
            m.U[1] = m.lds(m.su + SLOT_U);
            m.U[2] = m.zero();
#pragma unroll
            for (int q = 0; q < 3; ++q) m.F[q] = m.zero();
            m.accf = T(0);
            uint32_t ready = 0;
            m.ring_init();
            m.seen = -1;
            for (int it = 0; it < niter; it += 3) {
                if (it >= it_lo && it + 2 <= it_hi) {
                    T* gptr = fl.Gcol + (int64_t)(kf0 + it - 1) * fl.plane;
                    m.template lean<0>(it, ready, gptr, fl.plane);
                    m.template lean<1>(it + 1, ready, gptr, fl.plane);
                    m.template lean<2>(it + 2, ready, gptr, fl.plane);
                } else {
                    m.template step<0>(it, fl);
                    if (it + 1 < niter) m.template step<1>(it + 1, fl);
                    if (it + 2 < niter) m.template step<2>(it + 2, fl);
                    ready = 0;
                }
                m.ring_next_trip();
                acc2 += (double)m.accf;
                m.accf = T(0);
            }
        }
    } else if (warp == NR) {
        // ------------------------------------------------------------------ y-ring warp: F rows 0 and nrows+1
        const int x0 = tx0 + VW * lane;
        const int yT = ty0 - 1, yD = ty0 + nrows;
        const bool domT = yT >= 0, domD = yD < p.N1;
        const uint32_t suT = sU + (1 * BX + VW * (lane + 1)) * SZ, suD = sU + ((nrows + 2) * BX + VW * (lane + 1)) * SZ;
        const uint32_t scT = sC + (VW * (lane + 1)) * SZ, scD = sC + ((nrows + 1) * BX + VW * (lane + 1)) * SZ;
        const uint32_t sfT = sF + (VW * (lane + 1)) * SZ, sfD = sF + ((nrows + 1) * BX + VW * (lane + 1)) * SZ;
        const bool edge = lane == 0 || lane == 31, lane0 = lane == 0;
        const int eoffB = (lane == 0 ? -1 : VW) * SZ;
        const uint32_t pub_me = spub + 4 * NR, pub_a = spub, pub_b = spub + 4 * (nrows - 1);
        S7W<T, VW, XU> wfT, wfD;
        s7_load_w<T, VW, XU>(wfT, tab, rowbase(yT), x0, p.N2, p.R2);
        s7_load_w<T, VW, XU>(wfD, tab, rowbase(yD), x0, p.N2, p.R2);
        auto lds = [](uint32_t a) { return s7_lds(a, (PackT*)nullptr); };
        PackT zero;
#pragma unroll
        for (int j = 0; j < VW; ++j) zero.v[j] = T(0);
        s7_mbar_wait(sbarp, 0);
        PackT umT = lds(suT), umD = lds(suD), ucT = lds(suT + SLOT_U), ucD = lds(suD + SLOT_U);
        int ph = 0;
        S8Ring<SLOT_U, SLOT_C, Cfg::NSU, Cfg::NST> ring;
        ring.init();
        for (int it = 0; it < niter; ++it) {
            const int zg = p.z0 + kf0 + it;
            const uint32_t ocur = ring.ucur, onxt = ring.unxt, oc = ring.c;
            s7_mbar_wait(sbar + ring.bar, ring.par);
            ring.advance();
            const PackT upT = lds(suT + onxt), upD = lds(suD + onxt);
            const bool zin = zg >= 0 && zg < p.N0g;
            const bool zslow = zin && s7_cls(zg, p.N0g, p.R0) != p.R0;
            PackT fT = zero, fD = zero;
            if (zin && domT) {
                const PackT uym = lds(suT + ocur - BXB), uyp = lds(suT + ocur + BXB);
                const PackT cc = p.has_c ? lds(scT + oc) : zero;
                T ul = __shfl_up_sync(0xffffffffu, ucT.v[VW - 1], 1), ur = __shfl_down_sync(0xffffffffu, ucT.v[0], 1);
                const T e = s7_lds1_if(suT + ocur + eoffB, edge, (T*)nullptr);
                if (edge) {
                    ul = lane0 ? e : ul;
                    ur = lane0 ? ur : e;
                }
                if (zslow)
                    fT = s7_slow_fwd<T, VW>(tab, gm, zg, yT, x0, cc, ucT, umT, upT, uym, uyp, ul, ur);
                else
                    fT = s7_fwd<T, VW, XU>(wfT, cc, ucT, umT, upT, uym, uyp, ul, ur);
            }
            if (zin && domD) {
                const PackT uym = lds(suD + ocur - BXB), uyp = lds(suD + ocur + BXB);
                const PackT cc = p.has_c ? lds(scD + oc) : zero;
                T ul = __shfl_up_sync(0xffffffffu, ucD.v[VW - 1], 1), ur = __shfl_down_sync(0xffffffffu, ucD.v[0], 1);
                const T e = s7_lds1_if(suD + ocur + eoffB, edge, (T*)nullptr);
                if (edge) {
                    ul = lane0 ? e : ul;
                    ur = lane0 ? ur : e;
                }
                if (zslow)
                    fD = s7_slow_fwd<T, VW>(tab, gm, zg, yD, x0, cc, ucD, umD, upD, uym, uyp, ul, ur);
                else
                    fD = s7_fwd<T, VW, XU>(wfD, cc, ucD, umD, upD, uym, uyp, ul, ur);
            }
            {  // rows 0 and nrows-1 are done reading the ring slots this plane overwrites
                uint32_t spins = 0;
                while (min(s8_peek(pub_a), s8_peek(pub_b)) < it - 2)
                    if (++spins > (1u << 24)) __trap();
            }
            s7_sts(sfT + (it & 3) * SLOT_F, fT);
            s7_sts(sfD + (it & 3) * SLOT_F, fD);
            __syncwarp();
            if (lane0) s8_publish(pub_me, it);
            umT = ucT;
            ucT = upT;
            umD = ucD;
            ucD = upD;
            if (++ph == 3) ph = 0;
        }
    } else {
        // ------------------------------------------------------------------ x-ring warp (+ TMA producer)
        // stage j = { U plane kf0 + 1 + j -> U slot (j + 2) % NSU,  c plane kf0 + j -> c slot j % NST }, barrier j % NST
        constexpr int NST = Cfg::NST, NSU = Cfg::NSU, NFL = Cfg::NFL;
        auto issue_stage = [&](int j) {
            const int b = j % NST;
            mbar_expect_tx(&bar_stage[b], Cfg::BYTES_U + (p.has_c ? Cfg::BYTES_C : 0u));
            tma_load_3d(smem_raw + Cfg::OFF_U + ((j + 2) % NSU) * SLOT_U, &tmU, &bar_stage[b], tx0 - VW, ty0 - 2,
                        kf0 + 1 + j + p.halo);
            if (p.has_c)
                tma_load_3d(smem_raw + Cfg::OFF_C + b * SLOT_C, &tmC, &bar_stage[b], tx0 - VW, ty0 - 1, kf0 + j + p.halo);
        };
        if (lane == 0) {
            mbar_expect_tx(bar_pro_p, 2 * Cfg::BYTES_U);
            tma_load_3d(smem_raw + Cfg::OFF_U, &tmU, bar_pro_p, tx0 - VW, ty0 - 2, kf0 - 1 + p.halo);
            tma_load_3d(smem_raw + Cfg::OFF_U + SLOT_U, &tmU, bar_pro_p, tx0 - VW, ty0 - 2, kf0 + p.halo);
            for (int j = 0; j < NFL - 1 && j < niter; ++j) issue_stage(j);
        }
        const bool active = lane < 2 * nrows;
        const int side = lane & 1, f = active ? (lane >> 1) + 1 : 1;  // F rows 1 .. nrows
        const int y = ty0 - 1 + f, x = side ? tx0 + TX : tx0 - 1;
        const bool dom = active && y < p.N1 && x >= 0 && x < p.N2;
        const int col = side ? TX + VW : VW - 1;
        const uint32_t uo = sU + ((f + 1) * BX + col) * SZ, co = sC + (f * BX + col) * SZ, fo = sF + (f * BX + col) * SZ;
        const uint32_t pub_me = spub + 4 * (NR + 1), pub_row = spub + 4 * (f - 1);
        // as producer, lane l watches one working warp: rows 0 .. nrows-1, then the y-ring
        const uint32_t pub_w = spub + 4 * (lane < nrows ? lane : NR);
        T w7[7];
        {
            const int cls = dom ? (p.R0 * C1 + s7_cls(y, p.N1, p.R1)) * C2 + s7_cls(x, p.N2, p.R2) : 0;
#pragma unroll
            for (int o = 0; o < 7; ++o) w7[o] = dom ? tab[cls * 7 + o] : T(0);
        }
        s7_mbar_wait(sbarp, 0);
        T um = s7_lds1(uo, (T*)nullptr), uc = s7_lds1(uo + SLOT_U, (T*)nullptr);
        int ph = 0;
        S8Ring<SLOT_U, SLOT_C, Cfg::NSU, Cfg::NST> ring;
        ring.init();
        // Stage j >= NFL - 1 may be issued once every working warp has published plane max(j - NFL, 0) (it is then done
        // with the U and c slots the stage overwrites).  Blocking form (default): wait for that right after publishing
        // plane j - NFL, which makes this warp a per-plane meeting point of the CTA.  Non-blocking form (xr_async):
        // poll once per plane, issue whatever has become possible, and block only before waiting on a stage that has
        // not been issued yet -- the row warps are then coupled to their neighbouring rows only.
        const bool xr_async = p.xr_async != 0;
        int nxt = NFL - 1;  // next stage to issue (xr_async)
        for (int it = 0; it < niter; ++it) {
            const int zg = p.z0 + kf0 + it;
            const uint32_t ucur = uo + ring.ucur, unxt = uo + ring.unxt, oc = ring.c;
            if (xr_async) {
                uint32_t spins = 0;
                while (nxt <= it) {  // stage `it` itself is still to be issued (nxt < niter here)
                    const int mp = __reduce_min_sync(0xffffffffu, s8_peek(pub_w));
                    if (mp >= max(nxt - NFL, 0)) {
                        if (lane == 0) issue_stage(nxt);
                        ++nxt;
                    } else if (++spins > (1u << 26)) {
                        __trap();
                    }
                }
            }
            s7_mbar_wait(sbar + ring.bar, ring.par);
            ring.advance();
            const T up = s7_lds1(unxt, (T*)nullptr);
            T fv = T(0);
            if (dom && zg >= 0 && zg < p.N0g) {
                const T cc = p.has_c ? s7_lds1(co + oc, (T*)nullptr) : T(0);
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
            {  // the row warps are done reading the ring slot this plane overwrites
                uint32_t spins = 0;
                while (!__all_sync(0xffffffffu, s8_peek(pub_row) >= it - 2))
                    if (++spins > (1u << 24)) __trap();
            }
            if (active) s7_sts1(fo + (it & 3) * SLOT_F, fv);
            __syncwarp();
            if (lane == 0) s8_publish(pub_me, it);
            um = uc;
            uc = up;
            if (xr_async) {
                const int mp = min(it, __reduce_min_sync(0xffffffffu, s8_peek(pub_w)));  // this warp has published `it`
                while (nxt < niter && mp >= max(nxt - NFL, 0)) {
                    if (lane == 0) issue_stage(nxt);
                    ++nxt;
                }
            } else if ((it == 0 && NFL - 1 < niter) || it + NFL < niter) {
                // every other warp has published plane `it`: the U plane kf and the c plane kf are consumed
                uint32_t spins = 0;
                while (!__all_sync(0xffffffffu, s8_peek(pub_w) >= it))
                    if (++spins > (1u << 26)) __trap();
                if (lane == 0) {
                    if (it == 0 && NFL - 1 < niter) issue_stage(NFL - 1);
                    if (it + NFL < niter) issue_stage(it + NFL);
                }
            }
            if (++ph == 3) ph = 0;
        }
    }
    const double sum = block_sum(acc2, red);
    if (tid == 0) p.partials[blockIdx.x] = sum;
}

}  // namespace odil
