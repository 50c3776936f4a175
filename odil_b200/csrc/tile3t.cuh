// k_tile3t: fused residual + loss + adjoint gradient for 3-D grids with ANY offset set of small radius, TMA-fed.
//
// Same sweep as k_tile3d (tile3d.cuh: a CTA owns a 16 x 64 tile of (axis 1, axis 2) and marches along axis 0 with a ring
// of U planes and a ring of F planes in shared memory), rebuilt around what ncu showed there (3 % DRAM throughput:
// every plane was staged with scalar, index-wrapped loads and consumed behind two CTA barriers with nothing in flight)
// and in the first TMA version (1800 SASS instructions per warp and plane: generic 64-bit addressing of the rings):
//   * the U planes (tile + twice the in-plane radius; the x halo is fixed at 4 elements so that the row pitch is a
//     compile-time constant and the innermost TMA coordinate stays 16-byte aligned -- a start at -2 floats faults with
//     "illegal instruction") arrive by cp.async.bulk.tensor.3d, kT3tPF planes ahead of their first use, zero-filled
//     outside the array, one mbarrier per ring slot;
//   * ONE __syncthreads per plane: the F ring has 2*H0 + 2 slots, so F plane j can be written while slower threads
//     still gather g from F planes j-1-2*H0 .. j-1; the ring-slot offsets of the next plane are published (by warp 7)
//     before the barrier of the current one;
//   * a thread owns ONE COLUMN of the tile: <= 5 F cells (rows rg + 4 i of the F tile) and 4 g cells (rows rg + 4 i).
//     Per plane and stencil offset it forms one 32-bit shared address; the cells are then reached with immediate
//     offsets: 1 LDS + 1 FMA per cell and offset.  The 2*H2 halo columns of the F tile are one extra cell for the
//     threads 128 ..;
//   * warps whose cells all have the interior class on all three axes (F) / whose sources all have it (g) use the
//     interior coefficient row from registers; any other warp reads the coefficient per cell and offset from the table
//     in shared memory (class of the cell = z class of the plane + the (y, x) class byte of a tile-shaped map that is
//     filled once).  F is stored as 0 outside the array, so such sources contribute nothing (a wrap-free plan -- the
//     only kind dispatched here -- has a zero coefficient there anyway);
//   * c of the next plane is prefetched into registers behind the barrier; ring positions are counters (no division in
//     the plane loop).
// Per cell the operations and their order are those of k_tile3d / k_generic, so F and g are bit-identical to them; the
// loss partials are summed per thread and plane in T, then in fp64.
// Requires: wrap_free plan, <= 8 offsets, radii <= 2, rows a multiple of 16 bytes, no slab.  BASELINE configs[2]: the
// (t, x, y) wave footprint (7 offsets, radii 2 / 1 / 1).
#pragma once
#include "tile3d.cuh"
#include "tma.cuh"

namespace odil {

constexpr int kT3tY = 16, kT3tX = 64, kT3tThreads = 256, kT3tPF = 3, kT3tN = 8;
constexpr int kT3tHM = 2;                    // largest in-plane radius
constexpr int kT3tXH = 2 * kT3tHM;           // x halo of the U tile: 16 bytes (fp32) / 32 bytes (fp64)
constexpr int kT3tAW = kT3tX + 2 * kT3tXH;   // U tile row pitch (elements)
constexpr int kT3tFW = kT3tX + 2 * kT3tHM;   // F tile row pitch
constexpr int kT3tFR = (kT3tY + 2 * kT3tHM + 3) / 4;  // F rows per thread (main block)

template <typename T>
struct Tile3tParams {
    const T* c;        // nullable
    T* G;
    T* Fout;           // nullable
    const T* table;    // [ncls][noff]
    double* partials;  // one per CTA
    T scale;
    int N0, N1, N2;
    int R0, R1, R2;
    int H0, H1, H2;
    int noff, ncls;
    int zchunk;
    signed char dz[kT3tN], dy[kT3tN], dx[kT3tN];
};

struct Tile3tDims {
    int AH, FH, NU, NF, slotU, slotF;  // slots in bytes
    size_t off_f, off_tab, off_ofs, off_bar, off_cls, total;
};

template <typename T>
__host__ __device__ inline Tile3tDims t3t_dims(int H0, int H1, int ncls, int noff) {
    Tile3tDims d;
    d.AH = kT3tY + 4 * H1;
    d.FH = kT3tY + 2 * H1;
    d.NU = 2 * H0 + 1 + kT3tPF;
    d.NF = 2 * H0 + 2;
    d.slotU = ((d.AH * kT3tAW * (int)sizeof(T) + 127) / 128) * 128;
    d.slotF = d.FH * kT3tFW * (int)sizeof(T);
    d.off_f = (size_t)d.NU * d.slotU;
    d.off_tab = d.off_f + (size_t)d.NF * d.slotF;
    d.off_ofs = ((d.off_tab + (size_t)ncls * noff * sizeof(T) + 15) / 16) * 16;
    d.off_bar = d.off_ofs + 4 * kT3tN * sizeof(int);
    d.off_cls = d.off_bar + 8 * (size_t)d.NU;  // (y, x) class byte per F-tile position
    d.total = d.off_cls + (size_t)d.FH * kT3tFW + 128;  // + slack for the 128-byte alignment of the base
    return d;
}

__device__ __forceinline__ float t3t_lds(uint32_t a, float*) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double t3t_lds(uint32_t a, double*) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void t3t_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void t3t_sts(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void t3t_lds8i(uint32_t a, int* o) {
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(a));
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) : "r"(a + 16));
}

__device__ __forceinline__ int t3t_ldsb(uint32_t a) {
    int v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
template <typename V>
__device__ __forceinline__ void t3t_keep(V& v) {  // the value stays in its register: no rematerialisation from threadIdx
    if constexpr (sizeof(V) == 8)
        asm volatile("" : "+l"(v));
    else
        asm volatile("" : "+r"(v));
}

template <typename T>
__global__ void __launch_bounds__(kT3tThreads, sizeof(T) == 4 ? 2 : 1)
    k_tile3t(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ Tile3tParams<T> p) {
    constexpr int S = (int)sizeof(T);
    constexpr int AW = kT3tAW, FW = kT3tFW;
    constexpr uint32_t RU4 = 4 * AW * S, RF4 = 4 * FW * S;  // byte distance of the rows rg and rg + 4 in the U / F tile
    extern __shared__ unsigned char t3t_raw[];
    __shared__ double red[32];
    unsigned char* smem = t3t_raw + ((128u - (smem_u32(t3t_raw) & 127u)) & 127u);
    const Tile3tDims d = t3t_dims<T>(p.H0, p.H1, p.ncls, p.noff);
    const int FH = d.FH, NU = d.NU, NF = d.NF;
    T* sTab = reinterpret_cast<T*>(smem + d.off_tab);
    int* sOfs = reinterpret_cast<int*>(smem + d.off_ofs);  // [2][kT3tN] U offsets, then [2][kT3tN] F offsets (bytes)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + d.off_bar);
    const uint32_t sUb = smem_u32(smem), sFb = sUb + (uint32_t)d.off_f, sOb = sUb + (uint32_t)d.off_ofs;

    const int tid = threadIdx.x;
    const int ty0 = blockIdx.y * kT3tY, tx0 = blockIdx.x * kT3tX;
    const int zs = blockIdx.z * p.zchunk, ze = min(zs + p.zchunk, p.N0);
    const int N0 = p.N0, N1 = p.N1, N2 = p.N2, noff = p.noff;
    const int H0 = p.H0, H1 = p.H1, H2 = p.H2;
    const int C1 = 2 * p.R1 + 1, C2 = 2 * p.R2 + 1;
    const int64_t plane = (int64_t)N1 * N2;
    const int CI = (p.R0 * C1 + p.R1) * C2 + p.R2;  // class of a cell that is interior on every axis
    const int j0 = zs - H0, j1 = ze - 1 + H0;       // F planes of this chunk
    const int pbase = j0 - H0, plast = j1 + H0;     // U planes of this chunk

    if (tid == 0) {
        for (int i = 0; i < NU; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < p.ncls * noff; i += kT3tThreads) sTab[i] = p.table[i];
    // (y, x) class byte of every F-tile position (0 outside the array: F is 0 there)
    const uint32_t sTb = sUb + (uint32_t)d.off_tab, sCb = sUb + (uint32_t)d.off_cls;
    for (int e = tid; e < FH * FW; e += kT3tThreads) {
        const int r = e / FW, cc = e - r * FW;
        const int y = ty0 - H1 + r, xx = tx0 - kT3tHM + cc;
        const bool in = y >= 0 && y < N1 && xx >= 0 && xx < N2;
        smem[d.off_cls + e] = (unsigned char)(in ? t2_class(y, N1, p.R1) * C2 + t2_class(xx, N2, p.R2) : 0);
    }
    // ring-slot byte offsets per stencil offset, written by the threads 248 .. 255 (o = tid - 248) by plane parity:
    // U offsets for the F phase of plane j0 + jr, F offsets for the g phase of plane k = j0 + jr - H0
    const int po = tid - (kT3tThreads - kT3tN);
    auto publish_ou = [&](int jr) {
        const bool on = po < noff;
        const int q = jr + H0 + (on ? p.dz[po] : 0);  // U plane j + dz, relative to pbase
        sOfs[(jr & 1) * kT3tN + po] = on ? (q % NU) * d.slotU + (p.dy[po] * AW + p.dx[po]) * S : 0;
    };
    auto publish_of = [&](int jr) {
        const bool on = po < noff;
        const int q = jr - H0 - (on ? p.dz[po] : 0);  // F plane k - dz, relative to j0 (>= 0 whenever g is computed)
        sOfs[(2 + (jr & 1)) * kT3tN + po] = on ? ((q + NF) % NF) * d.slotF - (p.dy[po] * FW + p.dx[po]) * S : 0;
    };
    if (po >= 0) publish_ou(0);
    __syncthreads();

    auto issue = [&](int q, int s) {  // U plane pbase + q -> slot s = q % NU
        mbar_expect_tx(&full[s], (uint32_t)(d.AH * AW * S));
        tma_load_3d(smem + (size_t)s * d.slotU, &tmU, &full[s], tx0 - kT3tXH, ty0 - 2 * H1, pbase + q);
    };
    if (tid == 0)
        for (int q = 0; q < NU && pbase + q <= plast; ++q) issue(q, q);

    // ---- fixed per-thread tile positions: one column ----
    T wi[kT3tN];
#pragma unroll
    for (int o = 0; o < kT3tN; ++o) wi[o] = o < noff ? p.table[CI * noff + o] : T(0);
    const int col = tid & (kT3tX - 1), rg = tid >> 6;
    const int x = tx0 + col;
    const bool xin = x < N2;
    const bool xfast = x >= p.R2 && x < N2 - p.R2, xfastG = x >= p.R2 + H2 && x < N2 - p.R2 - H2;
    int cx = xin ? t2_class(x, N2, p.R2) : 0;
    // F cells of the main block: F row fr = rg + 4 i (y = ty0 - H1 + fr), F column HM + col; U row fr + H1, column XH + col
    uint32_t baseU = sUb + (uint32_t)(((rg + H1) * AW + kT3tXH + col) * S);
    uint32_t baseFw = sFb + (uint32_t)((rg * FW + kT3tHM + col) * S);
    const int yF0 = ty0 - H1 + rg;
    unsigned exF = 0, domF = 0, ownF = 0, fastF = 0;
#pragma unroll
    for (int i = 0; i < kT3tFR; ++i) {
        const int fr = rg + 4 * i, y = yF0 + 4 * i;
        if (fr < FH) exF |= 1u << i;
        const bool in = fr < FH && xin && y >= 0 && y < N1;
        if (in) domF |= 1u << i;
        if (in && fr >= H1 && fr < H1 + kT3tY) ownF |= 1u << i;
        if (in && xfast && y >= p.R1 && y < N1 - p.R1) fastF |= 1u << i;
    }
    // extra F cell (the 2*H2 halo columns of the F tile): threads 128 .. 128 + 2*H2*FH - 1
    const int xt = tid - 128;
    const bool exX = H2 > 0 && xt >= 0 && xt < 2 * H2 * FH;
    int xr = 0, xc = 0;  // F row / F column of the extra cell
    if (exX) {
        xr = xt / (2 * H2);
        const int xi = xt - xr * 2 * H2;
        xc = xi < H2 ? kT3tHM - H2 + xi : kT3tHM + kT3tX + (xi - H2);
    }
    const int yX = ty0 - H1 + xr, xX = tx0 - kT3tHM + xc;
    const bool domX = exX && yX >= 0 && yX < N1 && xX >= 0 && xX < N2;
    int clsX = domX ? t2_class(yX, N1, p.R1) * C2 + t2_class(xX, N2, p.R2) : 0;
    uint32_t uaX = sUb + (uint32_t)(((xr + H1) * AW + xc + kT3tXH - kT3tHM) * S);
    uint32_t faX = sFb + (uint32_t)((xr * FW + xc) * S);
    // g cells: tile row rg + 4 i, F row + H1, F column HM + col
    uint32_t baseFr = sFb + (uint32_t)(((rg + H1) * FW + kT3tHM + col) * S);
    uint32_t baseCr = sCb + (uint32_t)((rg + H1) * FW + kT3tHM + col);  // class byte of the own g cell 0
    unsigned okG = 0, fastG = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int y = ty0 + rg + 4 * i;
        if (y < N1 && xin) okG |= 1u << i;
        if (xfastG && y >= p.R1 + H1 && y < N1 - p.R1 - H1) fastG |= 1u << i;
    }
    int64_t gofsF = (int64_t)yF0 * N2 + x;  // in-plane offset of the first main F cell (valid where domF says so)
    int64_t gofsX = (int64_t)yX * N2 + xX;
    unsigned flagsX = (exX ? 1u : 0u) | (domX ? 2u : 0u);
    // keep all of this in registers (under register pressure ptxas re-derived it from threadIdx on every plane)
    t3t_keep(baseU); t3t_keep(baseFw); t3t_keep(baseFr); t3t_keep(baseCr); t3t_keep(uaX); t3t_keep(faX);
    t3t_keep(exF); t3t_keep(domF); t3t_keep(ownF); t3t_keep(fastF); t3t_keep(okG); t3t_keep(fastG);
    t3t_keep(gofsF); t3t_keep(gofsX); t3t_keep(flagsX); t3t_keep(cx); t3t_keep(clsX);

    // running global pointers (one 64-bit add per plane instead of re-deriving the address per cell)
    const T* cptr = p.c ? p.c + (int64_t)j0 * plane + gofsF : nullptr;    // c of the first main F cell, plane j
    const T* cptrX = p.c ? p.c + (int64_t)j0 * plane + gofsX : nullptr;
    T* fptr = p.Fout ? p.Fout + (int64_t)j0 * plane + gofsF : nullptr;     // Fout of the first main F cell, plane j
    T* gptr = p.G + (int64_t)(j0 - H0) * plane + (int64_t)(ty0 + rg) * N2 + x;  // g cell 0 of plane k = j - H0
    const int64_t rowstep = (int64_t)4 * N2;

    T cN[kT3tFR], cX = T(0);
    auto prefetch_c = [&](int j, const T* cp, const T* cpx) {  // c of plane j (cp, cpx point into that plane)
        const bool zin = cp != nullptr && j >= 0 && j < N0 && j <= j1;
#pragma unroll
        for (int i = 0; i < kT3tFR; ++i) cN[i] = (zin && ((domF >> i) & 1u)) ? __ldg(cp + i * rowstep) : T(0);
        cX = (zin && (flagsX & 2u)) ? __ldg(cpx) : T(0);
    };
    prefetch_c(j0, cptr, cptrX);

    // the first 2*H0 planes of the ring (the plane loop waits for one more plane per step)
    for (int q = 0; q < 2 * H0; ++q) mbar_wait(&full[q], 0);

    double acc = 0.0;
    // ring positions as counters: ws / wpar = slot and phase parity of the newest U plane of the step (q = jr + 2*H0),
    // rs = slot of the U plane retired by the step (q = jr), fsb = byte offset of the F slot jr % NF
    int ws = 2 * H0, rs = 0;
    uint32_t wpar = 0, fsb = 0;
    const uint32_t fs_end = (uint32_t)(NF * d.slotF);
    for (int j = j0; j <= j1; ++j) {
        const int jr = j - j0;
        mbar_wait(&full[ws], wpar);  // newest U plane of this step: j + H0 = pbase + jr + 2*H0
        if (++ws == NU) {
            ws = 0;
            wpar ^= 1u;
        }
        // ---------------- F plane j on the tile plus one in-plane radius
        {
            int ou[kT3tN];
            t3t_lds8i(sOb + (uint32_t)((jr & 1) * kT3tN * 4), ou);
            uint32_t ra[kT3tN];
#pragma unroll
            for (int o = 0; o < kT3tN; ++o) ra[o] = baseU + (uint32_t)ou[o];
            const bool zin = j >= 0 && j < N0;
            const bool zfast = j >= p.R0 && j < N0 - p.R0;
            const int czc = zin ? t2_class(j, N0, p.R0) * C1 : 0;
            const bool own_plane = j >= zs && j < ze;
            const unsigned needF = zin ? domF : 0u;  // cells that compute (cN is 0 for every other cell)
            T f[kT3tFR];
#pragma unroll
            for (int i = 0; i < kT3tFR; ++i) f[i] = cN[i];
            // one decision per warp and plane: every computing cell of the warp has the interior class
            if (__all_sync(0xffffffffu, zfast ? (needF & ~fastF) == 0u : needF == 0u)) {
                // offsets outside, cells inside: five independent FMA chains per thread.  Rows beyond the F tile
                // (bit of exF clear) read whatever follows in shared memory; their result is dropped below.
#pragma unroll
                for (int o = 0; o < kT3tN; ++o) {
                    if (o < noff) {
#pragma unroll
                        for (int i = 0; i < kT3tFR; ++i) f[i] += wi[o] * t3t_lds(ra[o] + i * RU4, (T*)nullptr);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < kT3tFR; ++i) {
                    if ((needF >> i) & 1u) {
                        const int cls = (czc + t2_class(yF0 + 4 * i, N1, p.R1)) * C2 + cx;
                        const uint32_t tb = sTb + (uint32_t)(cls * noff * S);
#pragma unroll
                        for (int o = 0; o < kT3tN; ++o)
                            if (o < noff) f[i] += t3t_lds(tb + o * S, (T*)nullptr) * t3t_lds(ra[o] + i * RU4, (T*)nullptr);
                    }
                }
            }
            T accp = T(0);
            const unsigned ownN = own_plane ? (needF & ownF) : 0u;
#pragma unroll
            for (int i = 0; i < kT3tFR; ++i) {
                if ((exF >> i) & 1u) {  // warp-uniform: the lanes of a warp share rg
                    const T fi = ((needF >> i) & 1u) ? f[i] : T(0);
                    if ((ownN >> i) & 1u) {
                        accp = fma(fi, fi, accp);
                        if (fptr) fptr[i * rowstep] = fi;
                    }
                    t3t_sts(baseFw + fsb + i * RF4, fi);
                }
            }
            if (flagsX & 1u) {
                T fx = cX;
                const uint32_t tb = sTb + (uint32_t)((czc * C2 + clsX) * noff * S);
#pragma unroll
                for (int o = 0; o < kT3tN; ++o)
                    if (o < noff) fx += t3t_lds(tb + o * S, (T*)nullptr) * t3t_lds(uaX + (uint32_t)ou[o], (T*)nullptr);
                if (!(zin && (flagsX & 2u))) fx = T(0);
                t3t_sts(faX + fsb, fx);
            }
            acc += (double)accp;
            fsb += (uint32_t)d.slotF;
            if (fsb == fs_end) fsb = 0;
        }
        if (po >= 0) {
            publish_of(jr);
            publish_ou(jr + 1);
        }
        __syncthreads();
        // every thread is done with U plane j - H0: its slot takes plane j - H0 + NU
        if (tid == 0) {
            const int q = jr + NU;
            if (pbase + q <= plast) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(q, rs);
            }
        }
        if (++rs == NU) rs = 0;
        if (cptr) {
            cptr += plane;
            cptrX += plane;
        }
        if (fptr) fptr += plane;
        prefetch_c(j + 1, cptr, cptrX);
        // ---------------- g plane k = j - H0 from the F planes k - H0 .. k + H0 (= j)
        const int k = j - H0;
        if (k >= zs) {
            int of[kT3tN];
            t3t_lds8i(sOb + (uint32_t)((2 + (jr & 1)) * kT3tN * 4), of);
            uint32_t ra[kT3tN];
#pragma unroll
            for (int o = 0; o < kT3tN; ++o) ra[o] = baseFr + (uint32_t)of[o];
            const bool zfast = k >= p.R0 + H0 && k < N0 - p.R0 - H0;
            T g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = T(0);
            if (__all_sync(0xffffffffu, zfast ? (okG & ~fastG) == 0u : okG == 0u)) {
#pragma unroll
                for (int o = 0; o < kT3tN; ++o) {
                    if (o < noff) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) g[i] += wi[o] * t3t_lds(ra[o] + i * RF4, (T*)nullptr);
                    }
                }
            } else {
                // coefficient of the SOURCE cell's class: z class of its plane + its (y, x) class byte; source planes
                // outside the array are skipped, source cells outside the array have F = 0
#pragma unroll
                for (int o = 0; o < kT3tN; ++o) {
                    if (o < noff) {
                        const int sz = k - p.dz[o];
                        if (sz >= 0 && sz < N0) {
                            const int czs = t2_class(sz, N0, p.R0) * C1 * C2;
                            const uint32_t ca = baseCr - (uint32_t)(p.dy[o] * FW + p.dx[o]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int cls = czs + t3t_ldsb(ca + i * 4 * FW);
                                g[i] += t3t_lds(sTb + (uint32_t)((cls * noff + o) * S), (T*)nullptr) *
                                        t3t_lds(ra[o] + i * RF4, (T*)nullptr);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if ((okG >> i) & 1u) gptr[i * rowstep] = g[i] * p.scale;
        }
        gptr += plane;
    }
    const double s = block_sum(acc, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

}  // namespace odil
