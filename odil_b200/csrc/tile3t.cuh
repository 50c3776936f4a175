// k_tile3t: fused residual + loss + adjoint gradient for 3-D grids with ANY offset set of small radius, TMA-fed.
//
// Same sweep as k_tile3d (tile3d.cuh: a CTA owns a 16 x 64 tile of (axis 1, axis 2) and marches along axis 0 with a ring
// of U planes and a ring of F planes in shared memory), rebuilt around what ncu showed there (3 % DRAM throughput:
// every plane was staged with scalar, index-wrapped loads and consumed behind two CTA barriers with nothing in flight):
//   * the U planes (tile + twice the in-plane radius, the x halo rounded up to 16 bytes) arrive by
//     cp.async.bulk.tensor.3d, kT3tPF planes ahead of their first use, zero-filled outside the array, one mbarrier per
//     ring slot;
//   * ONE __syncthreads per plane: the F ring has 2*H0 + 2 slots, so F plane j can be written while slower threads
//     still gather g from F planes j-1-2*H0 .. j-1; the ring-slot offsets of the next plane are published before the
//     barrier of the current one;
//   * every thread owns fixed tile positions (<= 6 F cells incl. the F halo, 4 g cells): their shared-memory offsets,
//     global offsets and class flags are computed once, outside the plane loop; c of the next plane is prefetched into
//     registers behind the barrier;
//   * cells whose class is interior on all three axes (F) / whose sources all are (g) use the interior coefficient row
//     from registers with the offset loop unrolled; everything else takes the table path per cell.  Sources outside
//     the array are skipped: a wrap-free plan (the only kind dispatched here) has a zero coefficient there.
// Per cell the operations and their order are those of k_tile3d / k_generic, so F and g are bit-identical to them; the
// loss partials are summed in a different order (fp64).
// Requires: wrap_free plan, <= 8 offsets, rows a multiple of 16 bytes, no slab.  BASELINE configs[2]: the (t, x, y)
// wave footprint (7 offsets, radii 2 / 1 / 1).
#pragma once
#include "tile3d.cuh"
#include "tma.cuh"

namespace odil {

constexpr int kT3tY = 16, kT3tX = 64, kT3tThreads = 256, kT3tPF = 3, kT3tN = 8, kT3tFC = 6;

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
    unsigned magicF;
    signed char dz[kT3tN], dy[kT3tN], dx[kT3tN];
};

struct Tile3tDims {
    int AH, AW, XH, FH, FW, NU, NF, slotU, slotF;  // slots in bytes
    size_t off_f, off_tab, off_ofs, off_bar, total;
};

template <typename T>
__host__ __device__ inline Tile3tDims t3t_dims(int H0, int H1, int H2, int ncls, int noff) {
    Tile3tDims d;
    d.AH = kT3tY + 4 * H1;
    // x halo of the U tile: 2*H2 rounded up to 16 bytes -- the innermost TMA coordinate (tx0 - XH) must address a
    // 16-byte aligned element (a start at -2 floats faults with "illegal instruction"; measured)
    constexpr int kVec = 16 / (int)sizeof(T);
    d.XH = ((2 * H2 + kVec - 1) / kVec) * kVec;
    d.AW = kT3tX + 2 * d.XH;
    d.FH = kT3tY + 2 * H1;
    d.FW = kT3tX + 2 * H2;
    d.NU = 2 * H0 + 1 + kT3tPF;
    d.NF = 2 * H0 + 2;
    d.slotU = ((d.AH * d.AW * (int)sizeof(T) + 127) / 128) * 128;
    d.slotF = ((d.FH * d.FW * (int)sizeof(T) + 15) / 16) * 16;
    d.off_f = (size_t)d.NU * d.slotU;
    d.off_tab = d.off_f + (size_t)d.NF * d.slotF;
    d.off_ofs = ((d.off_tab + (size_t)ncls * noff * sizeof(T) + 15) / 16) * 16;
    d.off_bar = d.off_ofs + 4 * kT3tN * sizeof(int);
    d.total = d.off_bar + 8 * (size_t)d.NU + 128;  // + slack for the 128-byte alignment of the base
    return d;
}

template <typename T>
__global__ void __launch_bounds__(kT3tThreads, sizeof(T) == 4 ? 3 : 1) k_tile3t(const __grid_constant__ CUtensorMap tmU,
                                                          const __grid_constant__ Tile3tParams<T> p) {
    extern __shared__ unsigned char t3t_raw[];
    __shared__ double red[32];
    unsigned char* smem = t3t_raw + ((128u - (smem_u32(t3t_raw) & 127u)) & 127u);
    const Tile3tDims d = t3t_dims<T>(p.H0, p.H1, p.H2, p.ncls, p.noff);
    const int AW = d.AW, FW = d.FW, FHW = d.FH * d.FW, NU = d.NU, NF = d.NF;
    const int slotU = d.slotU / (int)sizeof(T), slotF = d.slotF / (int)sizeof(T);  // in elements
    const T* sU = reinterpret_cast<const T*>(smem);
    T* sF = reinterpret_cast<T*>(smem + d.off_f);
    T* sTab = reinterpret_cast<T*>(smem + d.off_tab);
    int* sOU = reinterpret_cast<int*>(smem + d.off_ofs);  // [2][kT3tN]: U offsets of the offsets, by plane parity
    int* sOF = sOU + 2 * kT3tN;                           // [2][kT3tN]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + d.off_bar);

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

#ifdef ODIL_B200_DEBUG_MBAR
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("k_tile3t<%d> block z %d: NU %d NF %d slotU %d slotF %d off_bar %d total %d pbase %d plast %d smem 0x%x full 0x%x\n",
               (int)sizeof(T), blockIdx.z, NU, NF, d.slotU, d.slotF, (int)d.off_bar, (int)d.total, pbase, plast,
               smem_u32(smem), smem_u32(full));
#endif
    if (tid == 0) {
        for (int i = 0; i < NU; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < p.ncls * noff; i += kT3tThreads) sTab[i] = p.table[i];
    // ring-slot offsets of plane j (relative number jr = j - j0): written by the threads 0 .. 7
    auto publish_ou = [&](int jr) {  // U offsets for the F phase of plane j0 + jr
        const bool on = tid < noff;
        const int q = jr + H0 + (on ? p.dz[tid] : 0);  // U plane j + dz, relative to pbase
        sOU[(jr & 1) * kT3tN + tid] = on ? (q % NU) * slotU + p.dy[tid] * AW + p.dx[tid] : 0;
    };
    auto publish_of = [&](int jr) {  // F offsets for the g phase of plane k = j - H0
        const bool on = tid < noff;
        const int q = jr - H0 - (on ? p.dz[tid] : 0);  // F plane k - dz, relative to j0 (>= -2*H0)
        sOF[(jr & 1) * kT3tN + tid] = on ? ((q + NF) % NF) * slotF - (p.dy[tid] * FW + p.dx[tid]) : 0;
    };
    if (tid < kT3tN) publish_ou(0);
    __syncthreads();

    auto issue = [&](int q, int s) {  // U plane pbase + q -> slot s = q % NU
        mbar_expect_tx(&full[s], (uint32_t)(d.AH * AW * (int)sizeof(T)));
        tma_load_3d(smem + (size_t)s * d.slotU, &tmU, &full[s], tx0 - d.XH, ty0 - 2 * H1, pbase + q);
    };
    if (tid == 0)
        for (int q = 0; q < NU && pbase + q <= plast; ++q) issue(q, q);

    // ---- fixed per-thread tile positions ----
    T wi[kT3tN];
#pragma unroll
    for (int o = 0; o < kT3tN; ++o) wi[o] = o < noff ? p.table[CI * noff + o] : T(0);
    const int nF = (FHW + kT3tThreads - 1) / kT3tThreads;  // <= kT3tFC
    int atU[kT3tFC], gofs[kT3tFC], cyx[kT3tFC];
    unsigned dom = 0, own = 0, fastF = 0;
#pragma unroll
    for (int i = 0; i < kT3tFC; ++i) {
        const int e = tid + kT3tThreads * i;
        const int r = (int)__umulhi((unsigned)e, p.magicF);
        const int cc = e - r * FW;
        const int ly = ty0 - H1 + r, lx = tx0 - H2 + cc;
        const bool in = e < FHW && ly >= 0 && ly < N1 && lx >= 0 && lx < N2;
        atU[i] = (r + H1) * AW + cc + d.XH - H2;
        gofs[i] = in ? ly * N2 + lx : 0;
        cyx[i] = in ? t2_class(ly, N1, p.R1) * C2 + t2_class(lx, N2, p.R2) : 0;
        if (in) dom |= 1u << i;
        if (in && r >= H1 && r < H1 + kT3tY && cc >= H2 && cc < H2 + kT3tX) own |= 1u << i;
        if (in && ly >= p.R1 && ly < N1 - p.R1 && lx >= p.R2 && lx < N2 - p.R2) fastF |= 1u << i;
    }
    // g cells: e = tid + 256 i -> row (tid >> 6) + 4 i, column tid & 63
    const int gr0 = tid >> 6, gcc = tid & 63;
    const int gx = tx0 + gcc;
    const int atF0 = (gr0 + H1) * FW + gcc + H2;
    unsigned okG = 0, fastG = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int y = ty0 + gr0 + 4 * i;
        if (y < N1 && gx < N2) okG |= 1u << i;
        if (y >= p.R1 + H1 && y < N1 - p.R1 - H1 && gx >= p.R2 + H2 && gx < N2 - p.R2 - H2) fastG |= 1u << i;
    }

    // c of the first plane
    T cN[kT3tFC];
    auto prefetch_c = [&](int j) {
        const bool zin = j >= 0 && j < N0 && j <= j1;
#pragma unroll
        for (int i = 0; i < kT3tFC; ++i)
            cN[i] = (p.c && zin && ((dom >> i) & 1u)) ? __ldg(p.c + (int64_t)j * plane + gofs[i]) : T(0);
    };
    prefetch_c(j0);

    // the first 2*H0 planes of the ring (the plane loop waits for one more plane per step)
    for (int q = 0; q < 2 * H0; ++q) mbar_wait(&full[q], 0);

    double acc = 0.0;
    // ring positions, kept as counters (no division in the plane loop): ws / wpar = slot and phase parity of the newest
    // U plane of the step (q = jr + 2*H0), rs = slot of the U plane retired by the step (q = jr), fs = F slot jr % NF
    int ws = 2 * H0, rs = 0, fs = 0;
    uint32_t wpar = 0;
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
#pragma unroll
            for (int o = 0; o < kT3tN; ++o) ou[o] = sOU[(jr & 1) * kT3tN + o];
            const bool zin = j >= 0 && j < N0;
            const bool zfast = j >= p.R0 && j < N0 - p.R0;
            const int czc = zin ? t2_class(j, N0, p.R0) * C1 * C2 : 0;
            const bool own_plane = j >= zs && j < ze;
            T* fdst = sF + fs * slotF;
            if (++fs == NF) fs = 0;
#pragma unroll
            for (int i = 0; i < kT3tFC; ++i) {
                if (i < nF) {
                    const int e = tid + kT3tThreads * i;
                    T f = T(0);
                    if (zin && ((dom >> i) & 1u)) {
                        f = cN[i];
                        const T* a = sU + atU[i];
                        if (zfast && ((fastF >> i) & 1u)) {
#pragma unroll
                            for (int o = 0; o < kT3tN; ++o)
                                if (o < noff) f += wi[o] * a[ou[o]];
                        } else {
                            const T* trow = sTab + (czc + cyx[i]) * noff;
                            const int* so = sOU + (jr & 1) * kT3tN;  // (dynamic index: from shared memory, not registers)
                            for (int o = 0; o < noff; ++o) f += trow[o] * a[so[o]];
                        }
                        if (own_plane && ((own >> i) & 1u)) {
                            acc += (double)f * (double)f;
                            if (p.Fout) p.Fout[(int64_t)j * plane + gofs[i]] = f;
                        }
                    }
                    if (e < FHW) fdst[e] = f;
                }
            }
        }
        if (tid < kT3tN) {
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
        prefetch_c(j + 1);
        // ---------------- g plane k = j - H0 from the F planes k - H0 .. k + H0 (= j)
        const int k = j - H0;
        if (k >= zs) {
            int of[kT3tN];
#pragma unroll
            for (int o = 0; o < kT3tN; ++o) of[o] = sOF[(jr & 1) * kT3tN + o];
            const bool zfast = k >= p.R0 + H0 && k < N0 - p.R0 - H0;
            T* gdst = p.G + (int64_t)k * plane + (int64_t)(ty0 + gr0) * N2 + gx;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if ((okG >> i) & 1u) {
                    const T* a = sF + atF0 + 4 * i * FW;
                    T g = T(0);
                    if (zfast && ((fastG >> i) & 1u)) {
#pragma unroll
                        for (int o = 0; o < kT3tN; ++o)
                            if (o < noff) g += wi[o] * a[of[o]];
                    } else {
                        const int y = ty0 + gr0 + 4 * i;
                        const int* so = sOF + (jr & 1) * kT3tN;
                        for (int o = 0; o < noff; ++o) {
                            const int sz = k - p.dz[o], sy = y - p.dy[o], sx = gx - p.dx[o];
                            if (sz < 0 || sz >= N0 || sy < 0 || sy >= N1 || sx < 0 || sx >= N2) continue;
                            const int cls = (t2_class(sz, N0, p.R0) * C1 + t2_class(sy, N1, p.R1)) * C2 +
                                            t2_class(sx, N2, p.R2);
                            g += sTab[cls * noff + o] * a[so[o]];
                        }
                    }
                    gdst[(int64_t)(4 * i) * N2] = g * p.scale;
                }
            }
        }
    }
    const double s = block_sum(acc, red);
    if (tid == 0) p.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

}  // namespace odil
