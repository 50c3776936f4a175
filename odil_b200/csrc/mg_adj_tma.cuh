// k_interp_adjoint3t -- the transposed interpolation g_coarse = scale * I^T g_fine of mg_march.cuh, fed by TMA.
//
// k_interp_adjoint3m keeps 8-12 fine vectors per thread in flight with LDG.128 (128 registers, 16 warps per SM) and
// sat at 52 % of the HBM roofline: ncu showed half of the issue slots used and 3.7 long-scoreboard stall cycles per
// issued instruction -- latency-bound at that occupancy.  Here the fine planes come through shared memory instead:
//   * a CTA owns kA3RJ = 8 coarse rows x 64 coarse columns and marches along axis 0 over a chunk of coarse planes;
//   * one producer lane issues ONE cp.async.bulk.tensor.3d per step: the box {136, 20, 2} = both fine planes of the
//     step, the 16 own fine rows + 2 halo rows per side, the 128 own fine columns + 4 per side (16-byte granularity),
//     zero-filled outside the array -- so clipped rows / columns / planes need neither clamping nor predicates;
//   * NS stages (full / empty mbarriers) keep NS-1 boxes (21.8 KB each) in flight per CTA, 2 CTAs per SM;
//   * a consumer warp owns one coarse row: per fine plane 4 (interior) or 6 (near a y face) LDS.128 per lane, the
//     y taps, lane shuffles for the x neighbours (the two edge lanes read their pair from the halo columns), the x
//     taps, and the 6-plane window along axis 0 in registers -- the arithmetic of mg_adj_march, operation for operation
//     (results are bit-identical to k_interp_adjoint3m for finite inputs).
// The joint-pad correction (k_adjoint_joint_fix) runs after it as before.  fp32 only; fp64 stays on the LDG kernel.
#pragma once
#include "tma.cuh"

namespace odil {

constexpr int kA3RJ = 8;                 // coarse rows per CTA = consumer warps
constexpr int kA3Threads = 32 * (kA3RJ + 1);
constexpr int kA3BX = 128 + 8;           // fine columns in the box
constexpr int kA3BY = 2 * kA3RJ + 4;     // fine rows in the box

template <typename T, int NS>
struct A3Cfg {
    static constexpr int ROWB = kA3BX * (int)sizeof(T);
    static constexpr int PLANE = kA3BY * ROWB;
    static constexpr int STAGE = 2 * PLANE;
    static_assert(STAGE % 128 == 0, "stages stay 128-byte aligned");
    static constexpr int OFF_BAR = NS * STAGE;
    static constexpr size_t SMEM = OFF_BAR + 2 * NS * 8 + 128;  // + slack for the 128-byte alignment of the base
};

__device__ __forceinline__ float4 a3_lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 a3_lds2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}

// One consumer warp: coarse row J (smem rows 2*wrow + R0 ..), coarse cells 2k, 2k+1 per lane.
template <bool BY, int NS>
__device__ __forceinline__ void a3_march(const Mg3& m, float scale, float* __restrict__ gc, int Ibeg, int Iend, int out_z0,
                                         int J, int k, int lane, int wrow, uint32_t stage0, uint32_t full0,
                                         uint32_t empty0) {
    using Cfg = A3Cfg<float, NS>;
    using T = float;
    constexpr int NROW = BY ? 6 : 4, R0 = BY ? 0 : 1;
    const bool valid = 2 * k < m.n2;
    const bool lane0 = lane == 0, lane31 = lane == 31;
    const MgW6<T> wx0 = mg_adjw<T>(valid ? 2 * k : 2, m.n2), wx1 = mg_adjw<T>(valid ? 2 * k + 1 : 3, m.n2);
    const MgW6<T> wy = mg_adjw<T>(J, m.n1);
    // byte offsets inside a plane of the box: own vector (fine columns 4k .. 4k+3 = box columns 4*lane+4 ..) and the
    // pair beyond the warp's span (lane 0: box columns 2, 3; lane 31: 132, 133)
    const uint32_t vo = (uint32_t)((2 * wrow + R0) * Cfg::ROWB + (4 * lane + 4) * 4);
    const uint32_t eo = (uint32_t)((2 * wrow + R0) * Cfg::ROWB + (lane0 ? 2 : 132) * 4);
    const bool edge = lane0 || lane31;

    MgQ<T> Q0{T(0), T(0)}, Q1 = Q0, Q2 = Q0, Q3 = Q0;
    int s = 0;
    uint32_t par = 0;
    for (int I = Ibeg - 2; I < Iend; ++I) {
        mbar_wait_a(full0 + 8 * s, par);
        const uint32_t base = stage0 + (uint32_t)s * Cfg::STAGE;
        MgQ<T> Qn[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const uint32_t pb = base + (uint32_t)p * Cfg::PLANE;
            float4 v[NROW];
            float2 ev[NROW];
#pragma unroll
            for (int i = 0; i < NROW; ++i) v[i] = a3_lds4(pb + vo + i * Cfg::ROWB);
            if (edge) {
#pragma unroll
                for (int i = 0; i < NROW; ++i) ev[i] = a3_lds2(pb + eo + i * Cfg::ROWB);
            }
            T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0), e0 = T(0), e1 = T(0);
#pragma unroll
            for (int i = 0; i < NROW; ++i) {
                const T w = BY ? wy.w[R0 + i] : ((i == 0 || i == 3) ? T(0.25) : T(0.75));
                s0 = fma(w, v[i].x, s0);
                s1 = fma(w, v[i].y, s1);
                s2 = fma(w, v[i].z, s2);
                s3 = fma(w, v[i].w, s3);
            }
            if (edge) {
#pragma unroll
                for (int i = 0; i < NROW; ++i) {
                    const T w = BY ? wy.w[R0 + i] : ((i == 0 || i == 3) ? T(0.25) : T(0.75));
                    e0 = fma(w, ev[i].x, e0);
                    e1 = fma(w, ev[i].y, e1);
                }
            }
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
            Qn[p] = MgQ<T>{qa, qb};
        }
        // the stage is consumed (its values are in registers): hand it back to the producer
        __syncwarp();
        if (lane0) mbar_arrive_a(empty0 + 8 * s);
        if (++s == NS) {
            s = 0;
            par ^= 1u;
        }
        if (I >= Ibeg) {
            T a0, a1;
            if (I >= 2 && I <= m.n0 - 3) {
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

// grid (ceil(n2 / 64), ceil(n1 / kA3RJ), z-chunks), block kA3Threads.  `tm`: tensor map over the fine planes
// [zbase, zbase + extent) the range may touch (global fine plane numbers); everything outside reads as zero.
template <int NS>
__global__ void __launch_bounds__(kA3Threads, 2) k_interp_adjoint3t(const __grid_constant__ CUtensorMap tm, Mg3 m,
                                                                    float scale, float* __restrict__ gc, int cz_begin,
                                                                    int cz_end, int out_z0, int zbase, int zc) {
    using Cfg = A3Cfg<float, NS>;
    extern __shared__ unsigned char a3_raw[];
    // 128-byte aligned base (the dynamic block is only guaranteed 16)
    unsigned char* smem = a3_raw + ((128u - (smem_u32(a3_raw) & 127u)) & 127u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* empty = full + NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 32, J0 = blockIdx.y * kA3RJ;
    const int Ibeg = cz_begin + blockIdx.z * zc;
    const int Iend = min(Ibeg + zc, cz_end);
    if (Ibeg >= Iend) return;
    const int nact = min(kA3RJ, m.n1 - J0);  // consumer warps with a row inside the array
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], nact);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kA3RJ) {
        if (lane == 0) {
            const int nsteps = Iend - Ibeg + 2;
            int s = 0;
            uint32_t par = 1;  // parity of the phase BEFORE the one awaited: the first pass over the ring does not wait
            for (int i = 0; i < nsteps; ++i) {
                if (i >= NS) mbar_wait(&empty[s], par);
                mbar_expect_tx(&full[s], (uint32_t)Cfg::STAGE);
                tma_load_3d(smem + s * Cfg::STAGE, &tm, &full[s], 4 * k0 - 4, 2 * J0 - 2, 2 * (Ibeg - 1 + i) - zbase);
                if (++s == NS) {
                    s = 0;
                    par ^= 1u;
                }
            }
        }
        return;
    }
    if (warp >= nact) return;
    const int J = J0 + warp, k = k0 + lane;
    const uint32_t stage0 = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty);
    if (J <= 1 || J >= m.n1 - 2)
        a3_march<true, NS>(m, scale, gc, Ibeg, Iend, out_z0, J, k, lane, warp, stage0, full0, empty0);
    else
        a3_march<false, NS>(m, scale, gc, Ibeg, Iend, out_z0, J, k, lane, warp, stage0, full0, empty0);
}

}  // namespace odil
