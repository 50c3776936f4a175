// Region-typed affine stencil  F = A U + c,  loss = sum F^2,  g = scale * A^T F   (sm_100a).
//
// Replaces, on the reference side: ctx.field()=roll (core.py:910-975), the example operators'
// arithmetic (examples/poisson/poisson.py:57-68,100-113; examples/wave/wave.py:29-75), the loss
// reduction (core.py:1093) and the AD gradient (core.py:1100-1101).
//
// This file holds the plan, the launchers and the C ABI; the kernels live in headers included below:
//   generic.cuh      k_generic    any ndim <= 4, any offsets, per-cell class lookup (forward / adjoint / fused)
//   star8.cuh        k_star8      3-D star stencils: TMA-fed marching sweep, one row per warp (default)
//   star7.cuh        k_star7      its predecessor (tile of row strips)
//   star_legacy.cuh  k_star3d, k_star_v3, k_star_tma   earlier generations, selectable with plan_tune
//   tile2d.cuh       k_tile2d     2-D grids, any offsets of small radius, shared-memory tiles
//   tile3d.cuh       k_tile3d     non-star 3-D plans: 2-D tiles marching along axis 0 with plane rings
// Fused mode everywhere: forward residual + squared-loss partial + adjoint gradient in ONE sweep (U and c read once,
// g written once; F never touches HBM).
#include <algorithm>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "common.cuh"
#include "tma.cuh"

namespace odil {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
std::atomic<int64_t>& launch_counter() {
    static std::atomic<int64_t> c{0};
    return c;
}

double* reduction_scratch(int nslots) {
    static double* buf[64] = {nullptr};
    static int cap[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (cap[dev] < nslots) {
        if (buf[dev]) cudaFree(buf[dev]);
        int n = std::max(nslots, 4 * kMaxPartialBlocks);
        if (cudaMalloc(&buf[dev], sizeof(double) * n) != cudaSuccess) return nullptr;
        cap[dev] = n;
    }
    return buf[dev];
}

constexpr int kMaxBoxes = 8;
constexpr int kPartialCapacity = 1 << 16;

struct BoxList {
    int nbox;
    int64_t lo[kMaxBoxes][ODIL_B200_MAX_NDIM];  // local coords (axis 0 relative to first owned plane)
    int64_t sz[kMaxBoxes][ODIL_B200_MAX_NDIM];
    int64_t start[kMaxBoxes + 1];  // prefix sums of cell counts
};

struct GenParams {
    int ndim, noff;
    int64_t shape[ODIL_B200_MAX_NDIM];   // global shape
    int64_t stride[ODIL_B200_MAX_NDIM];  // element strides of the local arrays
    int64_t n0, z0;
    int halo;
    int off[ODIL_B200_MAX_OFFSETS][ODIL_B200_MAX_NDIM];
    int R[ODIL_B200_MAX_NDIM];
    int cstride[ODIL_B200_MAX_NDIM];  // class-index strides
    int zero_off;                     // index of the all-zero offset, or -1
    int count_mode;                   // 0: every cell counts toward the loss; 1: only cells with a non-interior class
};

}  // namespace odil

using namespace odil;

struct odil_b200_plan {
    int ndim, dtype, noff;
    int64_t shape[ODIL_B200_MAX_NDIM];
    int off[ODIL_B200_MAX_OFFSETS][ODIL_B200_MAX_NDIM];
    int R[ODIL_B200_MAX_NDIM];
    int ncls;
    std::vector<double> table;  // host copy [ncls][noff]
    void* table_dev;            // typed copy
    double* partials;           // device, kPartialCapacity doubles
    int device;
    // tiled-kernel eligibility
    int kind;      // 0 generic, 1 star tiled
    double w[7];   // interior weights: c, zm, zp, ym, yp, xm, xp  (3-D naming; 2-D uses y,x)
    int zchunk;    // 0 => auto
    int variant;   // tile shape variant
    int rmax0;     // max |off| along axis 0
    void* star_table;  // device, typed [ncls][7] (c, zm, zp, ym, yp, xm, xp) when kind == 1
    int star_has_z;    // some class row couples axis 0 (even if the interior row does not)
    int wrap_free;     // no coefficient multiplies a neighbour across a periodic boundary (TMA zero fill is exact)
    int use_tma;       // 1: TMA-fed kernel when eligible
    int use_v3;        // 1: column-group kernel (default when N2 % 4 == 0), 0: v2 tile kernel + shell
    int use_star7;     // 1: k_star7 (default when eligible)
    int no_xu;         // tuning: force the general coefficient registers
    int use_star8;     // 1: k_star8 (default when eligible): one row per warp, host-built work list
    // work list of k_star8 (device copy + the key it was built for)
    mutable void* work_dev;
    mutable int work_cap, work_n, work_key[6];
    int use_tile2d;    // 1: k_tile2d for 2-D grids (default; env ODIL_B200_TILE2D=0 or an explicit variant disables)
    int h2[2];         // stencil radius per axis (2-D plans)
    int use_tile3d;    // 1: k_tile3d for non-star 3-D plans (default; ODIL_B200_TILE3D=0 or variant 81 selects k_generic)
    int h3[3];         // stencil radius per axis (3-D plans)
    int star_xu;       // z-/y-arm coefficients of the interior y/z classes do not depend on the x class
};

#include "generic.cuh"
#include "star_legacy.cuh"   // k_star_v3 always; k_star3d / k_star_tma only with ODIL_B200_LEGACY
#include "star7.cuh"         // helpers shared with star8.cuh; the k_star7 kernel itself only with ODIL_B200_LEGACY
#include "star8.cuh"
#include "tile2d.cuh"
#include "tile3d.cuh"
#include "tile3t.cuh"
#include "tile2w.cuh"
#ifndef ODIL_B200_TILE2W_DEFAULT
#define ODIL_B200_TILE2W_DEFAULT 1  // k_tile2w where the plan fits it (never slower than k_tile2d: profiles/README.md)
#endif
namespace odil {

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static void fill_gen_params(const odil_b200_plan* plan, const odil_b200_slab* slab, GenParams& p) {
    memset(&p, 0, sizeof(p));
    p.ndim = plan->ndim;
    p.noff = plan->noff;
    int64_t st = 1;
    for (int a = plan->ndim - 1; a >= 0; --a) {
        p.shape[a] = plan->shape[a];
        p.stride[a] = st;
        st *= plan->shape[a];
    }
    int cs = 1;
    for (int a = plan->ndim - 1; a >= 0; --a) {
        p.R[a] = plan->R[a];
        p.cstride[a] = cs;
        cs *= 2 * plan->R[a] + 1;
    }
    p.n0 = slab->n0;
    p.z0 = slab->z0;
    p.halo = slab->halo;
    p.zero_off = -1;
    for (int o = 0; o < plan->noff; ++o) {
        bool z = true;
        for (int a = 0; a < plan->ndim; ++a) {
            p.off[o][a] = plan->off[o][a];
            z = z && plan->off[o][a] == 0;
        }
        if (z) p.zero_off = o;
    }
}

static void whole_box(const odil_b200_plan* plan, const odil_b200_slab* slab, BoxList& b) {
    memset(&b, 0, sizeof(b));
    b.nbox = 1;
    int64_t cnt = 1;
    for (int a = 0; a < plan->ndim; ++a) {
        b.lo[0][a] = 0;
        b.sz[0][a] = a == 0 ? slab->n0 : plan->shape[a];
        cnt *= b.sz[0][a];
    }
    b.start[0] = 0;
    b.start[1] = cnt;
}

// Boxes covering every owned cell within `thick[a]` of a domain face along some axis, disjoint.
static void shell_boxes(const odil_b200_plan* plan, const odil_b200_slab* slab, const int* thick, BoxList& b) {
    memset(&b, 0, sizeof(b));
    const int nd = plan->ndim;
    // global [lo, hi) ranges per axis: face-low, face-high, interior
    int64_t flo[ODIL_B200_MAX_NDIM][2], fhi[ODIL_B200_MAX_NDIM][2], inner[ODIL_B200_MAX_NDIM][2];
    bool has_hi[ODIL_B200_MAX_NDIM];
    for (int a = 0; a < nd; ++a) {
        const int64_t n = plan->shape[a];
        const int64_t t = thick[a];
        if (t <= 0) {
            flo[a][0] = flo[a][1] = 0;
            fhi[a][0] = fhi[a][1] = 0;
            has_hi[a] = false;
            inner[a][0] = 0;
            inner[a][1] = n;
        } else if (2 * t >= n) {
            flo[a][0] = 0;
            flo[a][1] = n;
            has_hi[a] = false;
            fhi[a][0] = fhi[a][1] = 0;
            inner[a][0] = inner[a][1] = 0;
        } else {
            flo[a][0] = 0;
            flo[a][1] = t;
            fhi[a][0] = n - t;
            fhi[a][1] = n;
            has_hi[a] = true;
            inner[a][0] = t;
            inner[a][1] = n - t;
        }
    }
    auto clip0 = [&](int64_t& lo, int64_t& hi) {  // global -> local along axis 0
        lo = std::max(lo, slab->z0) - slab->z0;
        hi = std::min(hi, slab->z0 + slab->n0) - slab->z0;
    };
    int nb = 0;
    int64_t total = 0;
    b.start[0] = 0;
    for (int a = 0; a < nd; ++a) {
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && !has_hi[a]) continue;
            int64_t lo[ODIL_B200_MAX_NDIM], hi[ODIL_B200_MAX_NDIM];
            bool empty = false;
            for (int c = 0; c < nd; ++c) {
                if (c == a) {
                    lo[c] = side ? fhi[a][0] : flo[a][0];
                    hi[c] = side ? fhi[a][1] : flo[a][1];
                } else if (c < a) {
                    lo[c] = inner[c][0];
                    hi[c] = inner[c][1];
                } else {
                    lo[c] = 0;
                    hi[c] = plan->shape[c];
                }
                if (c == 0) clip0(lo[c], hi[c]);
                if (hi[c] <= lo[c]) empty = true;
            }
            if (empty) continue;
            int64_t cnt = 1;
            for (int c = 0; c < nd; ++c) {
                b.lo[nb][c] = lo[c];
                b.sz[nb][c] = hi[c] - lo[c];
                cnt *= hi[c] - lo[c];
            }
            total += cnt;
            ++nb;
            b.start[nb] = total;
        }
    }
    b.nbox = nb;
}

template <typename T, int MODE>
static int launch_generic(const odil_b200_plan* plan, const GenParams& p, const BoxList& boxes, GenIO<T> io,
                          cudaStream_t st, int* nblocks_out) {
    const int64_t total = boxes.start[boxes.nbox];
    if (nblocks_out) *nblocks_out = 0;
    if (total == 0) return 0;
    const int64_t nb = (total + 255) / 256;
    ODIL_REQUIRE(nb < (1ll << 31), "generic stencil grid too large");
    if (MODE == 2) ODIL_REQUIRE(nb <= kPartialCapacity, "generic fused grid exceeds the partials workspace");
    k_generic<T, MODE><<<(unsigned)nb, 256, 0, st>>>(p, boxes, io);
    ODIL_LAUNCHED();
    if (nblocks_out) *nblocks_out = (int)nb;
    return 0;
}

#ifdef ODIL_B200_LEGACY
template <typename T, int TY, int TX, int NT>
static int launch_star_cfg(const StarParams<T>& sp, dim3 grid, bool vec, cudaStream_t st) {
    const size_t smem = (size_t)4 * (TY + 2) * (TX + 8) * sizeof(T);
    static bool attr_set[2] = {false, false};  // per template instantiation
    if (!attr_set[vec ? 1 : 0]) {
        if (vec)
            ODIL_CUDA(cudaFuncSetAttribute(k_star3d<T, TY, TX, NT, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            ODIL_CUDA(cudaFuncSetAttribute(k_star3d<T, TY, TX, NT, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[vec ? 1 : 0] = true;
    }
    if (vec)
        k_star3d<T, TY, TX, NT, true><<<grid, NT, smem, st>>>(sp);
    else
        k_star3d<T, TY, TX, NT, false><<<grid, NT, smem, st>>>(sp);
    ODIL_LAUNCHED();
    return 0;
}

#endif  // ODIL_B200_LEGACY

template <typename T, int TY, int TX>
static int launch_star_v3(const StarV3Params<T>& sp, dim3 grid, cudaStream_t st) {
    constexpr int NT = (TX / 4 + 2) * (TY + 4);
    const size_t smem = ((size_t)4 * (TY + 4) * ((TX / 4 + 2) * 4 + 8) + 512) * sizeof(T);
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star_v3<T, TY, TX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        ODIL_CUDA(cudaFuncSetAttribute(k_star_v3<T, TY, TX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        attr_set = true;
    }
    if (sp.has_z)
        k_star_v3<T, TY, TX, true><<<grid, NT, smem, st>>>(sp);
    else
        k_star_v3<T, TY, TX, false><<<grid, NT, smem, st>>>(sp);
    ODIL_LAUNCHED();
    return 0;
}

static void star_v3_tile(int variant, int& TY, int& TX) {
    switch (variant) {
        case 1: TY = 8; TX = 128; break;    // 408 threads
        case 2: TY = 12; TX = 128; break;   // 544 threads
        case 3: TY = 26; TX = 128; break;   // 1020 threads
        default: TY = 16; TX = 128; break;  // 680 threads
    }
}

#ifdef ODIL_B200_LEGACY
template <typename T, int TY, int TX>
static int launch_star_tma(const CUtensorMap& tmU, const CUtensorMap& tmC, const StarTmaParams<T>& sp, dim3 grid,
                           cudaStream_t st) {
    constexpr int NT = (TX / 4 + 2) * (TY + 2);
    constexpr int PLN = (((TY + 4) * (TX + 8) + 31) / 32) * 32;
    const size_t smem = 128 + (size_t)(4 + 2 + 4) * PLN * sizeof(T) + 512 * sizeof(T) + 128;
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star_tma<T, TY, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    k_star_tma<T, TY, TX><<<grid, NT, smem, st>>>(tmU, tmC, sp);
    ODIL_LAUNCHED();
    return 0;
}

template <typename T, int VW, int TY, bool XU>
static int launch_star7(const odil_b200_plan* plan, const odil_b200_slab* slab, const T* U, const T* c, T scale, T* G,
                        T* Fout, int* nparts, cudaStream_t st) {
    using Cfg = Star7Cfg<T, VW, TY>;
    Star7Params<T> sp;
    sp.G = G;
    sp.Fout = Fout;
    sp.partials = plan->partials;
    sp.table = (const T*)plan->star_table;
    if (plan->ndim == 3) {
        sp.n0 = (int)slab->n0;
        sp.N0g = (int)plan->shape[0];
        sp.z0 = (int)slab->z0;
        sp.halo = slab->halo;
        sp.N1 = (int)plan->shape[1];
        sp.N2 = (int)plan->shape[2];
        sp.R0 = plan->R[0];
        sp.R1 = plan->R[1];
        sp.R2 = plan->R[2];
    } else {
        sp.n0 = 1;
        sp.N0g = 1;
        sp.z0 = 0;
        sp.halo = 0;
        sp.N1 = (int)plan->shape[0];
        sp.N2 = (int)plan->shape[1];
        sp.R0 = 0;
        sp.R1 = plan->R[0];
        sp.R2 = plan->R[1];
    }
    sp.scale = scale;
    sp.has_c = c != nullptr;
    const int gx = (sp.N2 + Cfg::TX - 1) / Cfg::TX, gy = (sp.N1 + TY - 1) / TY;
    int zchunk = plan->zchunk;
    if (zchunk <= 0) {
        // fill the 148 x MINB CTA slots in whole waves while keeping the 2-plane lead-in of a chunk small:
        // cost(gz) ~ waves(gz) * (planes per chunk + lead-in)
        const int64_t slots = 148 * (XU ? Cfg::MINB_XU : Cfg::MINB), layer = (int64_t)gx * gy;
        int64_t best = -1;
        for (int gz = 1; gz <= std::max(1, sp.n0 / 8); ++gz) {
            const int zc = (sp.n0 + gz - 1) / gz;
            const int gzr = (sp.n0 + zc - 1) / zc;
            const int64_t cost = ((layer * gzr + slots - 1) / slots) * (zc + 5);
            if (best < 0 || cost < best) {
                best = cost;
                zchunk = zc;
            }
        }
    }
    if (zchunk > sp.n0) zchunk = sp.n0;
    if (zchunk < 1) zchunk = 1;
    sp.zchunk = zchunk;
    const int gz = (sp.n0 + zchunk - 1) / zchunk;
    ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity, "star grid exceeds the partials workspace");
    const int64_t planes_total = (int64_t)sp.n0 + 2 * sp.halo;
    const int64_t plane_elems = (int64_t)sp.N1 * sp.N2;
    CUtensorMap tmU, tmC;
    if (int rc = make_plane_map<T>(&tmU, U - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RU, Cfg::BX)) return rc;
    if (int rc = make_plane_map<T>(&tmC, (c ? c : U) - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RC, Cfg::BX))
        return rc;
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star7<T, VW, TY, XU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    k_star7<T, VW, TY, XU><<<dim3(gx, gy, gz), Cfg::NT, Cfg::SMEM, st>>>(tmU, tmC, sp);
    ODIL_LAUNCHED();
    *nparts = gx * gy * gz;
    return 0;
}

#endif  // ODIL_B200_LEGACY

// Work list of k_star8: tiles of up to NR rows x TX columns times z-chunks, chosen so that the CTAs fill the
// SMs in whole waves with equal work.  cost ~ waves * (planes per chunk + lead-in) * (rows per CTA + ring rows).
static void build_work8(int NR, int TX, int slots, int n0, int N1, int N2, int zchunk, std::vector<S8Work>& out) {
    const int gx = (N2 + TX - 1) / TX;
    int64_t best = -1;
    int best_rmax = NR, best_zc = n0;
    for (int rmax = 1; rmax <= NR; ++rmax) {
        const int nblk = (N1 + rmax - 1) / rmax;
        const int gz_lo = zchunk > 0 ? std::max(1, (n0 + zchunk - 1) / zchunk) : 1;
        const int gz_hi = zchunk > 0 ? gz_lo : std::max(1, n0 / 8);
        for (int gz = gz_lo; gz <= gz_hi; ++gz) {
            const int zc = (n0 + gz - 1) / gz;
            const int gzr = (n0 + zc - 1) / zc;
            const int64_t ncta = (int64_t)gx * nblk * gzr;
            if (ncta > kPartialCapacity) continue;
            const int64_t cost = ((ncta + slots - 1) / slots) * (zc + 6) * (rmax + 3);
            if (best < 0 || cost < best) {
                best = cost;
                best_rmax = rmax;
                best_zc = zc;
            }
        }
    }
    const int nblk = (N1 + best_rmax - 1) / best_rmax;
    out.clear();
    for (int zs = 0; zs < n0; zs += best_zc)
        for (int b = 0; b < nblk; ++b) {
            const int y0 = (int)((int64_t)b * N1 / nblk), y1 = (int)((int64_t)(b + 1) * N1 / nblk);
            if (y1 <= y0) continue;
            for (int bx = 0; bx < gx; ++bx) out.push_back(S8Work{bx * TX, y0, y1 - y0, zs, std::min(n0, zs + best_zc), 0, 0, 0});
        }
}

template <typename T, int VW, int NR, bool XU>
static int launch_star8(const odil_b200_plan* plan, const odil_b200_slab* slab, const T* U, const T* c, T scale, T* G,
                        T* Fout, int* nparts, cudaStream_t st) {
    using Cfg = Star8Cfg<T, VW, NR>;
    Star8Params<T> sp;
    sp.G = G;
    sp.Fout = Fout;
    sp.partials = plan->partials;
    sp.table = (const T*)plan->star_table;
    if (plan->ndim == 3) {
        sp.n0 = (int)slab->n0;
        sp.N0g = (int)plan->shape[0];
        sp.z0 = (int)slab->z0;
        sp.halo = slab->halo;
        sp.N1 = (int)plan->shape[1];
        sp.N2 = (int)plan->shape[2];
        sp.R0 = plan->R[0];
        sp.R1 = plan->R[1];
        sp.R2 = plan->R[2];
    } else {
        sp.n0 = 1;
        sp.N0g = 1;
        sp.z0 = 0;
        sp.halo = 0;
        sp.N1 = (int)plan->shape[0];
        sp.N2 = (int)plan->shape[1];
        sp.R0 = 0;
        sp.R1 = plan->R[0];
        sp.R2 = plan->R[1];
    }
    sp.scale = scale;
    sp.has_c = c != nullptr;
    {
        static const int xr_async = [] {
            const char* e = getenv("ODIL_B200_S8_ASYNC");
            return e ? atoi(e) : 0;
        }();
        sp.xr_async = xr_async;
    }
    const int key[6] = {sp.n0, sp.N1, sp.N2, NR, plan->zchunk, Cfg::TX};
    if (memcmp(key, plan->work_key, sizeof(key)) != 0) {
        std::vector<S8Work> work;
        build_work8(NR, Cfg::TX, 148 * Cfg::CTAS_PER_SM, sp.n0, sp.N1, sp.N2, plan->zchunk, work);
        ODIL_REQUIRE(!work.empty() && (int64_t)work.size() <= kPartialCapacity, "star work list has %lld entries",
                     (long long)work.size());
        if ((int)work.size() > plan->work_cap) {
            if (plan->work_dev) cudaFree(plan->work_dev);
            plan->work_dev = nullptr;
            plan->work_cap = 0;
            ODIL_CUDA(cudaMalloc(&plan->work_dev, sizeof(S8Work) * work.size()));
            plan->work_cap = (int)work.size();
        }
        // pageable source: the copy is staged before the call returns, so `work` may go out of scope
        ODIL_CUDA(cudaMemcpyAsync(plan->work_dev, work.data(), sizeof(S8Work) * work.size(), cudaMemcpyHostToDevice, st));
        plan->work_n = (int)work.size();
        memcpy(plan->work_key, key, sizeof(key));
    }
    sp.work = (const S8Work*)plan->work_dev;
    const int64_t planes_total = (int64_t)sp.n0 + 2 * sp.halo;
    const int64_t plane_elems = (int64_t)sp.N1 * sp.N2;
    CUtensorMap tmU, tmC;
    if (int rc = make_plane_map<T>(&tmU, U - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RU, Cfg::BX)) return rc;
    if (int rc = make_plane_map<T>(&tmC, (c ? c : U) - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, Cfg::RC, Cfg::BX))
        return rc;
    static bool attr_set = false;
    if (!attr_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_star8<T, VW, NR, XU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    k_star8<T, VW, NR, XU><<<plan->work_n, Cfg::NT, Cfg::SMEM, st>>>(tmU, tmC, sp);
    {
        launch_counter()++;
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFuncAttributes fa;
            memset(&fa, 0, sizeof(fa));
            cudaFuncGetAttributes(&fa, k_star8<T, VW, NR, XU>);
            cudaGetLastError();
            return fail("k_star8<NR=%d> launch (%d CTAs x %d threads, %zu B dynamic smem) -> %s [numRegs=%d maxThreadsPerBlock=%d "
                        "static smem=%zu maxDynamic=%d local=%zu]",
                        NR, plan->work_n, Cfg::NT, (size_t)Cfg::SMEM, cudaGetErrorString(e), fa.numRegs, fa.maxThreadsPerBlock,
                        fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, fa.localSizeBytes);
        }
    }
    *nparts = plan->work_n;
    return 0;
}

static void star_tile(int variant, int& TY, int& TX) {
    switch (variant) {
        case 1: TY = 16; TX = 64; break;
        case 2: TY = 16; TX = 128; break;
        case 3: TY = 4; TX = 128; break;
        default: TY = 8; TX = 128; break;
    }
}


// ------------------------------------------------------------------------------------------------
// 2-D tile kernel (tile2d.cuh)
// ------------------------------------------------------------------------------------------------
static bool tile2d_ok(const odil_b200_plan* plan, const odil_b200_slab* slab) {
    if (!plan->use_tile2d || plan->ndim != 2) return false;
    if (slab->halo > 0 || slab->n0 != plan->shape[0] || slab->z0 != 0) return false;
    if (plan->h2[0] > kT2MaxRadius || plan->h2[1] > kT2MaxRadius || plan->ncls > 255) return false;
    const int64_t N0 = plan->shape[0], N1 = plan->shape[1];
    if (N0 * N1 >= (1ll << 31) || N0 < 1 || N1 < 1) return false;
    const int64_t gx = (N1 + kT2X - 1) / kT2X, gy = (N0 + kT2Y - 1) / kT2Y;
    if (gy > 65535 || gx * gy > kPartialCapacity) return false;
    return (size_t)plan->ncls * plan->noff * 8 <= 64 * 1024;
}

template <typename T, int MODE>
static int launch_tile2d(const odil_b200_plan* plan, const T* A, const T* c, T scale, T* out, T* Fout, int* nparts,
                         cudaStream_t st) {
    Tile2Params<T> p;
    p.A = A;
    p.c = c;
    p.out = out;
    p.Fout = Fout;
    p.table = (const T*)plan->table_dev;
    p.partials = plan->partials;
    p.scale = scale;
    p.N0 = (int)plan->shape[0];
    p.N1 = (int)plan->shape[1];
    p.R0 = plan->R[0];
    p.R1 = plan->R[1];
    p.H0 = plan->h2[0];
    p.H1 = plan->h2[1];
    p.noff = plan->noff;
    p.ncls = plan->ncls;
    int AH, AW, FH, FW;
    t2_dims<MODE>(p.H0, p.H1, AH, AW, FH, FW);
    p.magicA = (unsigned)((1ull << 32) / (unsigned)AW + 1);
    p.magicF = (unsigned)((1ull << 32) / (unsigned)FW + 1);
    for (int o = 0; o < ODIL_B200_MAX_OFFSETS; ++o) {
        p.dy[o] = o < plan->noff ? (signed char)plan->off[o][0] : 0;
        p.dx[o] = o < plan->noff ? (signed char)plan->off[o][1] : 0;
    }
    const size_t smem = t2_smem_bytes<T, MODE>(p.H0, p.H1, p.ncls, p.noff);
    static size_t smem_set = 48 * 1024;  // per template instantiation
    if (smem > smem_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_tile2d<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    dim3 grid((p.N1 + kT2X - 1) / kT2X, (p.N0 + kT2Y - 1) / kT2Y);
    k_tile2d<T, MODE><<<grid, kT2Threads, smem, st>>>(p);
    ODIL_LAUNCHED();
    if (nparts) *nparts = (int)(grid.x * grid.y);
    return 0;
}


// k_tile2w (tile2w.cuh): warp-private marching version of the fused 2-D sweep for wrap-free plans with at most 8
// offsets of radius <= 2; ODIL_B200_TILE2W=0 (read per call, so that tests can compare the two) keeps k_tile2d.
template <typename T>
static bool tile2w_fits(const odil_b200_plan* plan, const T* U, const T* c, const T* G, const T* Fout) {
    return plan->wrap_free && plan->noff <= kT2wN && plan->h2[0] <= kT2wHM && plan->h2[1] <= kT2wHM &&
           plan->shape[1] % kT2wVW == 0 && (uintptr_t)U % 16 == 0 && (uintptr_t)c % 16 == 0 && (uintptr_t)G % 16 == 0 &&
           (uintptr_t)Fout % 16 == 0;
}

template <typename T>
static bool tile2w_ok(const odil_b200_plan* plan, const T* U, const T* c, const T* G, const T* Fout) {
    const char* e = getenv("ODIL_B200_TILE2W");
    if (!(e ? atoi(e) != 0 : ODIL_B200_TILE2W_DEFAULT)) return false;
    return tile2w_fits<T>(plan, U, c, G, Fout);
}

// one instantiation per offset count (2 rows in flight, 4 warps per CTA: the other combinations measured slower, see
// profiles/README.md)
template <typename T, int NOFF>
static int launch_tile2w_n(const Tile2wParams<T>& p, int64_t nitems, int* grid_out, cudaStream_t st) {
    constexpr int PF = 2, WARPS = kT2wWarps;
    // registers: at most 64 per thread for fp32 (32 resident warps per SM), 128 for fp64
    constexpr int MINB = (sizeof(T) == 4 ? 1024 : 512) / (32 * WARPS);
    const int grid = (int)((nitems + WARPS - 1) / WARPS);
    const size_t smem = t2w_smem_bytes<T>(WARPS, p.H0);
    *grid_out = grid;
    static size_t smem_set = 48 * 1024;  // per instantiation
    if (smem > smem_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_tile2w<T, NOFF, PF, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        smem_set = smem;
    }
    k_tile2w<T, NOFF, PF, WARPS, MINB><<<grid, 32 * WARPS, smem, st>>>(p);
    ODIL_LAUNCHED();
    return 0;
}

template <typename T>
static int launch_tile2w(const odil_b200_plan* plan, const T* U, const T* c, T scale, T* G, T* Fout, int* nparts,
                         double* sumsq, cudaStream_t st) {
    Tile2wParams<T> p;
    p.sumsq = sumsq;  // the last CTA to finish sums the partials (fixed order): no separate reduction launch
    p.counter = reinterpret_cast<unsigned*>(plan->partials + kPartialCapacity);
    p.U = U;
    p.c = c;
    p.G = G;
    p.Fout = Fout;
    p.table = (const T*)plan->table_dev;
    p.partials = plan->partials;
    p.scale = scale;
    p.N0 = (int)plan->shape[0];
    p.N1 = (int)plan->shape[1];
    p.R0 = plan->R[0];
    p.R1 = plan->R[1];
    p.H0 = plan->h2[0];
    p.H1 = plan->h2[1];
    p.noff = plan->noff;
    p.ncls = plan->ncls;
    auto fdiv4 = [](int v) { return v >= 0 ? v / 4 : -((-v + 3) / 4); };  // floor(v / 4)
    for (int o = 0; o < kT2wN; ++o) {
        p.dy[o] = o < plan->noff ? (signed char)plan->off[o][0] : 0;
        p.dx[o] = o < plan->noff ? (signed char)plan->off[o][1] : 0;
        for (int j = 0; j < kT2wVW; ++j) {
            const int a = j + p.dx[o], b = j - p.dx[o];
            p.kU[o][j] = ((a - 4 * fdiv4(a)) * kT2wW + fdiv4(a)) * (int)sizeof(T);
            p.kF[o][j] = ((b - 4 * fdiv4(b)) * kT2wW + fdiv4(b)) * (int)sizeof(T);
        }
        p.rU[o] = p.dy[o] - p.H0;
        p.rF[o] = -p.H0 - p.dy[o];
    }
    auto env_int = [](const char* name, int dflt) {
        const char* e = getenv(name);
        return e ? atoi(e) : dflt;
    };
    // A warp's march is a serial chain (~8 cycles per instruction when it runs alone), so the sweep wants MANY warps:
    // about four waves of 148 SMs x 32 resident warps, in chunks of max(4, 4*H0) .. 64 rows (each pays 2*H0 lead-in
    // rows).  Measured (profiles/README.md): 1024^2 star -> 4 rows, 2048 x 4096 wave footprint -> 8, 4096^2 star -> 8
    p.nstrips = (p.N1 + kT2wOwn - 1) / kT2wOwn;
    const int64_t want = 4 * 148 * 32;
    int rc = (int)(((int64_t)p.N0 * p.nstrips + want - 1) / want);
    rc = std::min(64, std::max(std::max(4, 4 * p.H0), (rc + 3) / 4 * 4));
    rc = std::max(1, env_int("ODIL_B200_T2W_ROWS", rc));  // measurement knob
    int64_t nitems = (int64_t)p.nstrips * ((p.N0 + rc - 1) / rc);
    while ((nitems + kT2wWarps - 1) / kT2wWarps > kPartialCapacity) {
        rc *= 2;
        nitems = (int64_t)p.nstrips * ((p.N0 + rc - 1) / rc);
    }
    p.rows_per_chunk = rc;
    p.nitems = (int)nitems;
    int grid = 0, rcode = 0;
    switch (plan->noff) {
#define ODIL_T2W(N_) case N_: rcode = launch_tile2w_n<T, N_>(p, nitems, &grid, st); break
        ODIL_T2W(1); ODIL_T2W(2); ODIL_T2W(3); ODIL_T2W(4); ODIL_T2W(5); ODIL_T2W(6); ODIL_T2W(7); ODIL_T2W(8);
#undef ODIL_T2W
        default: return fail("k_tile2w: %d offsets", plan->noff);
    }
    if (rcode) return rcode;
    if (nparts) *nparts = grid;
    return 0;
}


// ------------------------------------------------------------------------------------------------
// 3-D marching tile kernel for non-star plans (tile3d.cuh)
// ------------------------------------------------------------------------------------------------
static int tile3d_zchunk(const odil_b200_plan* plan) {
    if (plan->zchunk > 0) return (int)std::min<int64_t>(plan->zchunk, plan->shape[0]);
    // enough CTAs for a few waves of 148 SMs x 2-4 resident CTAs, chunks long enough to hide the 4*H0 lead-in planes
    const int64_t tiles = ((plan->shape[1] + kT3Y - 1) / kT3Y) * ((plan->shape[2] + kT3X - 1) / kT3X);
    int64_t zc = 64;
    while (zc > 16 && tiles * ((plan->shape[0] + zc - 1) / zc) < 148 * 8) zc /= 2;
    return (int)std::min<int64_t>(zc, plan->shape[0]);
}

static bool tile3d_ok(const odil_b200_plan* plan, const odil_b200_slab* slab) {
    if (!plan->use_tile3d || plan->ndim != 3 || plan->kind != 0) return false;
    if (slab->halo > 0 || slab->n0 != plan->shape[0] || slab->z0 != 0) return false;
    for (int a = 0; a < 3; ++a)
        if (plan->h3[a] > kT3MaxRadius || plan->shape[a] < 1 || plan->shape[a] >= (1 << 30)) return false;
    if (plan->ncls > 255 || plan->shape[1] * plan->shape[2] >= (1ll << 31)) return false;
    const int64_t gx = (plan->shape[2] + kT3X - 1) / kT3X, gy = (plan->shape[1] + kT3Y - 1) / kT3Y;
    const int zc = tile3d_zchunk(plan);
    const int64_t gz = (plan->shape[0] + zc - 1) / zc;
    return gy <= 65535 && gz <= 65535 && gx * gy * gz <= kPartialCapacity;
}

// k_tile3t (tile3t.cuh): TMA-fed version for wrap-free plans with at most 8 offsets; ODIL_B200_TILE3T=0 (read per
// call, so that tests can compare the two) keeps k_tile3d.
template <typename T>
static bool tile3t_ok(const odil_b200_plan* plan, const T* U) {
    const char* e = getenv("ODIL_B200_TILE3T");
    if (e && atoi(e) == 0) return false;
    return plan->wrap_free && plan->noff <= kT3tN && (plan->shape[2] * sizeof(T)) % 16 == 0 &&
           (uintptr_t)U % 16 == 0 && get_encode_tiled() != nullptr;
}

// z-chunk of k_tile3t.  The kernel is bound by instruction issue when the SMs are full (ncu: 50 % of the issue slots
// with 2 CTAs per SM), so an SM's time is the SUM of the planes of the CTAs it runs, and by the per-plane barrier when
// they are not: the chunk count minimises the larger of the two estimates, counting the 4*H0 lead-in planes.
static int tile3t_zchunk(const odil_b200_plan* plan, int ctas_per_sm) {
    const int N0 = (int)plan->shape[0];
    if (plan->zchunk > 0) return std::min(plan->zchunk, N0);
    const int64_t tiles = ((plan->shape[1] + kT3tY - 1) / kT3tY) * ((plan->shape[2] + kT3tX - 1) / kT3tX);
    (void)ctas_per_sm;
    const int64_t slots = 148;
    const int lead = 4 * plan->h3[0];
    int best_zc = N0;
    double best = -1.0;
    for (int gz = 1; gz <= std::max(1, N0 / 8); ++gz) {
        const int zc = (N0 + gz - 1) / gz;
        const int gzr = (N0 + zc - 1) / zc;
        if (gzr > 65535 || tiles * gzr > kPartialCapacity) break;
        // measured at 256 x 512 x 512 and 128 x 256 x 256 (fp32, wave footprint): a plane costs ~1.13 us of an SM's
        // issue slots and a pair of co-resident CTAs advances one plane per ~2.2 us (barrier / latency)
        const int64_t ctas = tiles * gzr;
        const double cost = std::max((double)((ctas + slots - 1) / slots) * 1.13, (double)((ctas + 2 * slots - 1) / (2 * slots)) * 2.2) *
                            (zc + lead);
        if (best < 0 || cost < best * 0.999) {
            best = cost;
            best_zc = zc;
        }
    }
    return best_zc;
}

template <typename T>
static int launch_tile3t(const odil_b200_plan* plan, const T* U, const T* c, T scale, T* G, T* Fout, int* nparts,
                         cudaStream_t st) {
    Tile3tParams<T> p;
    p.c = c;
    p.G = G;
    p.Fout = Fout;
    p.table = (const T*)plan->table_dev;
    p.partials = plan->partials;
    p.scale = scale;
    p.N0 = (int)plan->shape[0];
    p.N1 = (int)plan->shape[1];
    p.N2 = (int)plan->shape[2];
    p.R0 = plan->R[0];
    p.R1 = plan->R[1];
    p.R2 = plan->R[2];
    p.H0 = plan->h3[0];
    p.H1 = plan->h3[1];
    p.H2 = plan->h3[2];
    p.noff = plan->noff;
    p.ncls = plan->ncls;
    const Tile3tDims d = t3t_dims<T>(p.H0, p.H1, p.ncls, p.noff);
    ODIL_REQUIRE(d.total <= 227 * 1024, "tile3t: %zu bytes of shared memory needed", d.total);
    for (int o = 0; o < kT3tN; ++o) {
        p.dz[o] = o < plan->noff ? (signed char)plan->off[o][0] : 0;
        p.dy[o] = o < plan->noff ? (signed char)plan->off[o][1] : 0;
        p.dx[o] = o < plan->noff ? (signed char)plan->off[o][2] : 0;
    }
    static size_t smem_set = 0;  // per template instantiation
    static int occ = 1;
    if (d.total > smem_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_tile3t<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.total));
        smem_set = d.total;
    }
    ODIL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_tile3t<T>, kT3tThreads, d.total));
    p.zchunk = tile3t_zchunk(plan, occ);
    CUtensorMap tmU;
    if (int rc = make_plane_map<T>(&tmU, U, p.N0, p.N1, p.N2, d.AH, kT3tAW)) return rc;
    dim3 grid((p.N2 + kT3tX - 1) / kT3tX, (p.N1 + kT3tY - 1) / kT3tY, (p.N0 + p.zchunk - 1) / p.zchunk);
    k_tile3t<T><<<grid, kT3tThreads, d.total, st>>>(tmU, p);
    ODIL_LAUNCHED();
    *nparts = (int)(grid.x * grid.y * grid.z);
    return 0;
}

template <typename T>
static int launch_tile3d(const odil_b200_plan* plan, const T* U, const T* c, T scale, T* G, T* Fout, int* nparts,
                         cudaStream_t st) {
    if (tile3t_ok<T>(plan, U)) return launch_tile3t<T>(plan, U, c, scale, G, Fout, nparts, st);
    Tile3Params<T> p;
    p.U = U;
    p.c = c;
    p.G = G;
    p.Fout = Fout;
    p.table = (const T*)plan->table_dev;
    p.partials = plan->partials;
    p.scale = scale;
    p.N0 = (int)plan->shape[0];
    p.N1 = (int)plan->shape[1];
    p.N2 = (int)plan->shape[2];
    p.R0 = plan->R[0];
    p.R1 = plan->R[1];
    p.R2 = plan->R[2];
    p.H0 = plan->h3[0];
    p.H1 = plan->h3[1];
    p.H2 = plan->h3[2];
    p.noff = plan->noff;
    p.ncls = plan->ncls;
    p.zchunk = tile3d_zchunk(plan);
    const Tile3Dims d = t3_dims(p.H0, p.H1, p.H2);
    p.magicA = (unsigned)((1ull << 32) / (unsigned)d.AW + 1);
    p.magicF = (unsigned)((1ull << 32) / (unsigned)d.FW + 1);
    for (int o = 0; o < ODIL_B200_MAX_OFFSETS; ++o) {
        p.dz[o] = o < plan->noff ? (signed char)plan->off[o][0] : 0;
        p.dy[o] = o < plan->noff ? (signed char)plan->off[o][1] : 0;
        p.dx[o] = o < plan->noff ? (signed char)plan->off[o][2] : 0;
    }
    const size_t smem = t3_smem_bytes<T>(p.H0, p.H1, p.H2, p.ncls, p.noff);
    ODIL_REQUIRE(smem <= 227 * 1024, "tile3d: %zu bytes of shared memory needed", smem);
    static size_t smem_set = 48 * 1024;  // per template instantiation
    if (smem > smem_set) {
        ODIL_CUDA(cudaFuncSetAttribute(k_tile3d<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ODIL_CUDA(cudaFuncSetAttribute(k_tile3d8<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    dim3 grid((p.N2 + kT3X - 1) / kT3X, (p.N1 + kT3Y - 1) / kT3Y, (p.N0 + p.zchunk - 1) / p.zchunk);
    // k_tile3d8 measured SLOWER than k_tile3d on B200 (2.03 vs 1.50 ms at 256 x 512 x 512): opt-in for comparison only
    static const bool unrolled = getenv("ODIL_B200_TILE3D8") != nullptr;
    if (p.noff <= kT3N && unrolled)
        k_tile3d8<T><<<grid, kT3Threads, smem, st>>>(p);  // unrolled offsets, interior cells from registers
    else
        k_tile3d<T><<<grid, kT3Threads, smem, st>>>(p);
    ODIL_LAUNCHED();
    *nparts = (int)(grid.x * grid.y * grid.z);
    return 0;
}

template <typename T>
static int run_fused(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U, const void* c,
                     double scale, void* G, void* Fout, double* sumsq, cudaStream_t st) {
    GenParams gp;
    fill_gen_params(plan, slab, gp);
    GenIO<T> io;
    io.U = (const T*)U;
    io.c = (const T*)c;
    io.out = (T*)G;
    io.Fout = (T*)Fout;
    io.table = (const T*)plan->table_dev;
    io.scale = (T)scale;
    int nparts = 0;
    if (tile2d_ok(plan, slab)) {
        {  // ODIL_B200_TILE2W=2: never fall back silently (tests)
            const char* e = getenv("ODIL_B200_TILE2W");
            if (e && atoi(e) == 2 && !tile2w_fits<T>(plan, io.U, io.c, io.out, io.Fout))
                return fail("ODIL_B200_TILE2W=2: this plan does not fit k_tile2w (wrap-free, <= 8 offsets, radii <= 2, "
                            "row length a multiple of 4, 16-byte aligned arrays)");
        }
        if (tile2w_ok<T>(plan, io.U, io.c, io.out, io.Fout))
            return launch_tile2w<T>(plan, io.U, io.c, io.scale, io.out, io.Fout, &nparts, sumsq, st);
        if (int rc = launch_tile2d<T, 2>(plan, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)) return rc;
        k_reduce_partials<<<1, 1024, 0, st>>>(plan->partials, nparts, sumsq);
        ODIL_LAUNCHED();
        return 0;
    }
    if (tile3d_ok(plan, slab)) {
        if (int rc = launch_tile3d<T>(plan, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)) return rc;
        k_reduce_partials<<<1, 1024, 0, st>>>(plan->partials, nparts, sumsq);
        ODIL_LAUNCHED();
        return 0;
    }
    const bool slab_mode = slab->halo > 0 || slab->n0 != plan->shape[0] || slab->z0 != 0;
    bool tiled = plan->kind == 1 && !(plan->ndim == 2 && slab_mode);
    const int64_t n2 = plan->shape[plan->ndim - 1];
    const bool v3 = tiled && plan->use_v3 && (n2 % 4 == 0) && ((uintptr_t)U % 16 == 0) && ((uintptr_t)G % 16 == 0) &&
                    ((uintptr_t)c % 16 == 0) && ((uintptr_t)Fout % 16 == 0);
    bool tma = v3 && plan->use_tma && plan->wrap_free && get_encode_tiled() != nullptr;
#ifndef ODIL_B200_LEGACY
    tma = false;
#endif
    if (tma) {  // the three plane rings must fit the 227 KB of shared memory of one SM
        static const int tys[4] = {16, 8, 12, 26};
        const size_t need = 128 + (size_t)10 * (((tys[(plan->variant < 0 ? 1 : plan->variant) & 3] + 4) * (128 + 8) + 31) / 32 * 32) * sizeof(T) + 512 * sizeof(T) + 128;
        if (need > 227 * 1024) tma = false;
    }
    constexpr int VW7 = 16 / (int)sizeof(T);
    const bool star7 = tiled && plan->use_star7 && plan->wrap_free && get_encode_tiled() != nullptr && (n2 % VW7 == 0) &&
                       ((uintptr_t)U % 16 == 0) && ((uintptr_t)G % 16 == 0) && ((uintptr_t)c % 16 == 0) &&
                       ((uintptr_t)Fout % 16 == 0);
    if (star7 && plan->use_star8) {
        int rc;
#define ODIL_S8(NR_, XU_) launch_star8<T, VW7, NR_, XU_>(plan, slab, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)
        const bool xu = plan->star_xu && !plan->no_xu;
        switch (plan->variant) {
            case 1: rc = xu ? ODIL_S8(12, true) : ODIL_S8(12, false); break;
            case 2: rc = xu ? ODIL_S8(8, true) : ODIL_S8(8, false); break;
            default: rc = xu ? ODIL_S8(14, true) : ODIL_S8(14, false); break;
        }
#undef ODIL_S8
        if (rc) return rc;
#ifdef ODIL_B200_LEGACY
    } else if (star7) {
        int rc;
#define ODIL_S7(TY_, XU_) launch_star7<T, VW7, TY_, XU_>(plan, slab, io.U, io.c, io.scale, io.out, io.Fout, &nparts, st)
        const bool xu = plan->star_xu && !plan->no_xu;
        switch (plan->variant) {
            case 1: rc = xu ? ODIL_S7(16, true) : ODIL_S7(16, false); break;
            case 2: rc = xu ? ODIL_S7(8, true) : ODIL_S7(8, false); break;
            default: rc = xu ? ODIL_S7(12, true) : ODIL_S7(12, false); break;
        }
#undef ODIL_S7
        if (rc) return rc;
    } else if (tma) {
        StarTmaParams<T> sp;
        sp.G = io.out;
        sp.Fout = io.Fout;
        sp.partials = plan->partials;
        sp.table = (const T*)plan->star_table;
        if (plan->ndim == 3) {
            sp.n0 = (int)slab->n0;
            sp.N0g = (int)plan->shape[0];
            sp.z0 = (int)slab->z0;
            sp.halo = slab->halo;
            sp.N1 = (int)plan->shape[1];
            sp.N2 = (int)plan->shape[2];
            sp.R0 = plan->R[0];
            sp.R1 = plan->R[1];
            sp.R2 = plan->R[2];
        } else {
            sp.n0 = 1;
            sp.N0g = 1;
            sp.z0 = 0;
            sp.halo = 0;
            sp.N1 = (int)plan->shape[0];
            sp.N2 = (int)plan->shape[1];
            sp.R0 = 0;
            sp.R1 = plan->R[0];
            sp.R2 = plan->R[1];
        }
        for (int i = 0; i < 7; ++i) sp.w[i] = (T)plan->w[i];
        sp.scale = (T)scale;
        sp.has_c = io.c != nullptr;
        int TY = 16;
        const int TX = 128;
        const int tvar = plan->variant < 0 ? 1 : plan->variant;  // auto: TY = 8
        switch (tvar) {
            case 1: TY = 8; break;
            case 2: TY = 12; break;
            case 3: TY = 26; break;
            default: TY = 16; break;
        }
        const int gx = (sp.N2 + TX - 1) / TX, gy = (sp.N1 + TY - 1) / TY;
        int zchunk = plan->zchunk;
        if (zchunk <= 0) {
            zchunk = 128;
            while (zchunk > 16 && (int64_t)gx * gy * ((sp.n0 + zchunk - 1) / zchunk) < 148 * 2 * 4) zchunk /= 2;
        }
        if (zchunk > sp.n0) zchunk = sp.n0;
        if (zchunk < 1) zchunk = 1;
        sp.zchunk = zchunk;
        const int gz = (sp.n0 + zchunk - 1) / zchunk;
        ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity, "star grid exceeds the partials workspace");
        dim3 grid(gx, gy, gz);
        const int64_t planes_total = (int64_t)sp.n0 + 2 * sp.halo;
        const int64_t plane_elems = (int64_t)sp.N1 * sp.N2;
        CUtensorMap tmU, tmC;
        if (int rc = make_plane_map<T>(&tmU, io.U - sp.halo * plane_elems, planes_total, sp.N1, sp.N2, TY + 4, TX + 8))
            return rc;
        if (int rc = make_plane_map<T>(&tmC, (io.c ? io.c : io.U) - sp.halo * plane_elems, planes_total, sp.N1, sp.N2,
                                       TY + 4, TX + 8))
            return rc;
        int rc = 0;
        switch (tvar) {
            case 1: rc = launch_star_tma<T, 8, 128>(tmU, tmC, sp, grid, st); break;
            case 2: rc = launch_star_tma<T, 12, 128>(tmU, tmC, sp, grid, st); break;
            case 3: rc = launch_star_tma<T, 26, 128>(tmU, tmC, sp, grid, st); break;
            default: rc = launch_star_tma<T, 16, 128>(tmU, tmC, sp, grid, st); break;
        }
        if (rc) return rc;
        nparts = gx * gy * gz;
#endif  // ODIL_B200_LEGACY
    } else if (v3) {
        StarV3Params<T> sp;
        sp.U = io.U;
        sp.c = io.c;
        sp.G = io.out;
        sp.Fout = io.Fout;
        sp.partials = plan->partials;
        sp.table = (const T*)plan->star_table;
        if (plan->ndim == 3) {
            sp.n0 = slab->n0;
            sp.N0g = plan->shape[0];
            sp.z0 = slab->z0;
            sp.halo = slab->halo;
            sp.N1 = (int)plan->shape[1];
            sp.N2 = (int)plan->shape[2];
            sp.R0 = plan->R[0];
            sp.R1 = plan->R[1];
            sp.R2 = plan->R[2];
            sp.has_z = plan->w[1] != 0.0 || plan->w[2] != 0.0 || plan->star_has_z;
        } else {
            sp.n0 = 1;
            sp.N0g = 1;
            sp.z0 = 0;
            sp.halo = 0;
            sp.N1 = (int)plan->shape[0];
            sp.N2 = (int)plan->shape[1];
            sp.R0 = 0;
            sp.R1 = plan->R[0];
            sp.R2 = plan->R[1];
            sp.has_z = 0;
        }
        for (int i = 0; i < 7; ++i) sp.w[i] = (T)plan->w[i];
        sp.scale = (T)scale;
        int TY, TX;
        const int vvar = plan->variant < 0 ? 0 : plan->variant;  // auto: TY = 16
        star_v3_tile(vvar, TY, TX);
        const int gx = (sp.N2 + TX - 1) / TX, gy = (sp.N1 + TY - 1) / TY;
        int zchunk = plan->zchunk;
        if (zchunk <= 0) {
            zchunk = 128;
            while (zchunk > 16 && (int64_t)gx * gy * ((sp.n0 + zchunk - 1) / zchunk) < 148 * 2 * 4) zchunk /= 2;
        }
        if (zchunk > sp.n0) zchunk = (int)sp.n0;
        if (zchunk < 1) zchunk = 1;
        sp.zchunk = zchunk;
        const int gz = (int)((sp.n0 + zchunk - 1) / zchunk);
        ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity, "star grid exceeds the partials workspace");
        dim3 grid(gx, gy, gz);
        int rc = 0;
        switch (vvar) {
            case 1: rc = launch_star_v3<T, 8, 128>(sp, grid, st); break;
            case 2: rc = launch_star_v3<T, 12, 128>(sp, grid, st); break;
            case 3: rc = launch_star_v3<T, 26, 128>(sp, grid, st); break;
            default: rc = launch_star_v3<T, 16, 128>(sp, grid, st); break;
        }
        if (rc) return rc;
        nparts = gx * gy * gz;
#ifdef ODIL_B200_LEGACY
    } else if (tiled) {
        StarParams<T> sp;
        sp.U = io.U;
        sp.c = io.c;
        sp.G = io.out;
        sp.Fout = io.Fout;
        sp.partials = plan->partials;
        if (plan->ndim == 3) {
            sp.n0 = slab->n0;
            sp.N0g = plan->shape[0];
            sp.z0 = slab->z0;
            sp.halo = slab->halo;
            sp.N1 = (int)plan->shape[1];
            sp.N2 = (int)plan->shape[2];
            sp.R0 = plan->R[0];
            sp.R1 = plan->R[1];
            sp.R2 = plan->R[2];
            sp.has_z = plan->w[1] != 0.0 || plan->w[2] != 0.0;
        } else {
            sp.n0 = 1;
            sp.N0g = 1;
            sp.z0 = 0;
            sp.halo = 0;
            sp.N1 = (int)plan->shape[0];
            sp.N2 = (int)plan->shape[1];
            sp.R0 = 0;
            sp.R1 = plan->R[0];
            sp.R2 = plan->R[1];
            sp.has_z = 0;
        }
        sp.wc = (T)plan->w[0];
        sp.wzm = (T)plan->w[1];
        sp.wzp = (T)plan->w[2];
        sp.wym = (T)plan->w[3];
        sp.wyp = (T)plan->w[4];
        sp.wxm = (T)plan->w[5];
        sp.wxp = (T)plan->w[6];
        sp.scale = (T)scale;
        int TY, TX;
        const int wvar = plan->variant < 0 ? 2 : plan->variant;  // auto: 16 x 128 tiles
        star_tile(wvar, TY, TX);
        const int gx = (sp.N2 + TX - 1) / TX, gy = (sp.N1 + TY - 1) / TY;
        int zchunk = plan->zchunk;
        if (zchunk <= 0) {
            // aim for >= ~8 waves of 148 SMs x 4 CTAs, but keep the z-redundancy (2 extra planes) small
            zchunk = 64;
            while (zchunk > 16 && (int64_t)gx * gy * ((sp.n0 + zchunk - 1) / zchunk) < 148 * 4 * 6) zchunk /= 2;
        }
        if (zchunk > sp.n0) zchunk = (int)sp.n0;
        if (zchunk < 1) zchunk = 1;
        sp.zchunk = zchunk;
        const int gz = (int)((sp.n0 + zchunk - 1) / zchunk);
        ODIL_REQUIRE((int64_t)gx * gy * gz <= kPartialCapacity / 2, "star grid exceeds the partials workspace");
        dim3 grid(gx, gy, gz);
        const bool vec = (sp.N2 % 4 == 0) && ((uintptr_t)G % 16 == 0);
        int rc = 0;
        switch (wvar) {
            case 1: rc = launch_star_cfg<T, 16, 64, 256>(sp, grid, vec, st); break;
            case 2: rc = launch_star_cfg<T, 16, 128, 512>(sp, grid, vec, st); break;
            case 3: rc = launch_star_cfg<T, 4, 128, 128>(sp, grid, vec, st); break;
            default: rc = launch_star_cfg<T, 8, 128, 256>(sp, grid, vec, st); break;
        }
        if (rc) return rc;
        nparts = gx * gy * gz;
        // Boundary shell: per-cell class lookup, overwrites g within 2r of a face, adds the loss
        // of the cells whose own row is a boundary row.
        BoxList shell;
        int thick[ODIL_B200_MAX_NDIM];
        for (int a = 0; a < plan->ndim; ++a) thick[a] = 2 * plan->R[a];
        shell_boxes(plan, slab, thick, shell);
        gp.count_mode = 1;
        io.partials = plan->partials + nparts;
        int nb = 0;
        rc = launch_generic<T, 2>(plan, gp, shell, io, st, &nb);
        if (rc) return rc;
        ODIL_REQUIRE(nparts + nb <= kPartialCapacity, "partials workspace overflow");
        nparts += nb;
#endif  // ODIL_B200_LEGACY
    } else {
        BoxList all;
        whole_box(plan, slab, all);
        gp.count_mode = 0;
        io.partials = plan->partials;
        int nb = 0;
        int rc = launch_generic<T, 2>(plan, gp, all, io, st, &nb);
        if (rc) return rc;
        nparts = nb;
    }
    k_reduce_partials<<<1, 1024, 0, st>>>(plan->partials, nparts, sumsq);
    ODIL_LAUNCHED();
    return 0;
}

template <typename T, int VW, int NR>
static int worklist_for(int n0, int N1, int N2, int zchunk, int32_t* out, int cap) {
    using Cfg = Star8Cfg<T, VW, NR>;
    std::vector<S8Work> work;
    build_work8(NR, Cfg::TX, 148 * Cfg::CTAS_PER_SM, n0, N1, N2, zchunk, work);
    for (size_t i = 0; i < work.size() && (int)i < cap; ++i) {
        out[5 * i + 0] = work[i].tx0;
        out[5 * i + 1] = work[i].ty0;
        out[5 * i + 2] = work[i].nrows;
        out[5 * i + 3] = work[i].zs;
        out[5 * i + 4] = work[i].ze;
    }
    return (int)work.size();
}

}  // namespace odil

extern "C" {

int odil_b200_version(void) { return 100; }
const char* odil_b200_last_error(void) { return last_error_ref().c_str(); }
int64_t odil_b200_launch_count(void) { return launch_counter().load(); }

int odil_b200_stencil_plan_create(int ndim, const int64_t* shape, int dtype, int noff, const int32_t* offsets,
                                  const int32_t* rwidth, const double* table, odil_b200_plan** out) {
    ODIL_REQUIRE(out != nullptr, "plan out pointer is null");
    *out = nullptr;
    ODIL_REQUIRE(ndim >= 1 && ndim <= ODIL_B200_MAX_NDIM, "ndim=%d unsupported (1..%d)", ndim, ODIL_B200_MAX_NDIM);
    ODIL_REQUIRE(noff >= 1 && noff <= ODIL_B200_MAX_OFFSETS, "noff=%d unsupported (1..%d)", noff,
                 ODIL_B200_MAX_OFFSETS);
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    odil_b200_plan* p = new odil_b200_plan();
    p->ndim = ndim;
    p->dtype = dtype;
    p->noff = noff;
    p->ncls = 1;
    p->zchunk = 0;
    p->variant = -1;  // auto: best measured tile per kernel at 512^3 fp32 on B200 (tools/bench_kernels.py)
    p->table_dev = nullptr;
    p->partials = nullptr;
    int64_t total = 1;
    for (int a = 0; a < ndim; ++a) {
        p->shape[a] = shape[a];
        p->R[a] = rwidth[a];
        if (shape[a] < 1 || rwidth[a] < 0 || 2 * (int64_t)rwidth[a] > shape[a]) {
            delete p;
            return fail("axis %d: shape=%lld rwidth=%d invalid (need shape >= 2*rwidth)", a, (long long)shape[a],
                        rwidth[a]);
        }
        p->ncls *= 2 * rwidth[a] + 1;
        total *= shape[a];
    }
    p->rmax0 = 0;
    for (int o = 0; o < noff; ++o)
        for (int a = 0; a < ndim; ++a) {
            p->off[o][a] = offsets[o * ndim + a];
            // The kernels wrap an index with ONE conditional add / subtract, so |offset| must stay below the
            // axis size on every axis, size-1 axes included (the host folds offsets modulo the size first).
            if (std::abs(p->off[o][a]) >= shape[a] && p->off[o][a] != 0) {
                delete p;
                return fail("offset %d axis %d = %d exceeds the grid size %lld", o, a, p->off[o][a],
                            (long long)shape[a]);
            }
            if (a == 0) p->rmax0 = std::max(p->rmax0, std::abs(p->off[o][a]));
        }
    p->table.assign(table, table + (size_t)p->ncls * noff);
    // wrap_free: every (class, offset) whose neighbour would cross the periodic boundary has a zero coefficient.
    {
        bool wf = true;
        std::vector<int> cstr(ndim);
        int cs2 = 1;
        for (int a = ndim - 1; a >= 0; --a) {
            cstr[a] = cs2;
            cs2 *= 2 * p->R[a] + 1;
        }
        for (int cl = 0; cl < p->ncls && wf; ++cl)
            for (int o = 0; o < noff && wf; ++o) {
                if (p->table[(size_t)cl * noff + o] == 0.0) continue;
                for (int a = 0; a < ndim; ++a) {
                    const int d = p->off[o][a];
                    if (d == 0) continue;
                    const int r = p->R[a];
                    const int c = (cl / cstr[a]) % (2 * r + 1);
                    bool crosses;
                    if (c < r)
                        crosses = c + d < 0;            // low row i = c
                    else if (c > r)
                        crosses = (2 * r - c) < d;      // high row i = N-1-(2r-c)
                    else
                        crosses = std::abs(d) > r;      // interior rows next to the boundary rows
                    if (crosses) {
                        wf = false;
                        break;
                    }
                }
            }
        p->wrap_free = wf ? 1 : 0;
    }
    // Tiled eligibility: 2-D / 3-D, every offset a unit star arm, grid not degenerate.
    p->kind = 0;
    p->star_table = nullptr;
    p->star_has_z = 0;
    p->use_v3 = 1;
    p->use_tma = 1;
    p->use_star7 = 1;
    p->star_xu = 0;
    p->no_xu = 0;
    p->use_star8 = 1;
    {
        const char* e = getenv("ODIL_B200_TILE2D");
        p->use_tile2d = !(e && e[0] == '0');
        p->h2[0] = p->h2[1] = 0;
        if (ndim == 2)
            for (int o = 0; o < noff; ++o)
                for (int a = 0; a < 2; ++a) p->h2[a] = std::max(p->h2[a], std::abs(p->off[o][a]));
        const char* e3 = getenv("ODIL_B200_TILE3D");
        p->use_tile3d = !(e3 && e3[0] == '0');
        p->h3[0] = p->h3[1] = p->h3[2] = 0;
        if (ndim == 3)
            for (int o = 0; o < noff; ++o)
                for (int a = 0; a < 3; ++a) p->h3[a] = std::max(p->h3[a], std::abs(p->off[o][a]));
    }
    p->work_dev = nullptr;
    p->work_cap = 0;
    p->work_n = 0;
    for (int i = 0; i < 6; ++i) p->work_key[i] = -1;
    std::vector<double> star;
    for (int i = 0; i < 7; ++i) p->w[i] = 0.0;
    if ((ndim == 3 || ndim == 2) && total >= 512) {
        bool ok = true;
        int cint = 0, cs = 1;
        for (int a = ndim - 1; a >= 0; --a) {
            cint += p->R[a] * cs;
            cs *= 2 * p->R[a] + 1;
            if (shape[a] < 4 * p->R[a] + 1 || shape[a] < 4) ok = false;
        }
        double w[7] = {0, 0, 0, 0, 0, 0, 0};
        const int base = ndim == 3 ? 0 : 1;  // axis a of the plan maps to tiled axis base + a
        for (int o = 0; o < noff && ok; ++o) {
            int nz = 0, ax = -1, sg = 0;
            for (int a = 0; a < ndim; ++a)
                if (p->off[o][a] != 0) {
                    ++nz;
                    ax = a;
                    sg = p->off[o][a];
                }
            const double v = p->table[(size_t)cint * noff + o];
            if (nz == 0)
                w[0] += v;
            else if (nz == 1 && (sg == 1 || sg == -1))
                w[1 + 2 * (base + ax) + (sg > 0 ? 1 : 0)] += v;
            else
                ok = false;
        }
        if (ok) {
            p->kind = 1;
            for (int i = 0; i < 7; ++i) p->w[i] = w[i];
            star.assign((size_t)p->ncls * 7, 0.0);
            for (int o = 0; o < noff; ++o) {
                int slot = 0;
                for (int a = 0; a < ndim; ++a)
                    if (p->off[o][a] != 0) slot = 1 + 2 * (base + a) + (p->off[o][a] > 0 ? 1 : 0);
                for (int cl = 0; cl < p->ncls; ++cl) {
                    star[(size_t)cl * 7 + slot] += p->table[(size_t)cl * noff + o];
                    if ((slot == 1 || slot == 2) && p->table[(size_t)cl * noff + o] != 0.0) p->star_has_z = 1;
                }
            }
            // x-uniform arms: for the interior classes along every axis but the last, the four z/y arm
            // coefficients are the same for every class along the last axis
            {
                const int C2 = 2 * p->R[ndim - 1] + 1;
                const int C1 = ndim >= 2 ? 2 * p->R[ndim - 2] + 1 : 1;
                const int cz_int = ndim == 3 ? p->R[0] : 0;  // interior class of axis 0 (3-D only)
                bool xu = true;
                for (int cy = 0; cy < C1 && xu; ++cy) {
                    const int rowbase = (cz_int * C1 + cy) * C2;
                    for (int cx = 1; cx < C2 && xu; ++cx)
                        for (int sl = 1; sl <= 4; ++sl)
                            if (star[(size_t)(rowbase + cx) * 7 + sl] != star[(size_t)rowbase * 7 + sl]) xu = false;
                }
                p->star_xu = xu ? 1 : 0;
            }
        }
    }
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) {
        const size_t esz = dtype == ODIL_B200_F32 ? 4 : 8;
        e = cudaMalloc(&p->table_dev, esz * p->table.size());
        if (e == cudaSuccess) {
            if (dtype == ODIL_B200_F32) {
                std::vector<float> tf(p->table.begin(), p->table.end());
                e = cudaMemcpy(p->table_dev, tf.data(), esz * tf.size(), cudaMemcpyHostToDevice);
            } else {
                e = cudaMemcpy(p->table_dev, p->table.data(), esz * p->table.size(), cudaMemcpyHostToDevice);
            }
        }
        // + 2 doubles behind the partials: the "blocks done" counter of kernels that reduce their own partials (k_tile2w)
        if (e == cudaSuccess) e = cudaMalloc((void**)&p->partials, sizeof(double) * (kPartialCapacity + 2));
        if (e == cudaSuccess) e = cudaMemset(p->partials + kPartialCapacity, 0, 2 * sizeof(double));
        if (e == cudaSuccess && !star.empty()) {
            e = cudaMalloc(&p->star_table, esz * star.size());
            if (e == cudaSuccess) {
                if (dtype == ODIL_B200_F32) {
                    std::vector<float> tf(star.begin(), star.end());
                    e = cudaMemcpy(p->star_table, tf.data(), esz * tf.size(), cudaMemcpyHostToDevice);
                } else {
                    e = cudaMemcpy(p->star_table, star.data(), esz * star.size(), cudaMemcpyHostToDevice);
                }
            }
        }
    }
    if (e != cudaSuccess) {
        if (p->table_dev) cudaFree(p->table_dev);
        if (p->partials) cudaFree(p->partials);
        if (p->star_table) cudaFree(p->star_table);
        delete p;
        return fail("plan_create: %s", cudaGetErrorString(e));
    }
    *out = p;
    return 0;
}

int odil_b200_stencil_plan_destroy(odil_b200_plan* plan) {
    if (!plan) return 0;
    if (plan->table_dev) cudaFree(plan->table_dev);
    if (plan->partials) cudaFree(plan->partials);
    if (plan->star_table) cudaFree(plan->star_table);
    if (plan->work_dev) cudaFree(plan->work_dev);
    delete plan;
    return 0;
}

int odil_b200_stencil_plan_kind(const odil_b200_plan* plan) { return plan ? plan->kind : -1; }

int odil_b200_star_worklist(int dtype, int variant, int64_t n0, int64_t N1, int64_t N2, int zchunk, int32_t* out, int cap) {
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    ODIL_REQUIRE(n0 >= 1 && N1 >= 1 && N2 >= 1 && n0 < (1 << 30) && N1 < (1 << 30) && N2 < (1 << 30), "bad shape");
    ODIL_REQUIRE(cap >= 0 && (out != nullptr || cap == 0), "bad output buffer");
    const int v = variant < 0 ? 0 : variant % 10;
    const int a = (int)n0, b = (int)N1, c = (int)N2;
    if (dtype == ODIL_B200_F32) {
        if (v == 1) return worklist_for<float, 4, 12>(a, b, c, zchunk, out, cap);
        if (v == 2) return worklist_for<float, 4, 8>(a, b, c, zchunk, out, cap);
        return worklist_for<float, 4, 14>(a, b, c, zchunk, out, cap);
    }
    if (v == 1) return worklist_for<double, 2, 12>(a, b, c, zchunk, out, cap);
    if (v == 2) return worklist_for<double, 2, 8>(a, b, c, zchunk, out, cap);
    return worklist_for<double, 2, 14>(a, b, c, zchunk, out, cap);
}

int odil_b200_stencil_plan_tune(odil_b200_plan* plan, int zchunk, int variant) {
    ODIL_REQUIRE(plan != nullptr, "null plan");
    ODIL_REQUIRE((variant >= 0 && variant <= 3) || (variant >= 10 && variant <= 13) ||
                     (variant >= 20 && variant <= 23) || (variant >= 30 && variant <= 32) ||
                     (variant >= 40 && variant <= 42) || (variant >= 50 && variant <= 52) ||
                     (variant >= 60 && variant <= 62) || variant == 70 || variant == 71 || variant == 80 ||
                     variant == 81 || variant == -1,
                 "variant=%d unknown", variant);
#ifndef ODIL_B200_LEGACY
    ODIL_REQUIRE(!((variant >= 0 && variant <= 3) || (variant >= 10 && variant <= 13) || (variant >= 30 && variant <= 42)),
                 "variant=%d selects a superseded kernel generation (k_star_tma / k_star3d / k_star7): rebuild with "
                 "ODIL_B200_LEGACY=1", variant);
#endif
    plan->zchunk = zchunk;
    // 70 / 71: 2-D tile kernel on / off with the default 3-D choice; any explicit star variant also turns it off
    // (so the star kernels stay reachable on 2-D grids)
    if (variant == 80 || variant == 81) {  // 3-D marching tile kernel for non-star plans on / off
        plan->use_tile3d = variant == 80;
        return 0;
    }
    plan->use_tile2d = variant == 70 || variant == -1;
    if (variant >= 70) variant = -1;
    // -1 / 50..52: k_star8 (default; up to 14 / 12 / 8 rows per CTA); 60..62: same with per-cell z/y-arm
    // coefficients even when the plan allows the x-uniform form; 30..32 / 40..42: k_star7 (tiles of 12 / 16 / 8
    // rows, x-uniform / per-cell arms); 0..3: previous TMA-fed kernel; 10..13: v2 tile kernel + shell pass;
    // 20..23: LDG column-group kernel
    if (variant < 0) variant = 50;
    plan->use_star8 = variant >= 50;
    plan->use_star7 = variant >= 30;
    plan->no_xu = (variant >= 40 && variant < 50) || variant >= 60;
    plan->use_tma = variant < 10;
    plan->use_v3 = variant < 10 || variant >= 20;
    plan->variant = variant % 10;
    for (int i = 0; i < 6; ++i) plan->work_key[i] = -1;
    return 0;
}

static int check_slab(const odil_b200_plan* plan, const odil_b200_slab* slab, int need_halo) {
    ODIL_REQUIRE(plan && slab, "null plan/slab");
    ODIL_REQUIRE(slab->n0 >= 1 && slab->z0 >= 0 && slab->z0 + slab->n0 <= plan->shape[0],
                 "slab [%lld,+%lld) outside axis 0 of size %lld", (long long)slab->z0, (long long)slab->n0,
                 (long long)plan->shape[0]);
    if (slab->halo == 0)
        ODIL_REQUIRE(slab->n0 == plan->shape[0], "halo=0 requires the whole axis 0 (single-GPU semantics)");
    else
        ODIL_REQUIRE(slab->halo >= need_halo, "halo=%d too small, need %d", slab->halo, need_halo);
    return 0;
}

int odil_b200_stencil_forward(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U,
                              const void* F_in, void* F_out, void* stream) {
    if (int rc = check_slab(plan, slab, plan ? plan->rmax0 : 0)) return rc;
    ODIL_REQUIRE(U && F_out, "null array");
    if (tile2d_ok(plan, slab)) {
        if (plan->dtype == ODIL_B200_F32)
            return launch_tile2d<float, 0>(plan, (const float*)U, (const float*)F_in, 1.f, (float*)F_out, nullptr,
                                           nullptr, (cudaStream_t)stream);
        return launch_tile2d<double, 0>(plan, (const double*)U, (const double*)F_in, 1.0, (double*)F_out, nullptr,
                                        nullptr, (cudaStream_t)stream);
    }
    GenParams gp;
    fill_gen_params(plan, slab, gp);
    BoxList all;
    whole_box(plan, slab, all);
    if (plan->dtype == ODIL_B200_F32) {
        GenIO<float> io{(const float*)U, (const float*)F_in, (float*)F_out, nullptr, (const float*)plan->table_dev,
                        nullptr, 1.f};
        return launch_generic<float, 0>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
    }
    GenIO<double> io{(const double*)U, (const double*)F_in, (double*)F_out, nullptr, (const double*)plan->table_dev,
                     nullptr, 1.0};
    return launch_generic<double, 0>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
}

int odil_b200_stencil_adjoint(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* F, double scale,
                              const void* G_in, void* G_out, void* stream) {
    if (int rc = check_slab(plan, slab, plan ? plan->rmax0 : 0)) return rc;
    ODIL_REQUIRE(F && G_out, "null array");
    if (tile2d_ok(plan, slab)) {
        if (plan->dtype == ODIL_B200_F32)
            return launch_tile2d<float, 1>(plan, (const float*)F, (const float*)G_in, (float)scale, (float*)G_out,
                                           nullptr, nullptr, (cudaStream_t)stream);
        return launch_tile2d<double, 1>(plan, (const double*)F, (const double*)G_in, scale, (double*)G_out, nullptr,
                                        nullptr, (cudaStream_t)stream);
    }
    GenParams gp;
    fill_gen_params(plan, slab, gp);
    BoxList all;
    whole_box(plan, slab, all);
    if (plan->dtype == ODIL_B200_F32) {
        GenIO<float> io{(const float*)F, (const float*)G_in, (float*)G_out, nullptr, (const float*)plan->table_dev,
                        nullptr, (float)scale};
        return launch_generic<float, 1>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
    }
    GenIO<double> io{(const double*)F, (const double*)G_in, (double*)G_out, nullptr, (const double*)plan->table_dev,
                     nullptr, scale};
    return launch_generic<double, 1>(plan, gp, all, io, (cudaStream_t)stream, nullptr);
}

int odil_b200_stencil_fused(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U, const void* c,
                            double scale, void* G_out, void* F_out, double* sumsq_out, void* stream) {
    if (int rc = check_slab(plan, slab, plan ? 2 * plan->rmax0 : 0)) return rc;
    ODIL_REQUIRE(U && G_out && sumsq_out, "null array");
    if (plan->dtype == ODIL_B200_F32)
        return run_fused<float>(plan, slab, U, c, scale, G_out, F_out, sumsq_out, (cudaStream_t)stream);
    return run_fused<double>(plan, slab, U, c, scale, G_out, F_out, sumsq_out, (cudaStream_t)stream);
}

}  // extern "C"
