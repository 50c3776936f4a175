// k_generic: region-typed affine stencil for any ndim <= 4 and any offsets, one thread per cell with a per-cell
// class lookup; forward / adjoint / fused modes on a list of boxes (whole domain, or the boundary shell after the
// first tile kernel).  Included by stencil.cu after the plan / GenParams / BoxList definitions.
#pragma once
#include "common.cuh"

namespace odil {

// ------------------------------------------------------------------------------------------------
// Generic kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int axis_class(int64_t i, int64_t n, int r) {
    if (i < r) return (int)i;
    const int64_t d = n - 1 - i;
    if (d < r) return 2 * r - (int)d;
    return r;
}

template <typename T>
struct GenIO {
    const T* U;     // forward/fused: input field; adjoint: F
    const T* c;     // fused: constant term (nullable); forward: F_in; adjoint: G_in
    T* out;         // forward: F_out; adjoint/fused: G_out
    T* Fout;        // fused: optional F store
    const T* table;
    double* partials;
    T scale;
};

// Resolves the local element offset and class index of the cell `x + sgn*off` given local coords.
struct CellRef {
    int64_t lin;
    int cls;
};

__device__ __forceinline__ CellRef neighbour(const GenParams& p, const int64_t* xc /*local coords*/, const int* off,
                                             int sgn) {
    CellRef r;
    r.lin = 0;
    r.cls = 0;
#pragma unroll
    for (int a = 0; a < ODIL_B200_MAX_NDIM; ++a) {
        if (a >= p.ndim) break;
        const int s = sgn * off[a];
        const int64_t n = p.shape[a];
        int64_t il, ig;
        if (a == 0) {
            ig = p.z0 + xc[0] + s;
            if (ig < 0) ig += n;
            if (ig >= n) ig -= n;
            if (p.halo > 0) {
                il = xc[0] + s;  // physical halo plane
            } else {
                il = xc[0] + s;
                if (il < 0) il += p.n0;
                if (il >= p.n0) il -= p.n0;
            }
        } else {
            il = xc[a] + s;
            if (il < 0) il += n;
            if (il >= n) il -= n;
            ig = il;
        }
        r.lin += il * p.stride[a];
        r.cls += axis_class(ig, n, p.R[a]) * p.cstride[a];
    }
    return r;
}

template <typename T, int MODE>  // 0 forward, 1 adjoint, 2 fused
__global__ void __launch_bounds__(256) k_generic(GenParams p, BoxList boxes, GenIO<T> io) {
    __shared__ double red[32];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc2 = 0.0;
    if (gid < boxes.start[boxes.nbox]) {
        int b = 0;
        while (gid >= boxes.start[b + 1]) ++b;
        int64_t rem = gid - boxes.start[b];
        int64_t xc[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
        for (int a = p.ndim - 1; a >= 0; --a) {
            const int64_t s = boxes.sz[b][a];
            xc[a] = boxes.lo[b][a] + rem % s;
            rem /= s;
        }
        const int zero[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
        const CellRef self = neighbour(p, xc, zero, 0);
        if (MODE == 0) {
            T f = io.c ? io.c[self.lin] : T(0);
            for (int o = 0; o < p.noff; ++o) {
                const CellRef nb = neighbour(p, xc, p.off[o], +1);
                f += io.table[self.cls * p.noff + o] * io.U[nb.lin];
            }
            io.out[self.lin] = f;
        } else if (MODE == 1) {
            T g = T(0);
            for (int o = 0; o < p.noff; ++o) {
                const CellRef nb = neighbour(p, xc, p.off[o], -1);
                g += io.table[nb.cls * p.noff + o] * io.U[nb.lin];
            }
            g *= io.scale;
            if (io.c) g += io.c[self.lin];
            io.out[self.lin] = g;
        } else {
            // F at a cell y (given by local coords yc): sum_p table[cls(y)][p] * U[y + off_p] + c[y]
            auto eval_F = [&](const int64_t* yc, const CellRef& yref) -> T {
                T f = io.c ? io.c[yref.lin] : T(0);
                for (int q = 0; q < p.noff; ++q) {
                    const CellRef nb = neighbour(p, yc, p.off[q], +1);
                    f += io.table[yref.cls * p.noff + q] * io.U[nb.lin];
                }
                return f;
            };
            T g = T(0);
            T fself = T(0);
            bool have_self = false;
            for (int o = 0; o < p.noff; ++o) {
                // y = x - off_o, in local coords with the same wrapping rule as `neighbour`
                int64_t yc[ODIL_B200_MAX_NDIM] = {0, 0, 0, 0};
                for (int a = 0; a < p.ndim; ++a) {
                    int64_t v = xc[a] - p.off[o][a];
                    if (a == 0) {
                        if (p.halo == 0) {
                            if (v < 0) v += p.n0;
                            if (v >= p.n0) v -= p.n0;
                        }
                    } else {
                        if (v < 0) v += p.shape[a];
                        if (v >= p.shape[a]) v -= p.shape[a];
                    }
                    yc[a] = v;
                }
                const CellRef yref = neighbour(p, yc, zero, 0);
                const T fy = eval_F(yc, yref);
                if (o == p.zero_off) {
                    fself = fy;
                    have_self = true;
                }
                g += io.table[yref.cls * p.noff + o] * fy;
            }
            if (!have_self) fself = eval_F(xc, self);
            io.out[self.lin] = g * io.scale;
            if (io.Fout) io.Fout[self.lin] = fself;
            bool count = true;
            if (p.count_mode == 1) {
                // interior class index = sum_a R[a]*cstride[a]
                int cint = 0;
                for (int a = 0; a < p.ndim; ++a) cint += p.R[a] * p.cstride[a];
                count = self.cls != cint;
            }
            if (count) acc2 = (double)fself * (double)fself;
        }
    }
    if (MODE == 2) {
        const double s = block_sum(acc2, red);
        if (threadIdx.x == 0) io.partials[blockIdx.x] = s;
    }
}

}  // namespace odil
