/*
 * odil_b200.h -- C ABI of libodil_b200.so, the B200-native (sm_100a) residual-and-gradient engine
 * behind the ODIL Python API.
 *
 * The reference (cselab/odil) has NO native code and NO FFI: its per-iteration hot path is a
 * jitted XLA/TF program reached through `Problem._eval_loss_grad` (reference
 * src/odil/core.py:1027-1036 picks the backend, :1076-1111 is the JAX program) and the jitted
 * Adam step (src/odil/optimizer.py:311-326).  This header is the seam a maintainer binds with
 * ctypes to replace those programs.  Every entry point cites the reference
 * lines whose arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; odil_b200_last_error() gives the message
 *     (thread-local).  Nothing throws, nothing synchronises the device, nothing allocates after
 *     plan creation (so calls are CUDA-graph capturable).
 *   - all data pointers are DEVICE pointers owned by the caller, C-order (axis 0 slowest),
 *     16-byte aligned; `stream` is a cudaStream_t passed as void*.
 *   - dtype: 0 = float32, 1 = float64.  Scalars cross the ABI as double and are rounded to the
 *     array dtype inside.
 *   - slabs (multi-GPU, axis-0 decomposition): an array argument points at the FIRST OWNED plane;
 *     `halo` planes exist physically before and after the owned planes (0 => none; neighbours
 *     along axis 0 then wrap periodically inside the owned range, i.e. single-GPU semantics).
 */
#ifndef ODIL_B200_H
#define ODIL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODIL_B200_MAX_NDIM 4
#define ODIL_B200_MAX_OFFSETS 32
#define ODIL_B200_F32 0
#define ODIL_B200_F64 1

typedef struct odil_b200_plan odil_b200_plan; /* opaque */

int odil_b200_version(void);
const char* odil_b200_last_error(void);
/* Number of kernels this library has launched in the calling process (bench `gpu_launches`). */
int64_t odil_b200_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Region-typed affine stencil  F = A U + c.
 *
 * Replaces: ctx.field() = roll(U, -shift) (core.py:910-975, :963), the operator's elementwise
 * arithmetic incl. where(index-mask, ...) boundary rows (examples/poisson/poisson.py:57-68,
 * :100-113; examples/wave/wave.py:29-75), the loss reduction mean(square(F)) (core.py:1093) and
 * the reverse-mode gradient (2/n) A^T F (core.py:1100-1101).
 *
 *   shape[ndim]    GLOBAL grid shape of U and F.
 *   offsets        noff x ndim stencil shifts, same sign convention as ctx.field(key, *shift):
 *                  F[x] uses U[(x + off) mod N].
 *   rwidth[ndim]   region half-width r_a: per axis the classes are rows 0..r_a-1, interior,
 *                  rows N-r_a..N-1  (2 r_a + 1 classes); requires N_a >= 2 r_a.
 *   table          prod_a(2 r_a + 1) x noff coefficients, class index row-major over axes.
 * ------------------------------------------------------------------------------------------- */
int odil_b200_stencil_plan_create(int ndim, const int64_t* shape, int dtype, int noff, const int32_t* offsets,
                                  const int32_t* rwidth, const double* table, odil_b200_plan** plan);
int odil_b200_stencil_plan_destroy(odil_b200_plan* plan);

/* Slab geometry for one call: n0 owned planes starting at global plane z0; halo as above.
 * Single GPU: n0 = shape[0], z0 = 0, halo = 0. */
typedef struct {
    int64_t n0;
    int64_t z0;
    int32_t halo;
} odil_b200_slab;

/* F_out = A U + F_in      (F_in nullable => 0; F_in may alias F_out). Owned planes only.
 * U needs halo >= max|off_0| when slab.halo > 0. */
int odil_b200_stencil_forward(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U,
                              const void* F_in, void* F_out, void* stream);

/* G_out = scale * A^T F + G_in   (G_in nullable; may alias G_out). F needs halo >= max|off_0|. */
int odil_b200_stencil_adjoint(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* F, double scale,
                              const void* G_in, void* G_out, void* stream);

/* Fused single sweep:  F = A U + c (never written to HBM),  sumsq_out[0] = sum_owned F^2 (double,
 * deterministic two-stage reduction),  G_out = scale * A^T F.
 * c nullable (=> 0).  With slab.halo > 0: U needs halo >= 2 r0, c needs halo >= r0
 * (r0 = max|off_0|).  F_out nullable: when given, F is also stored (owned planes). */
int odil_b200_stencil_fused(const odil_b200_plan* plan, const odil_b200_slab* slab, const void* U, const void* c,
                            double scale, void* G_out, void* F_out, double* sumsq_out, void* stream);

/* Inspection of the launch plan of the fused star sweep (host only, no device work): the CTAs the work-list kernel
 * is launched with for a slab of n0 planes of an (N0, N1, N2) grid -- per CTA {x origin, y origin, rows, first plane,
 * end plane} (5 int32).  Fills up to `cap` entries and returns the number of CTAs (<0 on error).  `variant` and
 * `zchunk` as in odil_b200_stencil_plan_tune (-1 / 0 = defaults).  The list tiles the slab exactly once and is sized
 * to fill the 148 SMs in whole waves with equal work per CTA. */
int odil_b200_star_worklist(int dtype, int variant, int64_t n0, int64_t N1, int64_t N2, int zchunk, int32_t* out, int cap);

/* Classification of the plan's offsets: 1 = unit-arm star on a 2-D / 3-D grid (3-D: the TMA-fed marching sweep),
 * 0 = anything else.  2-D plans of either kind run the shared-memory tile kernel (any offsets of radius <= 4) on a
 * single GPU; 1-D, 4-D, non-star 3-D plans and 2-D slabs run the per-cell kernel. */
int odil_b200_stencil_plan_kind(const odil_b200_plan* plan);
/* Tuning knobs (0 / -1 keep the defaults): planes per z-chunk of the star kernels; `variant` selects a kernel
 * generation / tile shape of the star sweep (50-52, 60-62 current; 30-42, 20-23, 10-13, 0-3 earlier ones, kept as
 * measured history), 70 / 71 switch the 2-D tile kernel on / off (any explicit star variant also switches it off,
 * so the star kernels stay reachable on 2-D grids); 80 / 81 switch the marching tile kernel for non-star 3-D plans
 * on / off (default on; ODIL_B200_TILE3D=0 at plan creation also switches it off). */
int odil_b200_stencil_plan_tune(odil_b200_plan* plan, int zchunk, int variant);

/* sumsq_out[0] = sum x^2 (double accumulate).  Replaces mean(square(f)) of a materialised F
 * (core.py:1093) and the norm used by optimizers. `count` elements. */
int odil_b200_sum_squares(const void* x, int64_t count, int dtype, double* sumsq_out, void* stream);
/* out[0] = sum x*y (double). */
int odil_b200_dot(const void* x, const void* y, int64_t count, int dtype, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multigrid transfers (core.py:245-263 synthesis, :606-700 interp_to_finer, :703-755
 * restrict_to_coarser).  `loc` is one char per axis: 'c' cell, 'n' node, '.' axis not coarsened.
 * `cshape` is the coarse ARRAY shape (global); the fine array shape follows from loc
 * ('c': 2n, 'n': 2(n-1)+1, '.': n).
 *
 * interp_add:  out = ffac * fine_term + cfac * I(coarse)      (fine_term nullable => pure interp)
 *   Slabs: computes fine planes [fz_begin, fz_end) (GLOBAL indices). `out`/`fine_term` point at
 *   global fine plane out_z0; `coarse` points at global coarse plane coarse_z0.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t fz_begin, fz_end; /* fine planes to produce (global) */
    int64_t out_z0;           /* global fine plane index at the out / fine_term pointer */
    int64_t coarse_z0;        /* global coarse plane index at the coarse pointer */
} odil_b200_mg_range;

int odil_b200_mg_interp_add(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* coarse,
                            double cfac, const void* fine_term, double ffac, void* out,
                            const odil_b200_mg_range* range /* NULL => whole array */, void* stream);

/* g_coarse = scale * I^T g_fine  (exact transpose of interp incl. the joint 2*symmetric-reflect
 * pad; this is what AD of core.py:606-700 produces, NOT restrict_to_coarser).
 * Slabs: computes coarse planes [cz_begin, cz_end) (global); g_fine points at global fine plane
 * fine_z0 and must hold planes 2*cz_begin-2 .. 2*cz_end+1 (clipped to the domain; the outermost two
 * are only read for the pad corrections of the coarse planes 1 and n-2). */
typedef struct {
    int64_t cz_begin, cz_end;
    int64_t out_z0;  /* global coarse plane at the g_coarse pointer */
    int64_t fine_z0; /* global fine plane at the g_fine pointer */
} odil_b200_mg_adj_range;

int odil_b200_mg_interp_adjoint(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* g_fine,
                                double scale, void* g_coarse, const odil_b200_mg_adj_range* range, void* stream);

/* out = restrict_to_coarser(in)  (core.py:703-755). `fshape` = fine ARRAY shape. Single GPU. */
int odil_b200_mg_restrict(int ndim, const int64_t* fshape, const char* loc, int dtype, const void* in, void* out,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer updates.
 * adam_step replaces AdamNativeOptimizer._step (optimizer.py:311-319) for all tensors at once:
 *   m += (g - m) * one_minus_beta1;  v += (g*g - v) * one_minus_beta2;
 *   x -= (m * alpha) / (sqrt(v) + epsilon)
 * `alpha`, `one_minus_beta*` are computed by the host IN THE ARRAY DTYPE (optimizer.py:307-314)
 * and passed as doubles holding exactly representable values.
 * gd_step replaces GdOptimizer (optimizer.py:269-270): x -= g * lr.
 * ------------------------------------------------------------------------------------------- */
int odil_b200_adam_step(int ntensors, void* const* x, void* const* m, void* const* v, const void* const* g,
                        const int64_t* counts, int dtype, double alpha, double one_minus_beta1,
                        double one_minus_beta2, double epsilon, void* stream);
/* Same update with the step size read from DEVICE memory (alpha_dev[0], a double holding a value exactly
 * representable in the array dtype): the launch arguments no longer change from epoch to epoch, so a whole epoch
 * (residual + gradient + transfers + this update) can be captured once into a CUDA graph and replayed
 * (optimizer.py:307-326: alpha is the only per-epoch scalar). */
int odil_b200_adam_step_dev(int ntensors, void* const* x, void* const* m, void* const* v, const void* const* g,
                            const int64_t* counts, int dtype, const double* alpha_dev, double one_minus_beta1,
                            double one_minus_beta2, double epsilon, void* stream);

/* Step-size table of a replayed run: out[0] = table[step[0]], then step[0] += 1 (one launch).  The host tabulates
 * alpha(epoch) = lr * sqrt(1 - beta2^t) / (1 - beta1^t) in the array dtype for every epoch of the run
 * (optimizer.py:307-309) once; the captured epoch starts with this call and feeds `out` to odil_b200_adam_step_dev /
 * odil_b200_adam_synth as alpha_dev. */
int odil_b200_table_pick(const double* table, int64_t* step, double* out, void* stream);

int odil_b200_gd_step(int ntensors, void* const* x, const void* const* g, const int64_t* counts, int dtype,
                      double lr, void* stream);
/* y = a*x + b*y  (L-BFGS building block; also used for scaling). */
int odil_b200_axpby(int64_t count, int dtype, double a, const void* x, double b, void* y, void* stream);

/* Conjugate-gradient vector updates with the step scalars read from DEVICE memory (num[0] / den[0]; a zero
 * denominator gives a zero step), so an iteration never waits for the host:
 *   cg_update_xr:  alpha = num/den;  x += alpha p;  r -= alpha q
 *   cg_update_p:   beta  = num/den;  p = r + beta p
 * Together with odil_b200_dot and the stencil forward / adjoint products they replace the sparse solves the
 * reference calls for Newton on the normal equations M^T M x = M^T rhs (linsolver.py:18-26, util.py:152-187). */
int odil_b200_cg_update_xr(int64_t count, int dtype, const double* num, const double* den, const void* p, const void* q,
                           void* x, void* r, void* stream);
int odil_b200_cg_update_p(int64_t count, int dtype, const double* num, const double* den, const void* r, void* p,
                          void* stream);

/* L-BFGS building blocks (compact form). V: row-major [k][ld] matrix holding the history vectors
 * (k <= 256 rows of `count` elements).  Replaces the host vector algebra inside SciPy's
 * fmin_l_bfgs_b that the reference calls (optimizer.py:95-105).
 *   multi_dot:  out[r] = sum_i V[r][i] * g[i]            (device doubles, deterministic)
 *   multi_axpy: d[i]   = a0 * g[i] + sum_r coef[r] * V[r][i]   (coef: k device doubles; g nullable) */
int odil_b200_multi_dot(const void* V, int64_t ld, int k, const void* g, int64_t count, int dtype, double* out,
                        void* stream);
int odil_b200_multi_axpy(const void* V, int64_t ld, int k, const double* coef, double a0, const void* g, void* d,
                         int64_t count, int dtype, void* stream);

/* Fusion across the optimizer seam: the transposed interpolation of the finest level streams the
 * gradient g_fine of the finest multigrid term once and applies that term's Adam update (optimizer.py:311-319) on the
 * way, instead of k_adam reading g_fine a second time.  alpha_dev (nullable): step size in device memory (CUDA-graph
 * replay).  Returns 1 without doing anything when the arrays do not fit the marching kernel (caller: unfused pair). */
int odil_b200_mg_interp_adjoint_adam(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* g_fine,
                                     double scale, void* g_coarse, void* x, void* m, void* v, double alpha,
                                     const double* alpha_dev, double one_minus_beta1, double one_minus_beta2,
                                     double epsilon, void* stream);

/* The same seam, the other way round: the Adam update of the finest multigrid term t0 (x, m, v with gradient g;
 * optimizer.py:311-319) also writes the regular field of the NEXT evaluation, out = ffac * t0_new + cfac * I(coarse)
 * (multigrid_to_regular, core.py:245-263; coarse = the synthesised level 1, cshape = its shape), so the synthesis of
 * level 0 does not read t0 back.  x, m, v and out are bit-identical to odil_b200_adam_step followed by
 * odil_b200_mg_interp_add.  range (nullable; slabs): even fine planes [fz_begin, fz_end) to update, out_z0 = global
 * number of local plane 0 of x / m / v / g / out, coarse_z0 = that of `coarse` (as in odil_b200_mg_interp_add).
 * Returns 1 without doing anything when the arrays do not fit (caller: unfused pair). */
int odil_b200_adam_synth(int ndim, const int64_t* cshape, const char* loc, int dtype, const void* coarse, double cfac,
                         double ffac, void* x, void* m, void* v, const void* g, void* out, double alpha,
                         const double* alpha_dev, double one_minus_beta1, double one_minus_beta2, double epsilon,
                         const odil_b200_mg_range* range, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Run-time specialised kernels for operators that are not affine stencils.
 *
 * Replaces: the reference's per-operator XLA program -- `jax.jit(value_and_grad(eval_loss))` (core.py:1100-1107),
 * `tf.function(jit_compile=True)` (core.py:1069-1071), the per-(key, shift) Jacobian diagonals of
 * `_eval_operator_grad_tf` (core.py:1313-1361).  The host tracer (odil_b200/graph.py, codegen.py) writes CUDA C for
 * the traced operator (residuals, loss partial sums, reverse-mode adjoint, Jacobian products); these entry points
 * compile it with NVRTC to an sm_100a cubin and launch it.
 *
 *   jit_compile   source -> module (no device needed); jit_log() returns the compiler log of the calling thread
 *   jit_cubin     the compiled image (inspection with cuobjdump, caching)
 *   jit_kernel    handle of the `extern "C" __global__` function `name` (loads the image on first use)
 *   jit_launch    1-D launch; the kernel's single by-value struct parameter is copied from params[0..nbytes)
 * ------------------------------------------------------------------------------------------- */
int odil_b200_jit_compile(const char* source, const char* const* options, int noptions, void** module_out);
const char* odil_b200_jit_log(void);
int odil_b200_jit_cubin(void* module, const void** data, uint64_t* size);
int odil_b200_jit_kernel(void* module, const char* name, int max_dynamic_smem, void** kernel_out);
int odil_b200_jit_launch(void* kernel, uint32_t grid, uint32_t block, uint32_t smem, const void* params, uint64_t nbytes,
                         void* stream);
int odil_b200_jit_destroy(void* module);

/* ---------------------------------------------------------------------------------------------
 * Slab communicator: halo exchange and scalar all-reduce between the GPUs of one node.
 *
 * The reference has no multi-device path.  One process per GPU; `comm_create` allocates this rank's communication
 * block (flags + staging) and returns its 64-byte CUDA IPC handle; the host all-gathers the handles (any transport;
 * torch.distributed here) and `comm_connect` maps the peers' blocks (NVLink / NVSwitch peer access).  The data plane
 * is then pure device work on the caller's stream (two kernels per exchange, one per all-reduce), capturable in a
 * CUDA graph.  Every rank must issue the same sequence of exchanges / all-reduces.
 *
 *   halo_exchange      ring exchange along axis 0 of `narrays` arrays: send_lo[i] / send_hi[i] = this rank's first /
 *                      last nbytes[i] owned bytes (contiguous planes), recv_lo[i] / recv_hi[i] = its lower / upper halo
 *   halo_accumulate    the transpose of halo_exchange: send_lo[i] / send_hi[i] = partial sums this rank computed for
 *                      the planes just below / above its slab; they are ADDED to the owners' last / first owned planes
 *                      (acc_hi / acc_lo on the receiving side), lower neighbour's contribution first (deterministic)
 *   allreduce_scalars  in-place sum over ranks of count <= 16 doubles in device memory, summed in rank order
 * ------------------------------------------------------------------------------------------- */
typedef struct odil_b200_comm odil_b200_comm;
int odil_b200_comm_create(int rank, int world, int64_t halo_bytes, odil_b200_comm** comm, void* ipc_handle_out);
int odil_b200_comm_connect(odil_b200_comm* comm, const void* ipc_handles_in_rank_order);
int64_t odil_b200_comm_capacity(const odil_b200_comm* comm);
int odil_b200_halo_exchange(odil_b200_comm* comm, int narrays, const void* const* send_lo, const void* const* send_hi,
                            void* const* recv_lo, void* const* recv_hi, const int64_t* nbytes, void* stream);
int odil_b200_halo_accumulate(odil_b200_comm* comm, int narrays, const void* const* send_lo, const void* const* send_hi,
                              void* const* acc_lo, void* const* acc_hi, const int64_t* nbytes, int dtype, void* stream);
int odil_b200_allreduce_scalars(odil_b200_comm* comm, double* dev_scalars, int count, void* stream);
int odil_b200_comm_destroy(odil_b200_comm* comm);

#ifdef __cplusplus
}
#endif
#endif /* ODIL_B200_H */
