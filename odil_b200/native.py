"""
ctypes binding of libodil_b200.so (see include/odil_b200.h).  This is the ONLY compute path of the
package: there is no CPU or eager fallback -- if the library cannot be loaded, every entry point
raises.  Arguments are torch CUDA tensors (device memory plumbing) and Python scalars; work is
enqueued on torch's current stream.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ODIL_B200_LIB") or os.path.join(_HERE, "lib", "libodil_b200.so")  # override: A/B builds

F32, F64 = 0, 1
MAX_NDIM = 4
MAX_OFFSETS = 32

EXPORTS = [
    "odil_b200_version", "odil_b200_last_error", "odil_b200_launch_count",
    "odil_b200_stencil_plan_create", "odil_b200_stencil_plan_destroy", "odil_b200_stencil_forward",
    "odil_b200_stencil_adjoint", "odil_b200_stencil_fused", "odil_b200_stencil_plan_kind",
    "odil_b200_stencil_plan_tune", "odil_b200_sum_squares", "odil_b200_dot", "odil_b200_mg_interp_add",
    "odil_b200_mg_interp_adjoint", "odil_b200_mg_restrict", "odil_b200_adam_step", "odil_b200_gd_step",
    "odil_b200_axpby", "odil_b200_multi_dot", "odil_b200_multi_axpy", "odil_b200_cg_update_xr",
    "odil_b200_cg_update_p", "odil_b200_star_worklist", "odil_b200_adam_step_dev", "odil_b200_table_pick",
    "odil_b200_mg_interp_adjoint_adam", "odil_b200_adam_synth",
    "odil_b200_jit_compile", "odil_b200_jit_log", "odil_b200_jit_cubin", "odil_b200_jit_kernel",
    "odil_b200_jit_launch", "odil_b200_jit_destroy",
    "odil_b200_comm_create", "odil_b200_comm_connect", "odil_b200_comm_capacity", "odil_b200_halo_exchange",
    "odil_b200_halo_accumulate",
    "odil_b200_allreduce_scalars", "odil_b200_comm_destroy",
]


class NativeError(RuntimeError):
    pass


class Slab(ctypes.Structure):
    _fields_ = [("n0", ctypes.c_int64), ("z0", ctypes.c_int64), ("halo", ctypes.c_int32)]


class MgRange(ctypes.Structure):
    _fields_ = [("fz_begin", ctypes.c_int64), ("fz_end", ctypes.c_int64), ("out_z0", ctypes.c_int64),
                ("coarse_z0", ctypes.c_int64)]


class MgAdjRange(ctypes.Structure):
    _fields_ = [("cz_begin", ctypes.c_int64), ("cz_end", ctypes.c_int64), ("out_z0", ctypes.c_int64),
                ("fine_z0", ctypes.c_int64)]


_lib = None
_timer_hook = None


def set_timer_hook(hook):
    """hook(name, thunk) -> result: lets bench.py bracket each C-ABI call with CUDA events."""
    global _timer_hook
    _timer_hook = hook


def _call(name, thunk):
    return _timer_hook(name, thunk) if _timer_hook is not None else thunk()


def load(build_if_missing=False):
    """Loads the shared library (once). Raises NativeError if it is absent or lacks a symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build

            _build.build()
        else:
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -m odil_b200.build` "
                "(there is no CPU fallback for the ODIL hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise NativeError(f"{LIB_PATH} does not export {name}")
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    P = ctypes.POINTER
    lib.odil_b200_version.restype = ctypes.c_int
    lib.odil_b200_last_error.restype = ctypes.c_char_p
    lib.odil_b200_launch_count.restype = i64
    lib.odil_b200_stencil_plan_create.argtypes = [ctypes.c_int, P(i64), ctypes.c_int, ctypes.c_int, P(i32), P(i32),
                                                  P(dbl), P(vp)]
    lib.odil_b200_stencil_plan_destroy.argtypes = [vp]
    lib.odil_b200_stencil_forward.argtypes = [vp, P(Slab), vp, vp, vp, vp]
    lib.odil_b200_stencil_adjoint.argtypes = [vp, P(Slab), vp, dbl, vp, vp, vp]
    lib.odil_b200_stencil_fused.argtypes = [vp, P(Slab), vp, vp, dbl, vp, vp, vp, vp]
    lib.odil_b200_stencil_plan_kind.argtypes = [vp]
    lib.odil_b200_stencil_plan_tune.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    lib.odil_b200_sum_squares.argtypes = [vp, i64, ctypes.c_int, vp, vp]
    lib.odil_b200_dot.argtypes = [vp, vp, i64, ctypes.c_int, vp, vp]
    lib.odil_b200_mg_interp_add.argtypes = [ctypes.c_int, P(i64), ctypes.c_char_p, ctypes.c_int, vp, dbl, vp, dbl, vp,
                                            P(MgRange), vp]
    lib.odil_b200_mg_interp_adjoint.argtypes = [ctypes.c_int, P(i64), ctypes.c_char_p, ctypes.c_int, vp, dbl, vp,
                                                P(MgAdjRange), vp]
    lib.odil_b200_mg_interp_adjoint_adam.argtypes = [ctypes.c_int, P(i64), ctypes.c_char_p, ctypes.c_int, vp, dbl, vp, vp,
                                                     vp, vp, dbl, vp, dbl, dbl, dbl, vp]
    lib.odil_b200_adam_synth.argtypes = [ctypes.c_int, P(i64), ctypes.c_char_p, ctypes.c_int, vp, dbl, dbl, vp, vp, vp, vp,
                                         vp, dbl, vp, dbl, dbl, dbl, P(MgRange), vp]
    lib.odil_b200_mg_restrict.argtypes = [ctypes.c_int, P(i64), ctypes.c_char_p, ctypes.c_int, vp, vp, vp]
    lib.odil_b200_adam_step.argtypes = [ctypes.c_int, P(vp), P(vp), P(vp), P(vp), P(i64), ctypes.c_int, dbl, dbl,
                                        dbl, dbl, vp]
    lib.odil_b200_adam_step_dev.argtypes = [ctypes.c_int, P(vp), P(vp), P(vp), P(vp), P(i64), ctypes.c_int, vp, dbl,
                                            dbl, dbl, vp]
    lib.odil_b200_table_pick.argtypes = [vp, vp, vp, vp]
    lib.odil_b200_gd_step.argtypes = [ctypes.c_int, P(vp), P(vp), P(i64), ctypes.c_int, dbl, vp]
    lib.odil_b200_axpby.argtypes = [i64, ctypes.c_int, dbl, vp, dbl, vp, vp]
    lib.odil_b200_multi_dot.argtypes = [vp, i64, ctypes.c_int, vp, i64, ctypes.c_int, vp, vp]
    lib.odil_b200_multi_axpy.argtypes = [vp, i64, ctypes.c_int, vp, dbl, vp, vp, i64, ctypes.c_int, vp]
    lib.odil_b200_cg_update_xr.argtypes = [i64, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.odil_b200_cg_update_p.argtypes = [i64, ctypes.c_int, vp, vp, vp, vp, vp]
    lib.odil_b200_star_worklist.argtypes = [ctypes.c_int, ctypes.c_int, i64, i64, i64, ctypes.c_int, P(i32), ctypes.c_int]
    lib.odil_b200_jit_compile.argtypes = [ctypes.c_char_p, P(ctypes.c_char_p), ctypes.c_int, P(vp)]
    lib.odil_b200_jit_log.restype = ctypes.c_char_p
    lib.odil_b200_jit_cubin.argtypes = [vp, P(vp), P(ctypes.c_uint64)]
    lib.odil_b200_jit_kernel.argtypes = [vp, ctypes.c_char_p, ctypes.c_int, P(vp)]
    lib.odil_b200_jit_launch.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p,
                                         ctypes.c_uint64, vp]
    lib.odil_b200_jit_destroy.argtypes = [vp]
    lib.odil_b200_comm_create.argtypes = [ctypes.c_int, ctypes.c_int, i64, P(vp), ctypes.c_char_p]
    lib.odil_b200_comm_connect.argtypes = [vp, ctypes.c_char_p]
    lib.odil_b200_comm_capacity.argtypes = [vp]
    lib.odil_b200_comm_capacity.restype = i64
    lib.odil_b200_halo_exchange.argtypes = [vp, ctypes.c_int, P(vp), P(vp), P(vp), P(vp), P(i64), vp]
    lib.odil_b200_halo_accumulate.argtypes = [vp, ctypes.c_int, P(vp), P(vp), P(vp), P(vp), P(i64), ctypes.c_int, vp]
    lib.odil_b200_allreduce_scalars.argtypes = [vp, vp, ctypes.c_int, vp]
    lib.odil_b200_comm_destroy.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("odil_b200_last_error", "odil_b200_launch_count", "odil_b200_version", "odil_b200_jit_log",
                        "odil_b200_comm_capacity"):
            getattr(lib, name).restype = ctypes.c_int
    _lib = lib
    return lib


def is_loaded():
    return _lib is not None


def _check(rc):
    if rc != 0:
        raise NativeError(_lib.odil_b200_last_error().decode())


def dtype_code(dtype):
    if dtype in (torch.float32, np.float32) or (not isinstance(dtype, torch.dtype) and np.dtype(dtype) == np.float32):
        return F32
    if dtype in (torch.float64, np.float64) or (not isinstance(dtype, torch.dtype) and np.dtype(dtype) == np.float64):
        return F64
    raise NativeError(f"unsupported dtype {dtype}")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, what="array"):
    if t is None:
        return None
    if not (torch.is_tensor(t) and t.is_cuda):
        raise NativeError(f"{what}: expected a CUDA tensor (the ODIL hot path has no CPU fallback)")
    if not t.is_contiguous():
        raise NativeError(f"{what}: tensor must be C-contiguous")
    return ctypes.c_void_p(t.data_ptr())


def _plane_ptr(t, halo):
    """Pointer to the first OWNED plane of a slab tensor that carries `halo` planes on both sides."""
    if t is None:
        return None
    p = _ptr(t)
    if halo:
        plane = t.stride(0) * t.element_size()
        return ctypes.c_void_p(p.value + halo * plane)
    return p


_replayed = 0
FUSED_ADAM_APPLIED = 0
ADAM_SYNTH_APPLIED = 0  # calls of odil_b200_adam_synth that ran the fused kernel
# FUSED_ADAM_APPLIED: how many times odil_b200_mg_interp_adjoint_adam really ran (it declines unsuitable arrays)


def note_replayed_launches(n):
    """Kernels executed by a CUDA-graph replay do not pass through the library's launchers; whoever replays a
    captured epoch reports how many of this library's kernel nodes the graph holds."""
    global _replayed
    _replayed += int(n)


def launch_count():
    """Kernels of this library launched in this process: direct launches + kernel nodes of replayed graphs."""
    return int(load().odil_b200_launch_count()) + _replayed


# Handles whose destructor ran while a CUDA graph was being captured: destroying them frees device memory (cudaFree /
# cuModuleUnload), which invalidates a capture in progress.  Python's cyclic collector may run a destructor at any
# allocation, so `__del__` parks the handle here and the next creation or `flush_deferred()` releases it.
_deferred = []
_capture_depth = 0  # captures announced by capture_guard(): seen by destructors that run on ANY thread


class capture_guard:
    """`with native.capture_guard(): with torch.cuda.graph(g): ...` -- while it is open, destructors of native handles on
    every thread park their handle (a collector pass triggered on another thread, e.g. a sampler thread, does not see
    the capturing stream as its current stream); leaving it releases what was parked."""

    def __enter__(self):
        global _capture_depth
        _capture_depth += 1
        return self

    def __exit__(self, *exc):
        global _capture_depth
        _capture_depth -= 1
        flush_deferred()
        return False


def _capturing():
    if _capture_depth > 0:
        return True
    try:
        import torch

        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:
        return False


def _release(destroy, handle):
    if _capturing():
        _deferred.append((destroy, handle))
    else:
        destroy(handle)


def flush_deferred():
    """Destroys the handles parked by destructors that ran during a graph capture (no-op while capturing)."""
    if _deferred and not _capturing():
        while _deferred:
            destroy, handle = _deferred.pop()
            try:
                destroy(handle)
            except Exception:
                pass


class StencilPlan:
    """Region-typed affine stencil (include/odil_b200.h: odil_b200_stencil_plan_create)."""

    def __init__(self, shape, dtype, offsets, rwidth, table):
        lib = load()
        flush_deferred()
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.dtype = dtype
        self.code = dtype_code(dtype)
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32).reshape(-1, self.ndim))
        self.offsets = offsets
        self.noff = offsets.shape[0]
        self.rwidth = tuple(int(r) for r in rwidth)
        ncls = int(np.prod([2 * r + 1 for r in self.rwidth]))
        table = np.ascontiguousarray(np.asarray(table, dtype=np.float64).reshape(ncls, self.noff))
        self.table = table
        self.rmax0 = int(np.abs(offsets[:, 0]).max()) if self.noff else 0
        shp = (ctypes.c_int64 * self.ndim)(*self.shape)
        rw = (ctypes.c_int32 * self.ndim)(*self.rwidth)
        h = ctypes.c_void_p()
        _check(lib.odil_b200_stencil_plan_create(
            self.ndim, shp, self.code, self.noff, offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), rw,
            table.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _release(_lib.odil_b200_stencil_plan_destroy, self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def kind(self):
        return int(_lib.odil_b200_stencil_plan_kind(self.handle))

    def tune(self, zchunk=0, variant=0):
        _check(_lib.odil_b200_stencil_plan_tune(self.handle, int(zchunk), int(variant)))

    def _slab(self, slab):
        if slab is None:
            return Slab(self.shape[0], 0, 0)
        return Slab(int(slab[0]), int(slab[1]), int(slab[2]))

    def forward(self, U, F_in, F_out, slab=None):
        s = self._slab(slab)
        _check(_lib.odil_b200_stencil_forward(self.handle, ctypes.byref(s), _plane_ptr(U, s.halo),
                                              _plane_ptr(F_in, s.halo), _plane_ptr(F_out, s.halo), _stream()))

    def adjoint(self, F, scale, G_in, G_out, slab=None):
        s = self._slab(slab)
        _check(_lib.odil_b200_stencil_adjoint(self.handle, ctypes.byref(s), _plane_ptr(F, s.halo), float(scale),
                                              _plane_ptr(G_in, s.halo), _plane_ptr(G_out, s.halo), _stream()))

    def fused(self, U, c, scale, G_out, sumsq_out, F_out=None, slab=None):
        s = self._slab(slab)
        _call("stencil_fused", lambda: _check(_lib.odil_b200_stencil_fused(
            self.handle, ctypes.byref(s), _plane_ptr(U, s.halo), _plane_ptr(c, s.halo), float(scale),
            _plane_ptr(G_out, s.halo), _plane_ptr(F_out, s.halo), _ptr(sumsq_out), _stream())))


def sum_squares(x, out):
    load()
    _check(_lib.odil_b200_sum_squares(_ptr(x), x.numel(), dtype_code(x.dtype), _ptr(out), _stream()))


def dot(x, y, out):
    load()
    _check(_lib.odil_b200_dot(_ptr(x), _ptr(y), x.numel(), dtype_code(x.dtype), _ptr(out), _stream()))


def _cshape(shape):
    return (ctypes.c_int64 * len(shape))(*[int(s) for s in shape])


def mg_interp_add(cshape, loc, coarse, cfac, fine_term, ffac, out, rng=None):
    """out = ffac*fine_term + cfac*I(coarse). `cshape` = GLOBAL coarse array shape."""
    load()
    r = ctypes.byref(MgRange(*[int(v) for v in rng])) if rng is not None else None
    _call("mg_interp_add", lambda: _check(_lib.odil_b200_mg_interp_add(
        len(cshape), _cshape(cshape), loc.encode(), dtype_code(out.dtype), _ptr(coarse), float(cfac),
        _ptr(fine_term), float(ffac), _ptr(out), r, _stream())))


def mg_interp_adjoint(cshape, loc, g_fine, scale, g_coarse, rng=None):
    load()
    r = ctypes.byref(MgAdjRange(*[int(v) for v in rng])) if rng is not None else None
    _call("mg_interp_adjoint", lambda: _check(_lib.odil_b200_mg_interp_adjoint(
        len(cshape), _cshape(cshape), loc.encode(), dtype_code(g_fine.dtype), _ptr(g_fine), float(scale),
        _ptr(g_coarse), r, _stream())))


def mg_interp_adjoint_adam(cshape, loc, g_fine, scale, g_coarse, x, m, v, alpha, omb1, omb2, eps, alpha_dev=None):
    """g_coarse = scale * I^T g_fine and the Adam update of (x, m, v) with gradient g_fine in one pass over g_fine.
    Returns False (nothing done) if the arrays do not fit the fused kernel."""
    load()
    for t in (x, m, v):
        if t.dtype != g_fine.dtype or t.shape != g_fine.shape:
            raise NativeError("fused Adam: x, m, v must match the fine gradient in dtype and shape")
    res = {}

    def run():
        rc = _lib.odil_b200_mg_interp_adjoint_adam(
            len(cshape), _cshape(cshape), loc.encode(), dtype_code(g_fine.dtype), _ptr(g_fine), float(scale),
            _ptr(g_coarse), _ptr(x), _ptr(m), _ptr(v), float(alpha), _ptr(alpha_dev) if alpha_dev is not None else None,
            float(omb1), float(omb2), float(eps), _stream())
        if rc not in (0, 1):
            _check(rc)
        res["applied"] = rc == 0

    _call("mg_interp_adjoint_adam", run)
    global FUSED_ADAM_APPLIED
    FUSED_ADAM_APPLIED += int(res["applied"])
    return res["applied"]


def adam_synth(cshape, loc, coarse, cfac, ffac, x, m, v, g, out, alpha, omb1, omb2, eps, alpha_dev=None, rng=None):
    """Adam update of the finest multigrid term (x, m, v with gradient g) and out = ffac * x_new + cfac * I(coarse) in one
    pass (odil_b200_adam_synth).  rng = (fz_begin, fz_end, out_z0, coarse_z0) as in mg_interp_add (slabs).  Returns
    False (nothing done) if the arrays do not fit the fused kernel."""
    load()
    r = ctypes.byref(MgRange(*[int(q) for q in rng])) if rng is not None else None
    for t in (m, v, g, out):
        if t.dtype != x.dtype or t.shape != x.shape:
            raise NativeError("adam_synth: m, v, g, out must match x in dtype and shape")
    res = {}

    def run():
        rc = _lib.odil_b200_adam_synth(
            len(cshape), _cshape(cshape), loc.encode(), dtype_code(x.dtype), _ptr(coarse), float(cfac), float(ffac),
            _ptr(x), _ptr(m), _ptr(v), _ptr(g), _ptr(out), float(alpha),
            _ptr(alpha_dev) if alpha_dev is not None else None, float(omb1), float(omb2), float(eps), r, _stream())
        if rc not in (0, 1):
            _check(rc)
        res["applied"] = rc == 0

    _call("adam_synth", run)
    global ADAM_SYNTH_APPLIED
    ADAM_SYNTH_APPLIED += int(res["applied"])
    return res["applied"]


def mg_restrict(fshape, loc, fine, out):
    load()
    _check(_lib.odil_b200_mg_restrict(len(fshape), _cshape(fshape), loc.encode(), dtype_code(fine.dtype), _ptr(fine),
                                      _ptr(out), _stream()))


def _ptr_array(tensors, dtype=None, counts=None):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        if dtype is not None and t.dtype != dtype:
            raise NativeError(f"tensor {i}: dtype {t.dtype} does not match {dtype}")
        if counts is not None and t.numel() != counts[i]:
            raise NativeError(f"tensor {i}: {t.numel()} elements, expected {counts[i]}")
        arr[i] = _ptr(t).value
    return arr


def adam_step(x, m, v, g, alpha, omb1, omb2, eps):
    load()
    n = len(x)
    if n == 0:
        return
    cnt = [t.numel() for t in x]
    counts = (ctypes.c_int64 * n)(*cnt)
    dt = x[0].dtype
    _call("adam_step", lambda: _check(_lib.odil_b200_adam_step(
        n, _ptr_array(x, dt), _ptr_array(m, dt, cnt), _ptr_array(v, dt, cnt), _ptr_array(g, dt, cnt), counts,
        dtype_code(x[0].dtype), float(alpha), float(omb1), float(omb2), float(eps), _stream())))


def table_pick(table, step, out):
    """out[0] = table[step[0]]; step[0] += 1 on the device (float64 table, 1-element int64 `step`, 1-element float64 `out`)."""
    load()
    if table.dtype != torch.float64 or out.dtype != torch.float64 or step.dtype != torch.int64 or not table.is_cuda:
        raise NativeError("table_pick: float64 CUDA table / out and an int64 CUDA step counter")
    _call("table_pick", lambda: _check(_lib.odil_b200_table_pick(_ptr(table), _ptr(step), _ptr(out), _stream())))


def adam_step_dev(x, m, v, g, alpha_dev, omb1, omb2, eps):
    """adam_step with the step size in a 1-element float64 CUDA tensor (constant launch arguments: graph-capturable)."""
    load()
    if alpha_dev.dtype != torch.float64 or not alpha_dev.is_cuda or alpha_dev.numel() != 1:
        raise NativeError("adam_step_dev: alpha_dev must be a 1-element float64 CUDA tensor")
    n = len(x)
    if n == 0:
        return
    cnt = [t.numel() for t in x]
    counts = (ctypes.c_int64 * n)(*cnt)
    dt = x[0].dtype
    _call("adam_step", lambda: _check(_lib.odil_b200_adam_step_dev(
        n, _ptr_array(x, dt), _ptr_array(m, dt, cnt), _ptr_array(v, dt, cnt), _ptr_array(g, dt, cnt), counts,
        dtype_code(x[0].dtype), ctypes.c_void_p(alpha_dev.data_ptr()), float(omb1), float(omb2), float(eps),
        _stream())))


def gd_step(x, g, lr):
    load()
    n = len(x)
    cnt = [t.numel() for t in x]
    counts = (ctypes.c_int64 * n)(*cnt)
    _check(_lib.odil_b200_gd_step(n, _ptr_array(x, x[0].dtype), _ptr_array(g, x[0].dtype, cnt), counts,
                                  dtype_code(x[0].dtype), float(lr),
                                  _stream()))


def axpby(a, x, b, y):
    load()
    _check(_lib.odil_b200_axpby(x.numel(), dtype_code(x.dtype), float(a), _ptr(x), float(b), _ptr(y), _stream()))


def star_worklist(dtype, n0, N1, N2, variant=-1, zchunk=0):
    """CTAs of the fused star sweep for a slab of n0 planes: int32 array [ncta, 5] = x0, y0, rows, z begin, z end
    (host-side inspection, needs no GPU)."""
    load()
    import numpy as np

    code = dtype_code(dtype)
    n = _lib.odil_b200_star_worklist(code, int(variant), int(n0), int(N1), int(N2), int(zchunk), None, 0)
    if n < 0:
        raise NativeError(_lib.odil_b200_last_error().decode())
    out = np.zeros((n, 5), dtype=np.int32)
    _check(min(0, _lib.odil_b200_star_worklist(code, int(variant), int(n0), int(N1), int(N2), int(zchunk),
                                               out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n)))
    return out


def cg_update_xr(num, den, p, q, x, r):
    """alpha = num/den (device doubles); x += alpha p; r -= alpha q."""
    load()
    _check(_lib.odil_b200_cg_update_xr(x.numel(), dtype_code(x.dtype), _ptr(num), _ptr(den), _ptr(p), _ptr(q), _ptr(x),
                                       _ptr(r), _stream()))


def cg_update_p(num, den, r, p):
    """beta = num/den (device doubles); p = r + beta p."""
    load()
    _check(_lib.odil_b200_cg_update_p(p.numel(), dtype_code(p.dtype), _ptr(num), _ptr(den), _ptr(r), _ptr(p), _stream()))


def multi_dot(V, k, g, out):
    """out[r] = <V[r], g> for the first k rows of the row-major matrix V (device doubles)."""
    load()
    _check(_lib.odil_b200_multi_dot(_ptr(V), V.stride(0), int(k), _ptr(g), g.numel(), dtype_code(g.dtype), _ptr(out),
                                    _stream()))


def multi_axpy(V, k, coef, a0, g, d):
    """d = a0*g + sum_r coef[r]*V[r]  (coef: device float64 tensor with k entries)."""
    load()
    _check(_lib.odil_b200_multi_axpy(_ptr(V), V.stride(0), int(k), _ptr(coef), float(a0), _ptr(g), _ptr(d), d.numel(),
                                     dtype_code(d.dtype), _stream()))


# --------------------------------------------------------------------------------------------------
# Run-time specialised kernels (csrc/jit.cu): NVRTC -> sm_100a cubin -> driver launch
# --------------------------------------------------------------------------------------------------
class JitModule:
    """CUDA C source compiled for sm_100a (compilation needs no device; `kernel` / `launch` do)."""

    def __init__(self, source, options=()):
        lib = load()
        self.source = source
        self.handle = ctypes.c_void_p()
        opts = (ctypes.c_char_p * max(1, len(options)))(*[o.encode() for o in options])
        rc = lib.odil_b200_jit_compile(source.encode(), opts, len(options), ctypes.byref(self.handle))
        self.log = (lib.odil_b200_jit_log() or b"").decode(errors="replace").strip("\x00").strip()
        if rc != 0:
            raise NativeError(lib.odil_b200_last_error().decode() + "\n" + self.log[-4000:])
        self._kernels = {}

    def cubin(self):
        data, size = ctypes.c_void_p(), ctypes.c_uint64()
        _check(_lib.odil_b200_jit_cubin(self.handle, ctypes.byref(data), ctypes.byref(size)))
        return ctypes.string_at(data.value, size.value)

    def kernel(self, name, max_dynamic_smem=0):
        k = self._kernels.get(name)
        if k is None:
            k = ctypes.c_void_p()
            _check(_lib.odil_b200_jit_kernel(self.handle, name.encode(), int(max_dynamic_smem), ctypes.byref(k)))
            self._kernels[name] = k
        return k

    def launch(self, name, grid, block, params, smem=0):
        k = self.kernel(name)
        _call("jit:" + name, lambda: _check(_lib.odil_b200_jit_launch(k, int(grid), int(block), int(smem), params,
                                                                     len(params), _stream())))

    def __del__(self):
        try:
            if _lib is not None and self.handle:
                _release(_lib.odil_b200_jit_destroy, self.handle)
                self.handle = None
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------
# Slab communicator (csrc/comm.cu): halo exchange and scalar all-reduce over NVLink peer memory
# --------------------------------------------------------------------------------------------------
IPC_HANDLE_BYTES = 64


class Comm:
    """One per process (= per GPU).  `all_gather_bytes(bytes) -> [bytes per rank]` is the only thing asked of the host
    plumbing (torch.distributed): it carries the 64-byte IPC handles once, at creation."""

    def __init__(self, rank, world, halo_bytes, all_gather_bytes):
        lib = load()
        self.rank, self.world = int(rank), int(world)
        self.handle = ctypes.c_void_p()
        mine = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        _check(lib.odil_b200_comm_create(self.rank, self.world, int(halo_bytes), ctypes.byref(self.handle), mine))
        handles = all_gather_bytes(mine.raw)
        assert len(handles) == self.world and all(len(h) == IPC_HANDLE_BYTES for h in handles)
        _check(lib.odil_b200_comm_connect(self.handle, b"".join(handles)))
        self.capacity = int(lib.odil_b200_comm_capacity(self.handle))

    @staticmethod
    def bytes_needed(nbytes_list):
        return sum((n + 255) // 256 * 256 for n in nbytes_list)

    def halo_exchange(self, send_lo, send_hi, recv_lo, recv_hi):
        """Lists of contiguous CUDA tensor views (same length, pairwise equal sizes)."""
        n = len(send_lo)
        VP = ctypes.c_void_p * n
        nbytes = (ctypes.c_int64 * n)(*[t.numel() * t.element_size() for t in send_lo])
        args = [VP(*[_ptr(t).value for t in ts]) for ts in (send_lo, send_hi, recv_lo, recv_hi)]
        _call("halo_exchange", lambda: _check(_lib.odil_b200_halo_exchange(self.handle, n, *args, nbytes, _stream())))

    def halo_accumulate(self, send_lo, send_hi, acc_lo, acc_hi):
        """Adds the neighbours' partial sums (their send_hi / send_lo planes) to this rank's first / last owned planes."""
        n = len(send_lo)
        VP = ctypes.c_void_p * n
        nbytes = (ctypes.c_int64 * n)(*[t.numel() * t.element_size() for t in send_lo])
        args = [VP(*[_ptr(t).value for t in ts]) for ts in (send_lo, send_hi, acc_lo, acc_hi)]
        code = dtype_code(send_lo[0].dtype)
        if any(t.dtype != send_lo[0].dtype for ts in (send_lo, send_hi, acc_lo, acc_hi) for t in ts):
            raise NativeError("halo_accumulate: all arrays of one call must have the same dtype")
        _call("halo_accumulate", lambda: _check(_lib.odil_b200_halo_accumulate(self.handle, n, *args, nbytes, code,
                                                                                _stream())))

    def allreduce_scalars(self, t):
        """In-place sum over ranks of a small float64 CUDA tensor (deterministic: rank order)."""
        if t.dtype != torch.float64:
            raise NativeError("allreduce_scalars: float64 tensor expected")
        _call("allreduce_scalars", lambda: _check(_lib.odil_b200_allreduce_scalars(self.handle, _ptr(t), t.numel(),
                                                                                    _stream())))

    def destroy(self):
        if self.handle:
            _lib.odil_b200_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()
