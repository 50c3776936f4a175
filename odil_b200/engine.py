"""
ResidualEngine: traces the user operator once, lowers every output to region-typed stencil plans and
evaluates loss + gradient with the CUDA kernels.  This is what `Problem._eval_loss_grad_b200` calls;
it replaces the jitted XLA program of the reference (core.py:1076-1111):

    reference                                      here
    ---------------------------------------------  ---------------------------------------------
    multigrid_to_regular / interp_to_finer         odil_b200_mg_interp_add      (one launch / level)
    operator(ctx): roll, where, arithmetic         odil_b200_stencil_fused      (one sweep: F, sum F^2,
    mean(square(F)), jax.value_and_grad                                           g_U = (2/n) A^T F)
    AD through interp_to_finer                     odil_b200_mg_interp_adjoint  (one launch / level)

Outputs that are affine in the unknown fields with coefficients that are piecewise constant per
boundary region (Poisson, wave, identity-type data terms) are supported; anything else raises
NonAffineError at trace time (no silent fallback).
"""
import math
import os

import numpy as np
import torch

from . import native
from .backend import Affine, Known, NonAffineError, as_known, torch_dtype
from .core import Array, Context, Field, MultigridField, NeuralNet

MAX_REGION_WIDTH = 8


class _Fetch:
    """One device->host copy shared by all scalars of an evaluation, done on first use."""

    def __init__(self, sums, counts, dtype, raws=None):
        self.sums = sums
        self.counts = counts
        self.dtype = dtype
        self.raws = raws or [False] * len(counts)  # Context.Raw outputs: term = mean(value), norm = the term itself
        self._terms = None

    def terms(self):
        if self._terms is None:
            host = self.sums.detach().cpu().numpy()
            self._terms = [self.dtype.type(s / n) for s, n in zip(host, self.counts)]
        return self._terms

    def rearm(self):
        """Forget the host copy: the device sums were rewritten in place (CUDA-graph replay of the epoch)."""
        self._terms = None


class LazyScalar:
    """Device-resident scalar (loss / term / norm). Converts on np.array(), float(), format()."""
    __array_priority__ = 1000

    def __init__(self, fetch, kind, index=None):
        self._fetch, self._kind, self._index = fetch, kind, index

    def value(self):
        t = self._fetch.terms()
        if self._kind == "loss":
            return self._fetch.dtype.type(sum(t))
        if self._kind == "term" or self._fetch.raws[self._index]:
            return t[self._index]
        return np.sqrt(t[self._index])

    def __array__(self, dtype=None, copy=None):
        a = np.array(self.value())
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.value())

    def item(self):
        return self.value().item()

    def __format__(self, spec):
        return format(self.value(), spec)

    def __repr__(self):
        return repr(self.value())

    def _bin(self, other, op):
        return op(self.value(), other.value() if isinstance(other, LazyScalar) else other)

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._bin(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __lt__(self, o):
        return self._bin(o, lambda a, b: a < b)

    def __gt__(self, o):
        return self._bin(o, lambda a, b: a > b)


# --------------------------------------------------------------------------------------------------
# Lowering: Affine expression -> region-typed table
# --------------------------------------------------------------------------------------------------
def axis_region_width(t, axis, n):
    """Smallest r such that compact tensor `t` is constant along `axis` on [r, n - r)."""
    if t.shape[axis] == 1 or n <= 1:
        return 0
    diff = t.narrow(axis, 1, n - 1) != t.narrow(axis, 0, n - 1)
    other = [a for a in range(t.dim()) if a != axis]
    if other:
        diff = diff.sum(dim=other) > 0
    idx = torch.nonzero(diff.reshape(-1)).reshape(-1)
    if idx.numel() == 0:
        return 0
    idx = idx.to(torch.int64)
    return int(torch.minimum(idx + 1, (n - 1) - idx).max().item())


def representative_indices(n, r):
    mid = min(r, n - 1)
    return list(range(r)) + [mid] + list(range(n - r, n))


def lower_block(shape, coefs):
    """
    coefs: {offset tuple: Coef}.  Returns (offsets [noff, ndim], rwidth [ndim], table [ncls, noff]).
    Raises NonAffineError if some coefficient varies in the interior (not region-typed).
    """
    nd = len(shape)
    # roll() along an axis of size n is periodic with period n: reduce every offset to the shortest equivalent one
    # (on a size-1 axis every shift is the identity) and merge offsets that became equal.
    folded = {}
    for off, co in coefs.items():
        red = tuple(0 if shape[a] == 1 else int(o) - shape[a] * round(int(o) / shape[a]) if abs(int(o)) >= shape[a]
                    else int(o) for a, o in enumerate(off))
        folded[red] = folded[red].added(co) if red in folded else co.copy()
    coefs = folded
    offsets = sorted(coefs.keys())
    rr = [0] * nd
    for off in offsets:
        for t in coefs[off].terms:
            for a in range(nd):
                rr[a] = max(rr[a], axis_region_width(t, a, shape[a]))
    for a in range(nd):
        if rr[a] > MAX_REGION_WIDTH or 2 * rr[a] > shape[a]:
            raise NonAffineError(
                f"stencil coefficients vary inside the domain along axis {a} (region width {rr[a]}): per-cell "
                "coefficient arrays are not supported on the fused path yet")
    reps = [representative_indices(shape[a], rr[a]) for a in range(nd)]
    cls_shape = tuple(2 * r + 1 for r in rr)
    table = np.zeros(cls_shape + (len(offsets),), dtype=np.float64)
    for o, off in enumerate(offsets):
        acc = torch.zeros(cls_shape, dtype=torch.float64)
        for t in coefs[off].terms:
            s = t
            for a in range(nd):
                if s.shape[a] != 1:
                    s = s.index_select(a, torch.as_tensor(reps[a], device=s.device))
            acc = acc + s.to(torch.float64).cpu()
        table[..., o] = acc.cpu().numpy()
    keep = [o for o in range(len(offsets)) if np.any(table[..., o] != 0)]
    if not keep:
        keep = [0]
    offsets = [offsets[o] for o in keep]
    table = table[..., keep]
    return np.asarray(offsets, dtype=np.int32).reshape(len(offsets), nd), rr, table.reshape(-1, len(offsets))


def wraps_axis0(shape, offsets, rwidth, table):
    """True if some row of the region table multiplies a neighbour that lies across the axis-0
    boundary (i.e. the stencil really is periodic along axis 0, not overridden by boundary rows)."""
    n0, r0 = shape[0], rwidth[0]
    ncls0 = 2 * r0 + 1
    t = np.asarray(table).reshape((ncls0, -1, len(offsets)))
    rows = list(range(r0)) + [min(r0, n0 - 1)] + list(range(n0 - r0, n0))   # representative index per class
    interior_lo, interior_hi = r0, n0 - r0 - 1                               # index range of the interior class
    for o, off in enumerate(offsets):
        d = int(off[0])
        if d == 0:
            continue
        for cl, i in enumerate(rows):
            if cl == r0:
                out = (interior_lo + d < 0) or (interior_hi + d > n0 - 1) if interior_hi >= interior_lo else False
            else:
                out = (i + d < 0) or (i + d > n0 - 1)
            if out and np.any(t[cl, :, o] != 0):
                return True
    return False


def synthesize_slab(slab, arrays, shapes, factors, loc, buffers=None, exchanged=False, stop_level=0):
    """
    Multigrid synthesis on local slabs: every level is produced on its EXTENDED range (owned planes
    plus the halo, clipped to the domain) from the level below, after one batched halo exchange of
    all terms.  `shapes` are the GLOBAL array shapes per level.  Returns the local U (with halos).
    stop_level > 0: stop at that level and return (array, factor) of it (adam_synth_step: level 0 is produced by
    odil_b200_adam_synth).
    """
    L = len(arrays)
    if not exchanged:
        slab.exchange(list(arrays), width=slab.halo)
    if L == 1:
        return arrays[0] if factors[0] == 1 else arrays[0] * factors[0]
    H = slab.halo
    res, cfac = arrays[L - 1], float(factors[L - 1])
    if stop_level > 0:
        for lvl in range(L - 2, stop_level - 1, -1):
            res = _synth_slab_level(slab, arrays, shapes, factors, loc, buffers, lvl, res, cfac)
            cfac = 1.0
        return res, cfac
    for lvl in range(L - 2, -1, -1):
        fshape, cshape = shapes[lvl], shapes[lvl + 1]
        zf, nf = slab.owned_range(fshape)
        zc, _ = slab.owned_range(cshape)
        key = ("V", lvl)
        out = None if buffers is None else buffers.get(key)
        if out is None or tuple(out.shape) != tuple(arrays[lvl].shape):
            out = torch.zeros_like(arrays[lvl])
            if buffers is not None:
                buffers[key] = out
        lo, hi = max(zf - H, 0), min(zf + nf + H, fshape[0])
        if loc[0] == "c":
            lo -= lo % 2
            hi += hi % 2
        native.mg_interp_add(cshape, loc, res, cfac, arrays[lvl], float(factors[lvl]), out,
                             rng=(lo, hi, zf - H, zc - H))
        res, cfac = out, 1.0
    return res


def _synth_slab_level(slab, arrays, shapes, factors, loc, buffers, lvl, res, cfac):
    """One level of synthesize_slab: V_lvl = factors[lvl] * arrays[lvl] + cfac * I(res) on the extended range."""
    H = slab.halo
    fshape, cshape = shapes[lvl], shapes[lvl + 1]
    zf, nf = slab.owned_range(fshape)
    zc, _ = slab.owned_range(cshape)
    key = ("V", lvl)
    out = None if buffers is None else buffers.get(key)
    if out is None or tuple(out.shape) != tuple(arrays[lvl].shape):
        out = torch.zeros_like(arrays[lvl])
        if buffers is not None:
            buffers[key] = out
    lo, hi = max(zf - H, 0), min(zf + nf + H, fshape[0])
    if loc[0] == "c":
        lo -= lo % 2
        hi += hi % 2
    native.mg_interp_add(cshape, loc, res, cfac, arrays[lvl], float(factors[lvl]), out, rng=(lo, hi, zf - H, zc - H))
    return out


class TracerView(dict):
    """What the operator sees as `ctx.tracers` during the trace.  The reference passes tracers to the jitted
    function as run-time arguments (core.py:1076-1110); the affine path bakes their values into the coefficient
    tables, so every key that was READ is recorded with its value and `Problem` re-lowers the operator when one of
    them changes (EpochCallback updates tracers["epoch"] every epoch, util.py:370-378)."""

    def __init__(self, values):
        super().__init__(values or {})
        self.reads = {}

    def __getitem__(self, key):
        v = super().__getitem__(key)
        self.reads[key] = v
        return v

    def get(self, key, default=None):
        if key in self:
            return self[key]
        return default

    def stale(self, current):
        """True if a tracer the trace depended on now has another value."""
        for key, v in self.reads.items():
            c = (current or {}).get(key)
            try:
                same = bool(np.all(np.asarray(c) == np.asarray(v)))
            except Exception:
                same = c is v
            if not same:
                return True
        return False


class _Block:
    def __init__(self, key, frozen, plan):
        self.key, self.frozen, self.plan = key, frozen, plan


class _Output:
    def __init__(self, name, shape, const, blocks):
        self.name, self.shape, self.const, self.blocks = name, shape, const, blocks
        self.n = math.prod(shape)
        self.fused = False


class _Unknown:
    """How one state entry maps to arrays of the flat unknown list."""

    def __init__(self, key, field, first, domain):
        self.key, self.first = key, first
        self.kind = type(field).__name__
        if isinstance(field, MultigridField):
            self.narrays = len(field.terms)
            self.loc = field.loc
            self.shapes = [domain._get_field_shape(t.cshape, field.loc) for t in field.terms]  # GLOBAL shapes
        elif isinstance(field, Field):
            self.narrays = 1
            self.shapes = [domain._get_field_shape(field.cshape or domain.cshape, field.loc)]
        elif isinstance(field, Array):
            self.narrays = 1
            self.shapes = [tuple(field.array.shape)]
        elif isinstance(field, NeuralNet):
            self.narrays = len(field.weights) + len(field.biases)
            self.shapes = [tuple(a.shape) for a in list(field.weights) + list(field.biases)]
        else:
            raise TypeError("Unknown field type '{}'".format(type(field).__name__))
        self.shape = self.shapes[0]


class ResidualEngine:

    def __init__(self, problem, state, trace_only=False):
        """trace_only=True stops after lowering (no CUDA needed): used to inspect / test the plans."""
        self.trace_only = trace_only
        if not trace_only:
            native.load()
            if not torch.cuda.is_available():
                raise native.NativeError("the ODIL B200 engine needs a CUDA device (no CPU fallback)")
        self.problem = problem
        self.domain = domain = problem.domain
        self.dtype = np.dtype(domain.dtype)
        self.tdtype = torch_dtype(self.dtype)
        self.device = domain.mod.device
        # unknown layout
        self.unknowns = {}
        first = 0
        for key, field in state.fields.items():
            u = _Unknown(key, field, first, domain)
            if isinstance(field, MultigridField):
                u.mgloc = domain._mg_loc(field)
                u.factors = [float(f) for f in (field.factors or domain.mg_factors or [1] * len(field.terms))]
            self.unknowns[key] = u
            first += u.narrays
        self.narrays = first
        self.slab = getattr(domain, "slab", None)
        self._trace(state)
        self._buffers = {}
        if self.slab is not None and not trace_only:
            self._prepare_slabs()

    # ----------------------------------------------------------------------------------------------
    def _trace(self, state):
        problem, domain = self.problem, self.domain
        self.tracer_view = TracerView(problem.tracers if isinstance(problem.tracers, dict) else None)
        ctx = Context(domain, state, extra=problem.extra,
                      tracers=self.tracer_view if isinstance(problem.tracers, dict) else problem.tracers)
        ff = problem.operator(ctx)
        assert isinstance(ff, (tuple, list)) and len(ff), "Operator must return a non-empty list"
        names = [f[0] if isinstance(f, tuple) else "" for f in ff]
        nonempty = [n for n in names if n]
        assert len(nonempty) == len(set(nonempty)), "Name of fields must be unique, got {}".format(nonempty)
        values = [f[1] if isinstance(f, tuple) else f for f in ff]
        self.names = names
        self.outputs = []
        use_count = {}
        for name, v in zip(names, values):
            if isinstance(v, Context.Raw):
                raise NonAffineError("Context.Raw loss terms are not on the fused path yet")
            if not isinstance(v, Affine):
                v = as_known(v)
                const = v.full().to(device=self.device, dtype=self.tdtype).contiguous()
                self.outputs.append(_Output(name, tuple(v.shape), const, []))
                continue
            groups = {}
            for (key, off, frozen), coef in v.lin.items():
                groups.setdefault((key, frozen), {})[off] = coef
            blocks = []
            for (key, frozen), coefs in groups.items():
                unk = self.unknowns[key]
                if tuple(unk.shape) != tuple(v.shape):
                    raise NonAffineError(f"output shape {v.shape} differs from field '{key}' shape {unk.shape}")
                offsets, rr, table = lower_block(v.shape, coefs)
                blk = _Block(key, frozen, None)
                blk.spec = dict(shape=tuple(v.shape), offsets=offsets, rwidth=tuple(rr), table=table)
                if not self.trace_only:
                    blk.plan = native.StencilPlan(v.shape, self.tdtype, offsets, rr, table)
                blocks.append(blk)
                if not frozen:
                    use_count[key] = use_count.get(key, 0) + 1
            const = None
            if not v.const.is_zero():
                const = v.const.dense(v.shape, self.tdtype, self.device).contiguous()
            self.outputs.append(_Output(name, tuple(v.shape), const, blocks))
        for out in self.outputs:
            out.fused = (len(out.blocks) == 1 and not out.blocks[0].frozen and use_count.get(out.blocks[0].key) == 1)
        self.used_keys = set(use_count)

    def _prepare_slabs(self):
        """Slices the constant terms into local slabs and decides which fields need a U halo exchange."""
        slab = self.slab
        self.periodic_keys = set()
        for out in self.outputs:
            if not out.fused:
                raise NonAffineError("multi-GPU slabs support single-field fused outputs only (so far)")
            blk = out.blocks[0]
            if self.unknowns[blk.key].kind == "Array":
                raise NonAffineError("non-grid Array unknowns are not decomposed into slabs yet")
            out.const_local = slab.scatter(out.const) if out.const is not None else None
            sp = blk.spec
            if wraps_axis0(sp["shape"], sp["offsets"], sp["rwidth"], sp["table"]):
                self.periodic_keys.add(blk.key)
            out.const = None  # the global copy is no longer needed

    # ----------------------------------------------------------------------------------------------
    def _loss_grad_slab(self, arrays):
        slab, H = self.slab, self.slab.halo
        K = len(self.outputs)
        sums = torch.empty(K, dtype=torch.float64, device=self.device)
        used = [self.unknowns[k] for k in self.unknowns if k in self.used_keys]
        # regular fields left behind by adam_synth_step (computed from exactly these arrays, halos already exchanged)
        hits = {}
        for u in used:
            cached = self._synth_cache.pop(u.key, None) if hasattr(self, "_synth_cache") else None
            if cached is not None and cached[0] == self._signature(arrays[u.first: u.first + u.narrays]):
                hits[u.key] = cached[1]
        # one batched halo exchange of every array of every other used unknown
        slab.exchange([arrays[u.first + i] for u in used if u.key not in hits for i in range(u.narrays)], width=H)
        U = {}
        for u in used:
            a = arrays[u.first: u.first + u.narrays]
            if u.key in hits:
                U[u.key] = hits[u.key]
            elif u.kind == "MultigridField":
                bufs = self._buffers.setdefault(("slabV", u.key), {})
                U[u.key] = synthesize_slab(slab, a, u.shapes, u.factors, u.mgloc, bufs, exchanged=True)
                if u.key in self.periodic_keys:
                    slab.exchange([U[u.key]], width=H)
            else:
                U[u.key] = a[0]
        # Gradient work arrays carry GH = 3 zero halo planes per side (the state arrays carry H = 2).  The transposed
        # interpolation is done push-style: every rank computes the coarse planes of its slab AND the one plane beyond
        # each end from its own fine planes only (planes it does not own read as zero -- GH = 3 keeps every such read
        # inside the allocation), so those end planes and the first / last owned planes hold PARTIAL sums.  By
        # linearity the partial sums can be carried through all levels; ONE halo accumulation at the end hands them to
        # their owners (round 1 exchanged the gradient halo once per level: three more synchronisation points).
        GH = max(3, H)
        gshape = lambda shape: (slab.check(shape) + 2 * GH,) + tuple(shape[1:])
        view = lambda g: g[GH - H: g.shape[0] - (GH - H)]  # the H-halo layout of the state arrays
        grads = [None] * self.narrays
        acc_items = []
        for k, out in enumerate(self.outputs):
            blk = out.blocks[0]
            u = self.unknowns[blk.key]
            z0, n0 = slab.owned_range(out.shape)
            g = self._buf(("gU", blk.key), gshape(out.shape), zero=True)
            blk.plan.fused(U[blk.key], out.const_local, 2.0 / out.n, view(g), sums[k:k + 1], slab=(n0, z0, H))
            if u.kind != "MultigridField":
                grads[u.first] = view(g)
                continue
            gl = g
            for lvl in range(u.narrays):
                if lvl > 0:
                    zc, nc = slab.owned_range(u.shapes[lvl])
                    zf, _ = slab.owned_range(u.shapes[lvl - 1])
                    cb, ce = max(zc - 1, 0), min(zc + nc + 1, u.shapes[lvl][0])
                    gc = self._buf(("gV", u.key, lvl), gshape(u.shapes[lvl]), zero=True)
                    native.mg_interp_adjoint(u.shapes[lvl], u.mgloc, gl, 1.0, gc, rng=(cb, ce, zc - GH, zf - GH))
                    acc_items.append((gc[GH - 1:GH], gc[GH + nc:GH + nc + 1], gc[GH:GH + 1], gc[GH + nc - 1:GH + nc]))
                    gl = gc
                f = u.factors[lvl]
                grads[u.first + lvl] = (view(gl), f)
        slab.accumulate(acc_items)
        for i in range(self.narrays):
            if isinstance(grads[i], tuple):
                gv, f = grads[i]
                grads[i] = gv if f == 1 else gv * f
        for i in range(self.narrays):
            if grads[i] is None:
                grads[i] = torch.zeros_like(arrays[i])
        slab.all_reduce_sum(sums)
        fetch = _Fetch(sums, [o.n for o in self.outputs], self.dtype)
        return (LazyScalar(fetch, "loss"), grads, [LazyScalar(fetch, "term", k) for k in range(K)],
                [LazyScalar(fetch, "norm", k) for k in range(K)])

    # ----------------------------------------------------------------------------------------------
    def _buf(self, name, shape, zero=False):
        b = self._buffers.get(name)
        if b is None or tuple(b.shape) != tuple(shape):
            alloc = torch.zeros if zero else torch.empty
            b = alloc(tuple(shape), dtype=self.tdtype, device=self.device)
            self._buffers[name] = b
        return b

    def _check_arrays(self, arrays):
        if self.trace_only:
            raise native.NativeError("engine was built with trace_only=True")
        if len(arrays) != self.narrays:
            raise ValueError(f"expected {self.narrays} arrays, got {len(arrays)}")
        for a in arrays:
            if not (torch.is_tensor(a) and a.is_cuda and a.dtype == self.tdtype and a.is_contiguous()):
                raise native.NativeError(
                    "state arrays must be contiguous CUDA tensors of the domain dtype; use domain.init_state() "
                    "(the ODIL hot path has no CPU fallback)")

    @staticmethod
    def _signature(tensors):
        """Identity + torch version counter of a list of tensors: changes when an array is replaced or written to by a
        torch operation (the library's own kernels write behind torch's back; see adam_synth_step)."""
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def _regular(self, unk, arrays):
        """Regular field U of one unknown (multigrid synthesis when needed)."""
        a = arrays[unk.first: unk.first + unk.narrays]
        if unk.kind != "MultigridField":
            return a[0]
        cached = self._synth_cache.pop(unk.key, None) if hasattr(self, "_synth_cache") else None
        if cached is not None and cached[0] == self._signature(a):
            return cached[1]  # written by adam_synth_step together with the update that produced these arrays
        L = unk.narrays
        if L == 1:
            return a[0] if unk.factors[0] == 1 else a[0] * unk.factors[0]
        res, cfac = a[L - 1], unk.factors[L - 1]
        for lvl in range(L - 2, -1, -1):
            out = self._buf(("V", unk.key, lvl), unk.shapes[lvl])
            native.mg_interp_add(unk.shapes[lvl + 1], unk.mgloc, res, cfac, a[lvl], unk.factors[lvl], out)
            res, cfac = out, 1.0
        return res

    def adam_synth_step(self, x, m, v, grads, alpha, omb1, omb2, eps, alpha_dev=None):
        """The optimizer's Adam step (optimizer.py:311-319) on all arrays, with the update of the FINEST term of every
        used multigrid unknown fused with the synthesis of the regular field the next loss_grad() needs
        (odil_b200_adam_synth): coarser terms first, then the coarse levels are synthesised, then one pass over the
        finest term writes t0, m, v and U = t0 + I(V1).  The next loss_grad() on the same, untouched arrays picks U up
        instead of synthesising level 0 again.  Arrays the fused kernel does not fit take the plain update; results
        are bit-identical to native.adam_step followed by the usual synthesis."""
        if not hasattr(self, "_synth_cache"):
            self._synth_cache = {}
        keys = self.used_keys if self.slab is not None else (self.used_keys | self._frozen_keys())
        # candidates: 3-D multigrid unknowns (the fused kernel marches 3-D cell-centred grids; asking for it on a 2-D
        # grid cost configs[1] a wasted synthesis of the coarse levels and a second Adam launch per epoch: 51 vs 45.5 us)
        cand = [u for k, u in self.unknowns.items() if k in keys and u.kind == "MultigridField" and u.narrays >= 2
                and len(u.shapes[0]) == 3]
        first = {u.first for u in cand}
        # ODIL_B200_SYNTH_CHAIN=1: the intermediate levels take the same fused kernel (Adam of t_l + V_l = t_l + I(V_l+1))
        # instead of the multi-tensor Adam followed by mg_interp_add -- one pass over t_l instead of two
        chain = self.slab is None and os.environ.get("ODIL_B200_SYNTH_CHAIN", "0") not in ("", "0")
        if chain:
            first |= {u.first + l for u in cand for l in range(1, u.narrays - 1)}
        rest = [i for i in range(len(x)) if i not in first and grads[i] is not None]

        def plain(idx):
            if not idx:
                return
            args = ([x[i] for i in idx], [m[i] for i in idx], [v[i] for i in idx], [grads[i] for i in idx])
            if alpha_dev is not None:
                native.adam_step_dev(*args, alpha_dev, omb1, omb2, eps)
            else:
                native.adam_step(*args, alpha, omb1, omb2, eps)

        plain(rest)
        if self.slab is not None:
            # slabs: the coarser terms' halos are exchanged, the coarse levels synthesised on their extended ranges, the
            # finest term is updated on the OWNED planes (its halos are never needed again) and the halo of U is
            # exchanged instead -- the same bytes as the exchange of the finest term it replaces
            slab, H = self.slab, self.slab.halo
            slab.exchange([x[u.first + i] for u in cand for i in range(1, u.narrays)], width=H)
            done_u = []
            for u in cand:
                i0 = u.first
                a = x[i0: i0 + u.narrays]
                bufs = self._buffers.setdefault(("slabV", u.key), {})
                res, cfac = synthesize_slab(slab, a, u.shapes, u.factors, u.mgloc, bufs, exchanged=True, stop_level=1)
                out0 = bufs.get(("V", 0))
                if out0 is None or tuple(out0.shape) != tuple(a[0].shape):
                    out0 = bufs[("V", 0)] = torch.zeros_like(a[0])
                zf, nf = slab.owned_range(u.shapes[0])
                zc, _ = slab.owned_range(u.shapes[1])
                done = grads[i0] is not None and u.key not in self.periodic_keys and native.adam_synth(
                    u.shapes[1], u.mgloc, res, cfac, u.factors[0], x[i0], m[i0], v[i0], grads[i0], out0, alpha, omb1,
                    omb2, eps, alpha_dev, rng=(zf, zf + nf, zf - H, zc - H))
                if done:
                    done_u.append((u, a, out0))
                else:
                    self._synth_cache.pop(u.key, None)
                    if grads[i0] is not None:
                        plain([i0])
            if done_u:
                slab.exchange([out0 for _, _, out0 in done_u], width=H)
            for u, a, out0 in done_u:
                self._synth_cache[u.key] = (self._signature(a), out0)
            return
        for u in cand:
            i0, L = u.first, u.narrays
            a = x[i0: i0 + L]
            res, cfac = a[L - 1], u.factors[L - 1]
            for lvl in range(L - 2, 0, -1):
                out = self._buf(("V", u.key, lvl), u.shapes[lvl])
                i = i0 + lvl
                if not (chain and grads[i] is not None and native.adam_synth(
                        u.shapes[lvl + 1], u.mgloc, res, cfac, u.factors[lvl], x[i], m[i], v[i], grads[i], out, alpha,
                        omb1, omb2, eps, alpha_dev)):
                    if chain and grads[i] is not None:
                        plain([i])
                    native.mg_interp_add(u.shapes[lvl + 1], u.mgloc, res, cfac, a[lvl], u.factors[lvl], out)
                res, cfac = out, 1.0
            out0 = self._buf(("V", u.key, 0), u.shapes[0])
            done = grads[i0] is not None and native.adam_synth(
                u.shapes[1], u.mgloc, res, cfac, u.factors[0], x[i0], m[i0], v[i0], grads[i0], out0, alpha, omb1, omb2,
                eps, alpha_dev)
            if done:
                self._synth_cache[u.key] = (self._signature(a), out0)
            else:
                self._synth_cache.pop(u.key, None)
                if grads[i0] is not None:
                    plain([i0])

    def request_fused_adam(self, x, m, v, alpha, omb1, omb2, eps, alpha_dev=None):
        """Asks the NEXT loss_grad() to apply the Adam update of the finest multigrid term itself, inside the
        transposed interpolation that streams that term's gradient (odil_b200_mg_interp_adjoint_adam); the gradient
        entry of an updated array comes back as None.  x, m, v: the optimizer's lists, aligned with `arrays`."""
        self._fused_adam = dict(x=x, m=m, v=v, alpha=alpha, omb1=omb1, omb2=omb2, eps=eps, alpha_dev=alpha_dev)

    def _scatter_grad(self, unk, gU, grads):
        """Gradient of the regular field -> gradients of the stored arrays."""
        if unk.kind != "MultigridField":
            grads[unk.first] = gU
            return
        g = gU
        fused = getattr(self, "_fused_adam", None)
        for lvl in range(unk.narrays):
            if lvl > 0:
                gc = self._buf(("gV", unk.key, lvl), unk.shapes[lvl])
                done = False
                if lvl == 1 and fused is not None and unk.factors[0] == 1 and self.slab is None:
                    i = unk.first
                    done = native.mg_interp_adjoint_adam(unk.shapes[1], unk.mgloc, g, 1.0, gc, fused["x"][i],
                                                         fused["m"][i], fused["v"][i], fused["alpha"], fused["omb1"],
                                                         fused["omb2"], fused["eps"], fused["alpha_dev"])
                    if done:
                        grads[i] = None  # already applied
                if not done:
                    native.mg_interp_adjoint(unk.shapes[lvl], unk.mgloc, g, 1.0, gc)
                g = gc
            f = unk.factors[lvl]
            grads[unk.first + lvl] = g if f == 1 else g * f

    # ----------------------------------------------------------------------------------------------
    def loss_grad(self, arrays):
        self._check_arrays(arrays)
        if self.slab is not None:
            return self._loss_grad_slab(arrays)
        K = len(self.outputs)
        sums = torch.empty(K, dtype=torch.float64, device=self.device)
        U = {key: self._regular(self.unknowns[key], arrays) for key in self.used_keys | self._frozen_keys()}
        gU = {}
        for k, out in enumerate(self.outputs):
            if out.fused:
                blk = out.blocks[0]
                g = self._buf(("gU", blk.key), out.shape)
                blk.plan.fused(U[blk.key], out.const, 2.0 / out.n, g, sums[k:k + 1])
                gU[blk.key] = g
                continue
            if not out.blocks:
                native.sum_squares(out.const, sums[k:k + 1])
                continue
            F = self._buf(("F", k), out.shape)
            src = out.const
            for blk in out.blocks:
                blk.plan.forward(U[blk.key], src, F)
                src = F
            native.sum_squares(F, sums[k:k + 1])
            for blk in out.blocks:
                if blk.frozen:
                    continue
                g = self._buf(("gU", blk.key), out.shape)
                blk.plan.adjoint(F, 2.0 / out.n, g if blk.key in gU else None, g)
                gU[blk.key] = g
        grads = [None] * self.narrays
        applied = set()
        for key, unk in self.unknowns.items():
            if key in gU:
                self._scatter_grad(unk, gU[key], grads)
                if grads[unk.first] is None:
                    applied.add(unk.first)
        self._fused_adam = None
        for i in range(self.narrays):
            if grads[i] is None and i not in applied:
                grads[i] = torch.zeros_like(arrays[i])
        fetch = _Fetch(sums, [o.n for o in self.outputs], self.dtype)
        loss = LazyScalar(fetch, "loss")
        terms = [LazyScalar(fetch, "term", k) for k in range(K)]
        norms = [LazyScalar(fetch, "norm", k) for k in range(K)]
        return loss, grads, terms, norms

    def _frozen_keys(self):
        return {b.key for o in self.outputs for b in o.blocks if b.frozen}

    def operator_values(self, arrays):
        """Materialised operator outputs F_k (Problem.eval_operator, core.py:1298-1311)."""
        self._check_arrays(arrays)
        if self.slab is not None:
            raise NotImplementedError("eval_operator on slab-decomposed grids")
        U = {key: self._regular(self.unknowns[key], arrays) for key in self.used_keys | self._frozen_keys()}
        res = []
        for out in self.outputs:
            if not out.blocks:
                res.append(Known(out.const.clone()))
                continue
            F = torch.empty(out.shape, dtype=self.tdtype, device=self.device)
            src = out.const
            for blk in out.blocks:
                blk.plan.forward(U[blk.key], src, F)
                src = F
            res.append(Known(F))
        return res
