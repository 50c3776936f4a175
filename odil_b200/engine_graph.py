"""
GraphEngine: loss, gradient, operator values and Jacobian products of operators that are not affine stencils
(SURVEY.md 8f-1, 8f-2, 8f-3).  `Problem` falls back to it when the affine tracer raises NonAffineError.

    reference                                         here
    ------------------------------------------------  -------------------------------------------------------------
    jax.jit trace of operator(ctx) (core.py:1106)     one trace into odil_b200.graph nodes (Context(trace=...))
    XLA program of eval_loss + value_and_grad         one NVRTC-compiled kernel per output shape (codegen 'lossgrad')
    multigrid_to_regular / its AD                     the hand-written transfer kernels (shared with ResidualEngine)
    tracers as jit arguments (core.py:1076-1110)      run-time scalars in the kernel's parameter block
    _eval_operator_grad_tf + linearize (TF only)      'jvp' / 'vjp' kernels (matrix-free) and 'jac' (COO rows -> CSR)

Supported unknowns: Field, MultigridField (through its synthesised regular field, or its terms directly), Array,
NeuralNet.  Not decomposed into slabs.
"""
import math
import os

import numpy as np
import torch

from . import codegen, graph, native
from .backend import Known, NonAffineError
from .core import Array, Context, Field, MultigridField, NeuralNet, State
from .engine import LazyScalar, ResidualEngine, _Fetch


class ParamDict(dict):
    """`ctx.tracers` during a general trace: numeric scalars become run-time parameters of the generated kernels
    (the reference passes tracers as arguments of the jitted function, core.py:1076-1110)."""

    def __init__(self, values, dtype):
        super().__init__(values or {})
        self.dtype = dtype
        self.used = []

    def __getitem__(self, key):
        v = super().__getitem__(key)
        if isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool):
            if key not in self.used:
                self.used.append(key)
            return graph.g_param(key, self.dtype)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default


class GraphTrace:
    """Input registry of one trace.  Slots 0 .. narrays-1 are the state arrays in `arrays_from_state` order; further
    slots are the synthesised regular fields of multigrid unknowns."""

    def __init__(self, engine, state):
        self.engine = engine
        domain = engine.domain
        self.shapes = []
        self.exprs = []
        for key, unk in engine.unknowns.items():
            for shape in unk.shapes:
                self.exprs.append(graph.g_input(len(self.shapes), shape, engine.dtype))
                self.shapes.append(tuple(shape))
        self.regular_slot = {}
        self.fields_read = {}  # (key, shift, loc) -> (source loc, pad flags, trim flags): what ctx.field() produced
        # shadow state: same containers, arrays replaced by graph inputs
        fields = {}
        for key, field in state.fields.items():
            unk = engine.unknowns[key]
            ex = self.exprs[unk.first: unk.first + unk.narrays]
            if isinstance(field, MultigridField):
                terms = [Field(e, loc=t.loc, cshape=t.cshape) for e, t in zip(ex, field.terms)]
                fields[key] = MultigridField(terms, loc=field.loc, factors=field.factors, axes=field.axes,
                                             method=field.method)
            elif isinstance(field, Field):
                fields[key] = Field(ex[0], loc=field.loc, cshape=field.cshape)
            elif isinstance(field, Array):
                fields[key] = Array(ex[0], field.shape)
            elif isinstance(field, NeuralNet):
                nw = len(field.weights)
                fields[key] = NeuralNet(ex[:nw], ex[nw:], func_in=field.func_in, func_out=field.func_out,
                                        activation=field.activation)
        self.state = State(fields=fields, initialized=True)

    def note_field(self, desc, src_loc, pad_flag, trim_flag):
        self.fields_read[desc] = (src_loc, tuple(pad_flag), tuple(trim_flag))

    def column_map(self, desc, col0):
        """Packed-state column that ctx.field(*desc) reads in every cell of its result (-1 where it reads the zero
        pad): the index array of the source field pushed through the same pad / roll / trim (core.py:956-969)."""
        key, shift, loc = desc
        unk = self.engine.unknowns[key]
        src_loc, pad_flag, trim_flag = self.fields_read[desc]
        idx = np.arange(math.prod(unk.shapes[0]), dtype=np.int64).reshape(unk.shapes[0]) + int(col0[unk.first])
        if any(pad_flag):
            idx = np.pad(idx, [(1, 0) if f else (0, 0) for f in pad_flag], mode="constant", constant_values=-1)
        if any(shift):
            idx = np.roll(idx, tuple(int(-s) for s in shift), axis=tuple(range(idx.ndim)))
        if any(trim_flag):
            idx = idx[tuple(slice(0, -1 if f else None) for f in trim_flag)]
        return idx

    def regular(self, key):
        unk = self.engine.unknowns[key]
        if unk.kind != "MultigridField" or unk.narrays == 1 and unk.factors[0] == 1:
            return self.exprs[unk.first]
        if key not in self.regular_slot:
            self.regular_slot[key] = len(self.shapes)
            self.exprs.append(graph.g_input(len(self.shapes), unk.shapes[0], self.engine.dtype))
            self.shapes.append(tuple(unk.shapes[0]))
        return self.exprs[self.regular_slot[key]]


class _Out:
    def __init__(self, name, shape, raw):
        self.name, self.shape, self.raw = name, tuple(shape), raw
        self.n = math.prod(self.shape)
        self.blocks, self.fused, self.const = [], False, None


class GraphEngine(ResidualEngine):

    # ----------------------------------------------------------------------------------------------
    def _trace(self, state):
        problem, domain = self.problem, self.domain
        if self.slab is not None:
            raise NonAffineError("operators that are not affine stencils are not decomposed into slabs")
        self.trace = trace = GraphTrace(self, state)
        self.params = ParamDict(problem.tracers if isinstance(problem.tracers, dict) else None, self.dtype)
        from .engine import TracerView

        self.tracer_view = TracerView(None)  # nothing is baked in: tracers are run-time parameters here
        ctx = Context(domain, trace.state, extra=problem.extra, tracers=self.params, trace=trace)
        ff = problem.operator(ctx)
        assert isinstance(ff, (tuple, list)) and len(ff), "Operator must return a non-empty list"
        names = [f[0] if isinstance(f, tuple) else "" for f in ff]
        nonempty = [n for n in names if n]
        assert len(nonempty) == len(set(nonempty)), "Name of fields must be unique, got {}".format(nonempty)
        values = [f[1] if isinstance(f, tuple) else f for f in ff]
        self.names = names
        outs, self.outputs = [], []
        for name, v in zip(names, values):
            raw = isinstance(v, Context.Raw)
            if raw:
                v = v.value
            e = graph.node(v, self.dtype)
            if e.kind != "f":
                e = graph.g_unary("cast_f", e)
            outs.append((e, raw))
            self.outputs.append(_Out(name, e.shape, raw))
        self.used_keys = set()
        grad_slots = set(range(len(trace.shapes)))
        self.gen = codegen.Generator(outs, "float" if self.dtype == np.float32 else "double", trace.shapes, grad_slots,
                                     self.params.used)
        # `gen.params` may have grown by parameters first met during emission; shapes may have grown by regular slots
        self.modules = {}
        self.sources = {}
        self.consts_dev = None

    # ----------------------------------------------------------------------------------------------
    def source(self, mode):
        if mode not in self.sources:
            self.sources[mode] = self.gen.source(mode)
        return self.sources[mode]

    def module(self, mode):
        if mode not in self.modules:
            self.modules[mode] = native.JitModule(self.source(mode))
        return self.modules[mode]

    def _consts(self):
        if self.consts_dev is None:
            dev = []
            for t, strides, is_bool in self.gen.consts:
                t = t.to(self.device)
                dev.append((t.to(torch.uint8) if is_bool else t.to(self.tdtype)).contiguous())
            self.consts_dev = dev
        return self.consts_dev

    def _inputs(self, arrays):
        """Device tensor per input slot: state arrays, then synthesised regular fields."""
        ins = list(arrays)
        for key, slot in self.trace.regular_slot.items():
            ins.append(self._regular(self.unknowns[key], arrays))
        return ins

    def _prm(self):
        tr = self.problem.tracers if isinstance(self.problem.tracers, dict) else {}
        return [1.0 / o.n for o in self.outputs] + [float(tr[name]) for name in self.gen.params]

    def _nblk(self):
        return max(1, min(codegen.MAX_GRID, max((g.ncell + codegen.BLOCK - 1) // codegen.BLOCK
                                                 for g in self.gen.groups)))

    def _launch(self, mode, ins, gin=None, tin=None, out=None, seed=None, partials=None, sums=None, prm=None,
                jcol=None, jval=None, colbase=None, only=None):
        mod = self.module(mode)
        nblk = self._nblk()
        ptr = lambda xs: [x.data_ptr() if x is not None else 0 for x in xs] if xs is not None else []
        for g, name, which in self.gen.kernels(mode):
            if only is not None and (g.gid, which) != only:
                continue
            params = self.gen.pack(g.ncell, nblk, ptr(ins), ptr(gin), ptr(tin), colbase or [], ptr(self._consts()),
                                   ptr(out), ptr(seed), partials.data_ptr() if partials is not None else 0,
                                   sums.data_ptr() if sums is not None else 0,
                                   jcol.data_ptr() if jcol is not None else 0,
                                   jval.data_ptr() if jval is not None else 0, prm if prm is not None else self._prm())
            grid = nblk if mode == "lossgrad" else max(1, min(codegen.MAX_GRID * 2,
                                                               (g.ncell + codegen.BLOCK - 1) // codegen.BLOCK))
            mod.launch("k_" + name, grid, codegen.BLOCK, params)
        if mode == "lossgrad":
            params = self.gen.pack(0, nblk, [], [], [], [], [], [], [], partials.data_ptr(), sums.data_ptr(), 0, 0, [])
            mod.launch("k_reduce", len(self.outputs), codegen.BLOCK, params)

    def _zero_grads(self, arrays, name):
        """Zeroed accumulation targets per input slot (gradients are scattered with atomicAdd)."""
        res = []
        for i, shape in enumerate(self.trace.shapes):
            b = self._buf((name, i), shape)
            b.zero_()
            res.append(b)
        return res

    def _fold_regular(self, gin, grads):
        """Adds the gradient of each synthesised regular field to the gradients of its multigrid terms."""
        for key, slot in self.trace.regular_slot.items():
            unk = self.unknowns[key]
            tmp = [None] * self.narrays
            self._scatter_grad(unk, gin[slot], tmp)
            for i in range(unk.first, unk.first + unk.narrays):
                if tmp[i] is not None:
                    grads[i].add_(tmp[i])

    # ----------------------------------------------------------------------------------------------
    def loss_grad(self, arrays):
        self._check_arrays(arrays)
        K = len(self.outputs)
        ins = self._inputs(arrays)
        gin = self._zero_grads(arrays, "g")
        sums = torch.empty(K, dtype=torch.float64, device=self.device)
        partials = self._buf64("partials", K * self._nblk())
        self._launch("lossgrad", ins, gin=gin, partials=partials, sums=sums)
        grads = gin[: self.narrays]
        self._fold_regular(gin, grads)
        fetch = _Fetch(sums, [o.n for o in self.outputs], self.dtype, raws=[o.raw for o in self.outputs])
        return (LazyScalar(fetch, "loss"), grads, [LazyScalar(fetch, "term", k) for k in range(K)],
                [LazyScalar(fetch, "norm", k) for k in range(K)])

    def _buf64(self, name, n):
        b = self._buffers.get(name)
        if b is None or b.numel() != n:
            b = torch.zeros(n, dtype=torch.float64, device=self.device)
            self._buffers[name] = b
        return b

    def operator_values(self, arrays):
        self._check_arrays(arrays)
        ins = self._inputs(arrays)
        out = [torch.empty(o.shape, dtype=self.tdtype, device=self.device) for o in self.outputs]
        self._launch("values", ins, out=out)
        return [Known(t) for t in out]

    def jacobian(self, arrays):
        return GraphJacobian(self, arrays)

    def operator_grad(self, arrays):
        """Per output: {(key, shift, loc): dF/d(that shifted field), an array on the output's grid} for every
        ctx.field() the operator made, and {(key, None, None): dense block} for Array unknowns -- what the reference's
        `_eval_operator_grad_tf` returns (core.py:1313-1361), read off the rows the 'jac' kernels write."""
        jac = GraphJacobian(self, arrays)
        return diagonals_from_rows(self, jac.col0, jac.rows())


def diagonals_from_rows(engine, col0, rows_per_output):
    """rows_per_output[k] = (cell index, packed column, value) arrays of output k (COO rows of the Jacobian)."""
    trace = engine.trace
    maps = {desc: trace.column_map(desc, col0) for desc in trace.fields_read}
    res = []
    for k, out in enumerate(engine.outputs):
        cell, col, val = rows_per_output[k]
        d = {}
        taken = np.zeros(len(cell), dtype=bool)
        for desc, cmap in maps.items():
            if tuple(cmap.shape) != tuple(out.shape):
                continue
            hit = (~taken) & (cmap.reshape(-1)[cell] == col)
            g = np.zeros(out.n, dtype=np.float64)
            np.add.at(g, cell[hit], val[hit])
            taken |= hit
            if np.any(g != 0) or desc[1] == (0,) * len(desc[1]):
                d[desc] = Known(torch.as_tensor(g.reshape(out.shape), dtype=engine.tdtype))
        for key, unk in engine.unknowns.items():
            if unk.kind != "Array":
                continue
            lo, n = int(col0[unk.first]), math.prod(unk.shapes[0])
            hit = (~taken) & (col >= lo) & (col < lo + n)
            if np.any(hit):
                block = np.zeros((out.n, n), dtype=np.float64)
                np.add.at(block, (cell[hit], col[hit] - lo), val[hit])
                d[(key, None, None)] = Known(torch.as_tensor(block.reshape(out.shape + tuple(unk.shapes[0])),
                                                             dtype=engine.tdtype))
                taken |= hit
        res.append(d)
    return res


class GraphJacobian:
    """J = dF/d(packed state) at a fixed state: matrix-free products on the device (`matvec`, `rmatvec`: the
    generated forward- and reverse-mode kernels), the residual vector, and `tocsr()` = the matrix the reference's
    `linearize` assembles (core.py:1113-1217), from per-cell partial derivatives written by the 'jac' kernels.
    Behaves like a SciPy sparse matrix where the reference's scripts use one (`.T`, `@`, `.dot`)."""

    def __init__(self, engine, arrays):
        for key, unk in engine.unknowns.items():
            if unk.kind == "MultigridField" and unk.narrays > 1:
                raise NotImplementedError("Newton needs multigrid off (as in the reference, examples/wave/README.md:27)")
        self.engine = engine
        self.dtype, self.device = engine.tdtype, engine.device
        self.arrays = [a.clone() for a in arrays]
        self.prm = engine._prm()
        self.sizes = [a.numel() for a in self.arrays]
        self.col0 = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.row0 = np.concatenate([[0], np.cumsum([o.n for o in engine.outputs])]).astype(np.int64)
        self.shape = (int(self.row0[-1]), int(self.col0[-1]))
        self._csr = None
        self._dia = None  # per group: stored diagonals [npairs * ncell], built on the first product

    def _diagonals(self):
        """Per group the arrays D[pair][cell] = dF_k/d(load) at the fixed state (codegen mode 'jacd'), or False when the
        operator does not allow it / they would not fit (ODIL_B200_NEWTON_DIA=0 disables, ODIL_B200_NEWTON_DIA_GB caps the
        memory, default 24).  A Newton step evaluates the operator's arithmetic once here instead of once per CG product."""
        if self._dia is None:
            eng = self.engine
            self._dia = False
            if os.environ.get("ODIL_B200_NEWTON_DIA", "1") not in ("", "0") and eng.gen.dia_ok():
                sizes = {g.gid: len(g.pairs()) * g.ncell for g in eng.gen.groups}
                cap = float(os.environ.get("ODIL_B200_NEWTON_DIA_GB", "24")) * 2 ** 30
                if sum(sizes.values()) * self.arrays[0].element_size() <= cap:
                    dia = {}
                    for g in eng.gen.groups:
                        dia[g.gid] = torch.empty(max(1, sizes[g.gid]), dtype=self.dtype, device=self.device)
                        eng._launch("jacd", self.arrays, jval=dia[g.gid], prm=self.prm, only=(g.gid, None))
                    self._dia = dia
        return self._dia

    def _split(self, x, starts, shapes):
        return [x[int(starts[i]): int(starts[i + 1])].view(shapes[i]) for i in range(len(shapes))]

    def matvec(self, x):
        eng = self.engine
        x = x.to(self.dtype).contiguous()
        tin = self._split(x, self.col0, [tuple(a.shape) for a in self.arrays])
        y = torch.empty(self.shape[0], dtype=self.dtype, device=self.device)
        out = self._split(y, self.row0, [o.shape for o in eng.outputs])
        dia = self._diagonals()
        if dia:
            for g in eng.gen.groups:
                eng._launch("jvpd", self.arrays, tin=tin, out=out, jval=dia[g.gid], prm=self.prm, only=(g.gid, None))
        else:
            eng._launch("jvp", self.arrays, tin=tin, out=out, prm=self.prm)
        return y

    def rmatvec(self, y):
        eng = self.engine
        y = y.to(self.dtype).contiguous()
        seed = self._split(y, self.row0, [o.shape for o in eng.outputs])
        x = torch.zeros(self.shape[1], dtype=self.dtype, device=self.device)
        gin = self._split(x, self.col0, [tuple(a.shape) for a in self.arrays])
        dia = self._diagonals()
        if dia:
            mode = "vjpg" if eng.gen.gather_ok() and os.environ.get("ODIL_B200_NEWTON_GATHER", "1") not in ("", "0") \
                else "vjpd"
            for g in eng.gen.groups:
                eng._launch(mode, self.arrays, gin=gin, seed=seed, jval=dia[g.gid], prm=self.prm, only=(g.gid, None))
        else:
            eng._launch("vjp", self.arrays, gin=gin, seed=seed, prm=self.prm)
        return x

    def dot(self, x):
        if torch.is_tensor(x):
            return self.matvec(x)
        return self.tocsr().dot(x)

    def rows(self):
        """Per output k: (cell, column, value) of the non-zero Jacobian entries, from the 'jac' kernels."""
        eng = self.engine
        if self.shape[0] * 8 > 2 ** 31:
            raise MemoryError("explicit Jacobian rows are meant for small problems; use the matrix-free products")
        colbase = [int(c) for c in self.col0[:-1]]
        empty = (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0))
        res = [empty] * len(eng.outputs)
        for g, name, which in eng.gen.kernels("jac"):
            k = g.results[which][0]
            nl = eng.gen.nloads(g)
            if nl == 0:
                continue
            jcol = torch.zeros(g.ncell * nl, dtype=torch.int64, device=self.device)
            jval = torch.zeros(g.ncell * nl, dtype=self.dtype, device=self.device)
            eng._launch("jac", self.arrays, jcol=jcol, jval=jval, colbase=colbase, prm=self.prm, only=(g.gid, which))
            v = jval.cpu().numpy().astype(np.float64)
            c = jcol.cpu().numpy()
            r = np.repeat(np.arange(g.ncell, dtype=np.int64), nl)
            keep = v != 0
            res[k] = (r[keep], c[keep], v[keep])
        return res

    def tocsr(self):
        import scipy.sparse

        if self._csr is not None:
            return self._csr
        rows, cols, vals = [], [], []
        for k, (r, c, v) in enumerate(self.rows()):
            if len(r):
                rows.append(r + int(self.row0[k]))
                cols.append(c)
                vals.append(v)
        if rows:
            m = scipy.sparse.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                        shape=self.shape)
        else:
            m = scipy.sparse.coo_matrix(self.shape)
        self._csr = m.tocsr()
        return self._csr

    # SciPy-matrix surface used by reference scripts (tests/test_newton.py: `matrix.T @ matrix`, `matrix.T @ vector`)
    @property
    def T(self):
        return self.tocsr().T

    def __matmul__(self, other):
        return self.tocsr() @ other

    def __rmatmul__(self, other):
        return other @ self.tocsr()

    def toarray(self):
        return self.tocsr().toarray()
