"""
Grid, state containers, operator context and the Problem engine seam.

Host-side mirror of the reference's L1/L2 layers (src/odil/core.py): same class names, argument
meaning and error behaviour, so problem scripts written for cselab/odil run unchanged, while the
per-iteration arithmetic goes to hand-written CUDA (odil_b200.engine -> libodil_b200.so).

Reference anchors: Domain core.py:11-503, containers :506-603, transfers :606-755,
Context :865-990, Problem :993-1386, checkpoint :1389-1436, extrapolation helpers :1439-1457.
"""
import math
import pickle

import numpy as np
import torch

from . import native
from . import graph
from .backend import Affine, Known, Lazy, ModB200, NonAffineError, as_known, torch_dtype


def assert_equal(first, second, msg=""):
    if not (first == second):
        raise ValueError("Expected equal '{:}' and '{:}'{}".format(first, second, msg))


# --------------------------------------------------------------------------------------------------
# Containers (core.py:506-603)
# --------------------------------------------------------------------------------------------------
class Field:
    """Grid field. `loc`: one char per axis, 'c' cell centre / 'n' node. `cshape`: grid size in cells."""

    def __init__(self, array=None, loc=None, cshape=None):
        self.array = array
        self.loc = loc
        self.cshape = cshape

    def __repr__(self):
        return "odil.Field({}, loc='{}', cshape={:})".format(repr(self.array), self.loc, self.cshape)

    __str__ = __repr__


class MultigridField:
    """u = f0 t0 + I(f1 t1 + I(...)): `terms` are Fields from the finest to the coarsest level."""

    def __init__(self, terms=None, loc=None, factors=None, axes=None, method=None):
        self.terms = terms
        self.loc = loc
        self.factors = factors
        self.axes = axes
        self.method = method


class NeuralNet:
    """Fully connected network: weights[i] has shape (n_out, n_in), biases[i] shape (n_out,)."""

    def __init__(self, weights=None, biases=None, func_in=None, func_out=None, activation=None):
        self.weights = weights
        self.biases = biases
        self.func_in = func_in
        self.func_out = func_out
        self.activation = activation or "tanh"


class Array:
    """Unknown array that does not live on the grid."""

    def __init__(self, array=None, shape=None):
        self.array = array
        self.shape = shape

    def __repr__(self):
        return "odil.Array({}, shape={:})".format(repr(self.array), self.shape)

    __str__ = __repr__


class State:
    """Ordered mapping key -> Field | MultigridField | NeuralNet | Array; the order defines the unknown vector."""

    def __init__(self, fields=None, initialized=False):
        self.fields = fields if fields is not None else dict()
        self.initialized = initialized


# --------------------------------------------------------------------------------------------------
# Multigrid transfers: public functions of the reference (core.py:606-755) on the CUDA kernels
# --------------------------------------------------------------------------------------------------
def _device_array(u, mod, dtype=None):
    """Contiguous CUDA tensor from Known / numpy / torch."""
    if isinstance(u, graph.Expr):
        raise graph.GraphError("multigrid interpolation of a traced expression inside an operator is not supported")
    if isinstance(u, Affine):
        raise NonAffineError("multigrid transfer of a field expression inside an operator is not on the fused path")
    if isinstance(u, Known):
        t = u.full()
    elif torch.is_tensor(u):
        t = u
    else:
        t = torch.from_numpy(np.ascontiguousarray(u))
    if dtype is not None:
        t = t.to(torch_dtype(dtype))
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise native.NativeError("multigrid transfers run on the GPU only (no CPU fallback)")
        t = t.cuda()
    return t.contiguous()


def _fine_shape(shape, loc):
    return tuple({"c": 2 * n, "n": 2 * (n - 1) + 1, ".": n}[l] for n, l in zip(shape, loc))


def _coarse_shape(shape, loc):
    return tuple({"c": n // 2, "n": (n - 1) // 2 + 1, ".": n}[l] for n, l in zip(shape, loc))


def interp_to_finer(u, loc=None, method=None, mod=None, depth=1):
    """
    Interpolates a field to the next finer grid (core.py:606-700).  'c' axes double, 'n' axes go to
    2(n-1)+1, '.' axes keep their size.  `method` ('conv' | 'stack') is accepted for compatibility;
    both reference variants give the same numbers and map to one kernel.
    """
    if depth == 0:
        return u
    method = method or "stack"
    if method not in ["conv", "stack"]:
        raise ValueError("Unknown method='{}'".format(method))
    assert_equal(len(loc), len(u.shape))
    for l in loc:
        assert l in "cn.", "Invalid loc={}".format(loc)
    t = _device_array(u, mod)
    out = torch.empty(_fine_shape(t.shape, loc), dtype=t.dtype, device=t.device)
    native.mg_interp_add(tuple(t.shape), loc, t, 1.0, None, 0.0, out)
    return interp_to_finer(Known(out), loc, method, mod, depth - 1)


def restrict_to_coarser(u, loc=None, method=None, mod=None, depth=1):
    """Restricts a field to the next coarser grid (core.py:703-755)."""
    if depth == 0:
        return u
    method = method or "conv"
    if method not in ["conv"]:
        raise ValueError("Unknown method='{}'".format(method))
    assert_equal(len(loc), len(u.shape))
    for l in loc:
        assert l in "cn.", "Invalid loc={}".format(loc)
    if isinstance(u, graph.Expr):  # inside an operator (e.g. the `mgloss` outputs of examples/poisson/poisson.py:116-122)
        return restrict_to_coarser(graph.g_restrict(u, loc), loc, method, mod, depth - 1)
    t = _device_array(u, mod)
    out = torch.empty(_coarse_shape(t.shape, loc), dtype=t.dtype, device=t.device)
    native.mg_restrict(tuple(t.shape), loc, t, out)
    return restrict_to_coarser(Known(out), loc, method, mod, depth - 1)


def check_multigrid_cshapes(cshapes, axes=None):
    """Each level must be exactly half of the previous one along the multigrid axes (core.py:758-776)."""
    if not len(cshapes):
        return
    dim = len(cshapes[0])
    axes = axes or [True] * dim
    assert_equal(len(axes), dim)
    for fine, coarse in zip(cshapes[:-1], cshapes[1:]):
        for j in range(dim):
            if axes[j]:
                assert_equal(fine[j], coarse[j] * 2, " with cshapes={:}".format(cshapes))


def synthesize_multigrid(term_arrays, factors, loc):
    """U = f0 t0 + I(f1 t1 + I(...)) on device tensors (core.py:245-263). Returns a new tensor."""
    res = None
    cfac = 1.0
    for lvl in range(len(term_arrays) - 1, -1, -1):
        t = term_arrays[lvl]
        if res is None:
            res, cfac = t, float(factors[lvl])
            continue
        out = torch.empty_like(t)
        native.mg_interp_add(tuple(res.shape), loc, res, cfac, t, float(factors[lvl]), out)
        res, cfac = out, 1.0
    if cfac != 1.0:
        res = res * cfac
    return res


# --------------------------------------------------------------------------------------------------
# Domain (core.py:11-503)
# --------------------------------------------------------------------------------------------------
class Domain:

    def __init__(self, cshape, dimnames=None, lower=0.0, upper=1.0, dtype=None, multigrid=False,
                 mg_convert_all=True, mg_nlvl=None, mg_factors=None, mg_axes=None, mg_interp=None, mod=None):
        ndim = len(cshape)
        dimnames = dimnames or ["x", "y", "z"][:ndim]
        if mod is None:
            from .runtime import mod
        assert_equal(len(dimnames), ndim, f" with dimnames={dimnames}")
        if dtype is None:
            from . import runtime

            dtype = runtime.dtype
        self.ndim = ndim
        self.cshape = cshape
        self.dimnames = dimnames
        self.dtype = np.dtype(dtype)
        self.lower = (np.ones(ndim, dtype=dtype) * lower).astype(dtype)
        self.upper = (np.ones(ndim, dtype=dtype) * upper).astype(dtype)
        self.mod = mod
        # Slab decomposition along axis 0 when launched with one process per GPU (odil_b200.slab).
        from .slab import SlabInfo

        self.slab = SlabInfo.from_environment()
        self.multigrid = multigrid
        if multigrid:
            self.mg_factors = mg_factors
            mg_axes = mg_axes or [True] * ndim
            nlvl_max = min(round(np.log2(n)) if ax else max(cshape) for n, ax in zip(cshape, mg_axes))
            if mg_nlvl is not None:
                assert mg_nlvl >= 1
                mg_nlvl = min(mg_nlvl, nlvl_max)
            else:
                mg_nlvl = nlvl_max
            self.mg_nlvl = mg_nlvl
            self.mg_cshapes = [tuple(n >> lvl if ax else n for n, ax in zip(cshape, mg_axes))
                               for lvl in range(mg_nlvl)]
            check_multigrid_cshapes(self.mg_cshapes, mg_axes)
            self.mg_axes = mg_axes
            self.mg_interp = mg_interp
            self.mg_convert_all = mg_convert_all

    # -- geometry -------------------------------------------------------------------------------
    @staticmethod
    def _names_to_indices(dims, dimnames):
        sel = dims if dims is not None and len(dims) else range(len(dimnames))
        return tuple(dimnames.index(d) if isinstance(d, str) else d for d in sel)

    def cast(self, value, dtype=None):
        return self.mod.cast(value, dtype or self.dtype)

    def _points_1d(self, d, loc):
        n = self.cshape[d]
        if loc == "c":
            x = np.linspace(self.lower[d], self.upper[d], n, endpoint=False, dtype=self.dtype)
            if n > 1:
                x += (x[1] - x[0]) * 0.5
            return x
        if loc == "n":
            return np.linspace(self.lower[d], self.upper[d], n + 1, dtype=self.dtype)
        raise ValueError("Unknown loc=" + loc)

    def _indices_1d(self, d, loc):
        if loc == "c":
            return np.arange(self.cshape[d], dtype=int)
        if loc == "n":
            return np.arange(self.cshape[d] + 1, dtype=int)
        raise ValueError("Unknown loc=" + loc)

    def points_1d(self, *dims, loc=None):
        loc = loc or "c" * self.ndim
        idims = self._names_to_indices(dims, self.dimnames)
        res = [self._points_1d(i, c) for i, c in zip(idims, loc)]
        return res[0] if len(dims) == 1 else res

    def _grid(self, gen, dims, loc, skip):
        loc = loc or "c" * self.ndim
        assert_equal(len(loc), self.ndim, f" with loc={loc}")
        names = [v for v, c in zip(self.dimnames, loc) if c not in skip]
        idims = self._names_to_indices(dims, names)
        xx = [gen(d, loc[d]) for d in range(self.ndim) if loc[d] not in skip]
        data = self.mod.meshgrid(*xx, indexing="ij")
        res = tuple(data[i] for i in idims)
        return res[0] if len(dims) == 1 else res

    def points(self, *dims, loc=None):
        return self._grid(self._points_1d, dims, loc, ".")

    def indices(self, *dims, loc=None):
        return self._grid(self._indices_1d, dims, loc, ".")

    @staticmethod
    def _get_field_shape(cshape, loc=None):
        loc = loc or "c" * len(cshape)
        assert all(c in "cn" for c in loc)
        return tuple(int(s) + 1 if c == "n" else int(s) for s, c in zip(cshape, loc))

    def get_field_shape(self, loc=None):
        return self._get_field_shape(self.cshape, loc=loc)

    def size(self, *dims, loc=None):
        loc = loc or "c" * self.ndim
        assert_equal(len(loc), self.ndim, f" with loc={loc}")
        idims = self._names_to_indices(dims, self.dimnames)
        res = []
        for i in idims:
            if loc[i] not in "cn":
                raise ValueError("Unknown loc=" + loc[i])
            res.append(self.cshape[i] + (1 if loc[i] == "n" else 0))
        return res[0] if len(dims) == 1 else res

    def step_by_dim(self, i):
        return (self.upper[i] - self.lower[i]) / self.cshape[i]

    def step(self, *dims):
        idims = self._names_to_indices(dims, self.dimnames)
        res = tuple(self.step_by_dim(i) for i in idims)
        return res[0] if len(dims) == 1 else res

    def random_inner(self, size):
        res = latin_hypercube(self.ndim, size, dtype=self.dtype).T
        return [self.lower[i] + (self.upper[i] - self.lower[i]) * res[i] for i in range(self.ndim)]

    def random_boundary(self, normal, side, size):
        assert normal < self.ndim and side in (0, 1)
        res = latin_hypercube(self.ndim - 1, size, dtype=self.dtype).T
        res = np.vstack((res[:normal], np.ones(size, dtype=self.dtype) * side, res[normal:]))
        return [self.lower[i] + (self.upper[i] - self.lower[i]) * res[i] for i in range(self.ndim)]

    # -- multigrid ------------------------------------------------------------------------------
    def _mg_loc(self, mgfield):
        axes = mgfield.axes or self.mg_axes
        assert_equal(len(axes), len(mgfield.terms[0].cshape))
        return "".join(l if ax else "." for l, ax in zip(mgfield.loc, axes))

    def multigrid_to_regular(self, mgfield):
        """Regular field from multigrid terms (core.py:245-263), on the device."""
        factors = mgfield.factors or self.mg_factors or [1] * len(mgfield.terms)
        assert_equal(len(factors), len(mgfield.terms))
        arrays = [_device_array(t.array, self.mod) for t in mgfield.terms]
        if self.slab is not None:
            from .engine import synthesize_slab

            shapes = [self._get_field_shape(t.cshape, mgfield.loc) for t in mgfield.terms]
            res = self.slab.gather(synthesize_slab(self.slab, arrays, shapes, factors, self._mg_loc(mgfield)))
        else:
            res = synthesize_multigrid(arrays, factors, self._mg_loc(mgfield))
        return Field(Known(res), loc=mgfield.loc)

    def get_regular_array(self, field):
        if isinstance(field, Field) and self.slab is not None and torch.is_tensor(field.array):
            return Known(self.slab.gather(field.array))
        if isinstance(field, (Field, Array)):
            return field.array
        if isinstance(field, MultigridField):
            return self.multigrid_to_regular(field).array
        raise TypeError("Expected Field or MultigridField got {}".format(type(field).__name__))

    def regular_to_multigrid(self, field, cshapes=None, factors=None, method=None):
        """t0 = u / f0, deeper terms zero (core.py:276-297)."""
        mod = self.mod
        if isinstance(field, (MultigridField, NeuralNet)):
            raise TypeError("Expected Field or ndarray, got type {}".format(type(field).__name__))
        field = self.init_field(field)
        cshapes = cshapes or self.mg_cshapes
        factors = factors or self.mg_factors or [1] * len(cshapes)
        assert_equal(len(cshapes), len(factors))
        method = method or self.mg_interp
        first = field.array if factors[0] == 1 else field.array / factors[0]
        terms = [Field(first, loc=field.loc, cshape=field.cshape)]
        for cshape in cshapes[1:]:
            terms.append(Field(self._zeros_storage(self._get_field_shape(cshape, loc=field.loc)), loc=field.loc,
                               cshape=cshape))
        return MultigridField(terms=terms, loc=field.loc, factors=factors, method=method)

    # -- state ----------------------------------------------------------------------------------
    def _zeros_storage(self, shape):
        """Zero state array of GLOBAL field shape `shape` (the local slab with halos when decomposed)."""
        if self.slab is not None:
            shape = self.slab.local_shape(shape)
        return torch.zeros(tuple(shape), dtype=torch_dtype(self.dtype), device=self.mod.device)

    def init_field(self, field):
        """Moves a field description into backend storage, filling defaults (core.py:299-346)."""
        mod = self.mod
        if field is None:
            return self.init_field(Field(None, loc="c" * self.ndim, cshape=self.cshape))
        if isinstance(field, np.ndarray) or mod.is_tensor(field):
            return self.init_field(Field(field, loc="c" * len(field.shape), cshape=tuple(field.shape)))
        if isinstance(field, Field):
            cshape = field.cshape or self.cshape
            loc = field.loc or "c" * len(cshape)
            assert_equal(len(loc), len(cshape))
            shape = self._get_field_shape(cshape, loc=loc)
            array = field.array
            local = self.slab.local_shape(shape) if self.slab is not None else shape
            if array is None:
                array = self._zeros_storage(shape)
            elif torch.is_tensor(array) and tuple(array.shape) == local and self.slab is not None:
                array = mod.variable(array, dtype=self.dtype)  # already a local slab
            else:
                array = mod.variable(array, dtype=self.dtype)
                assert_equal(tuple(array.shape), shape)
                if self.slab is not None:
                    array = self.slab.scatter(array)
            assert_equal(tuple(array.shape), local)
            return Field(array, loc=loc, cshape=cshape)
        if isinstance(field, MultigridField):
            return MultigridField([self.init_field(t) for t in field.terms], loc=field.loc, factors=field.factors,
                                  axes=field.axes, method=field.method)
        if isinstance(field, NeuralNet):
            return NeuralNet([mod.variable(w, dtype=self.dtype) for w in field.weights],
                             [mod.variable(b, dtype=self.dtype) for b in field.biases],
                             func_in=field.func_in, func_out=field.func_out, activation=field.activation)
        if isinstance(field, list):
            u = np.array(field, dtype=self.dtype)
            return self.init_field(Array(u, shape=u.shape))
        if isinstance(field, Array):
            array = field.array
            if array is None:
                array = mod.zeros(field.shape, dtype=self.dtype)
            return Array(mod.variable(array, dtype=self.dtype), field.shape)
        raise TypeError("Unknown field type '{}'".format(type(field).__name__))

    def init_state(self, state):
        fields = dict()
        for key, raw in state.fields.items():
            field = self.init_field(raw)
            if self.multigrid and self.mg_convert_all and not isinstance(field, (MultigridField, NeuralNet, Array)):
                field = self.regular_to_multigrid(field)
            fields[key] = field
        return State(fields=fields, initialized=True)

    def arrays_from_field(self, field):
        if isinstance(field, (Field, Array)):
            return [field.array]
        if isinstance(field, MultigridField):
            return [t.array for t in field.terms]
        if isinstance(field, NeuralNet):
            return list(field.weights) + list(field.biases)
        raise TypeError("Unknown field type '{}'".format(type(field).__name__))

    def arrays_from_state(self, state):
        res = []
        for field in state.fields.values():
            res += self.arrays_from_field(field)
        return res

    @staticmethod
    def arrays_to_field(arrays, field):
        """Stores `arrays` into `field`; returns how many were consumed."""
        if isinstance(field, (Field, Array)):
            field.array = arrays[0]
            return 1
        if isinstance(field, MultigridField):
            for i, term in enumerate(field.terms):
                term.array = arrays[i]
            return len(field.terms)
        if isinstance(field, NeuralNet):
            nw, nb = len(field.weights), len(field.biases)
            field.weights[:] = arrays[:nw]
            field.biases[:] = arrays[nw:nw + nb]
            return nw + nb
        raise TypeError("Unknown field type '{}'".format(type(field).__name__))

    @staticmethod
    def arrays_to_state(arrays, state):
        offset = 0
        for field in state.fields.values():
            offset += Domain.arrays_to_field(arrays[offset:], field)
        return offset

    def _pack(self, arrays):
        """Flat vector of all entries (core.py:427-441).  Returned as a Known so that host arithmetic written for
        the reference (`packed + delta` with a NumPy `delta`, np.array(packed)) works on the device-resident data."""
        parts = [a.reshape(-1) if torch.is_tensor(a) else as_known(a).full().reshape(-1) for a in arrays]
        return Known(torch.cat(parts)) if parts else Known(torch.zeros(0))

    def pack_field(self, field):
        return self._pack(self.arrays_from_field(field))

    def pack_state(self, state):
        return self._pack(self.arrays_from_state(state))

    def _unpack(self, packed, arrays):
        """Splits a flat vector (Known / tensor / ndarray) into storage arrays shaped like `arrays`."""
        sizes = [math.prod(a.shape) for a in arrays]
        flat = packed if torch.is_tensor(packed) else as_known(packed).full()
        flat = flat.reshape(-1)[: sum(sizes)]
        res, start = [], 0
        for n, a in zip(sizes, arrays):
            t = flat[start:start + n].reshape(tuple(a.shape))
            res.append(self.mod.variable(t, dtype=self.dtype) if torch.is_tensor(a) else Known(t))
            start += n
        return res, sum(sizes)

    def unpack_field(self, packed, field):
        arrays, used = self._unpack(packed, self.arrays_from_field(field))
        self.arrays_to_field(arrays, field)
        return used

    def unpack_state(self, packed, state):
        arrays, used = self._unpack(packed, self.arrays_from_state(state))
        self.arrays_to_state(arrays, state)
        return used

    def make_neural_net(self, layers, initializer="lecun", func_in=None, func_out=None, activation=None):
        return make_neural_net(layers, self.dtype, self.mod, initializer, func_in, func_out, activation)

    def field(self, state, key, *shift):
        """Host accessor: regular field `key` shifted like ctx.field (core.py:474-489)."""
        field = state.fields[key]
        if not isinstance(field, (Field, MultigridField, Array)):
            raise TypeError("Expected Field or MultigridField, got type {} for field '{}'".format(
                type(field).__name__, key))
        if isinstance(field, Array):
            if len(shift):
                raise RuntimeError("Array requires an empty shift")
            return as_known(field.array)
        shift = shift or (0,) * self.ndim
        if len(shift) != self.ndim:
            raise RuntimeError("Expected {} shift components, got shift={}".format(self.ndim, shift))
        array = as_known(self.get_regular_array(field))
        return self.mod.roll(array, np.negative(shift), range(self.ndim))

    def neural_net(self, state, key):
        net = state.fields[key]
        if not isinstance(net, NeuralNet):
            raise TypeError("Expected NeuralNet, got type {} for key='{}'".format(type(net).__name__, key))
        return lambda *inputs: eval_neural_net(net, inputs, self.mod)


# --------------------------------------------------------------------------------------------------
# Neural networks (core.py:779-862) -- host/Known evaluation only; NN unknowns are not on the hot path
# --------------------------------------------------------------------------------------------------
def make_neural_net(layers, dtype, mod, initializer="lecun", func_in=None, func_out=None, activation=None):
    scale_of = {"legacy": lambda ni, no: np.sqrt(1.0 / ni), "glorot": lambda ni, no: np.sqrt(6.0 / (ni + no)),
                "lecun": lambda ni, no: np.sqrt(3.0 / ni), "he": lambda ni, no: np.sqrt(6.0 / ni)}
    if initializer not in scale_of:
        raise ValueError("Unknown initializer=" + initializer)
    weights, biases = [], []
    for ni, no in zip(layers[:-1], layers[1:]):
        s = scale_of[initializer](ni, no)
        weights.append(mod.random.uniform(shape=(no, ni), minval=-s, maxval=s, dtype=dtype))
        biases.append(mod.zeros(no, dtype=dtype))
    return NeuralNet(weights, biases, func_in=func_in, func_out=func_out, activation=activation)


def eval_neural_net(net, inputs, mod, frozen=False):
    weights, biases = net.weights, net.biases
    assert_equal(len(weights), len(biases), "Weights and biases do not match")
    assert_equal(weights[0].shape[1], len(inputs), "Weights and inputs do not match")
    act = {"tanh": mod.tanh, "relu": mod.relu, "none": lambda x: x}[net.activation]
    if net.func_in is not None:
        inputs = net.func_in(*inputs)
    if graph.any_expr(*inputs, *weights, *biases):
        # traced evaluation (operator under ctx.neural_net): elementwise over the inputs' common shape
        if frozen:
            weights = [mod.stop_gradient(w) for w in weights]
            biases = [mod.stop_gradient(b) for b in biases]
        outputs = graph.g_mlp(weights, biases, list(inputs), net.activation)
        if net.func_out is not None:
            outputs = net.func_out(*outputs)
        return outputs
    tmp = mod.stack(inputs, axis=0)
    nd = len(tmp.shape)
    tmp = mod.transpose(tmp, list(range(1, nd)) + [0])[..., None]
    for i, (w, b) in enumerate(zip(weights, biases)):
        tmp = mod.matmul(as_known(w), tmp) + as_known(b)[:, None]
        if i < len(weights) - 1:
            tmp = act(tmp)
    tmp = mod.transpose(tmp[..., 0], [nd - 1] + list(range(nd - 1)))
    outputs = [tmp[i] for i in range(tmp.shape[0])]
    if net.func_out is not None:
        outputs = net.func_out(*outputs)
    return outputs


# --------------------------------------------------------------------------------------------------
# Context: what the operator sees (core.py:865-990)
# --------------------------------------------------------------------------------------------------
class Context:

    class Raw:
        """Marks an operator output whose loss term is mean(value) instead of mean(value^2)."""

        def __init__(self, value):
            self.value = value

    def __init__(self, domain, state, watch_func=None, extra=None, tracers=None, distinct_shift=False, trace=None):
        """trace: None = affine tracing (closed-form stencil symbols); an `engine_graph.GraphTrace` = general tracing,
        where `state` is the trace's shadow state whose arrays are graph inputs."""
        self.domain = domain
        self.state = state
        self.trace = trace
        self.watch_func = watch_func or (lambda _: None)
        self.extra = extra
        self.tracers = tracers
        self.dtype = domain.dtype
        self.mod = domain.mod
        self.distinct_shift = distinct_shift
        self.desc_to_array = dict()
        self.key_to_array_jac = dict()
        self.step = domain.step
        self.size = domain.size
        self.indices = domain.indices
        self.points = domain.points

    def cast(self, value, dtype=None):
        return self.mod.cast(value, dtype or self.dtype)

    def field(self, key, *shift, loc=None, frozen=False):
        """
        Symbol for `roll(U_key, -shift)` (core.py:910-975): a NEGATIVE shift component reads the
        LOWER-index neighbour.  Multigrid fields appear as their synthesised regular field.
        """
        domain = self.domain
        field = self.state.fields[key]
        if not isinstance(field, (Field, MultigridField, Array)):
            raise TypeError("Expected Field or MultigridField, got type {} for key='{}'".format(
                type(field).__name__, key))
        if isinstance(field, Array):
            if len(shift):
                raise RuntimeError("Array requires an empty shift")
            if self.trace is not None:
                return self.mod.stop_gradient(field.array) if frozen else field.array
            shape = tuple(field.array.shape)
            return Affine.symbol(key, (0,) * len(shape), shape, self.dtype, frozen, device=self.mod.device)
        shift = tuple(int(s) for s in shift) or (0,) * domain.ndim
        if len(shift) != domain.ndim:
            raise RuntimeError("Expected {} shift components, got shift={}".format(domain.ndim, shift))
        if self.trace is not None:
            return self._field_graph(key, field, shift, loc or field.loc, frozen)
        if loc is not None and loc != field.loc:
            raise NonAffineError("ctx.field(loc=...) with a change of location is not on the fused path yet")
        desc = (key, shift, field.loc, frozen)
        if desc not in self.desc_to_array:
            if isinstance(field, MultigridField):
                cshape = field.terms[0].cshape
            else:
                cshape = field.cshape or domain.cshape
            shape = domain._get_field_shape(cshape, field.loc)
            self.desc_to_array[desc] = Affine.symbol(key, shift, shape, self.dtype, frozen, device=self.mod.device)
        return self.desc_to_array[desc]

    def _field_graph(self, key, field, shift, loc, frozen):
        """General tracing of ctx.field, statement by statement as the reference does it (core.py:934-975): source
        field -> zero pad (1, 0) where a cell-centred axis is read at nodes -> roll by -shift -> trim the last entry
        where a node-centred axis is read at cells -> stop_gradient if frozen."""
        mod, ndim = self.mod, self.domain.ndim
        desc = (key, shift, loc)
        if desc not in self.desc_to_array:
            array = self.trace.regular(key)
            pad_flag = [lf == "c" and l == "n" for lf, l in zip(field.loc, loc)]
            if any(pad_flag):
                array = mod.pad(array, pad_width=[(1, 0) if f else (0, 0) for f in pad_flag], mode="constant")
            if any(shift):
                array = mod.roll(array, np.negative(shift), range(ndim))
            trim_flag = [lf == "n" and l == "c" for lf, l in zip(field.loc, loc)]
            if any(trim_flag):
                array = array[tuple(slice(0, -1 if f else None) for f in trim_flag)]
            self.desc_to_array[desc] = array
            self.trace.note_field(desc, field.loc, pad_flag, trim_flag)
        array = self.desc_to_array[desc]
        return mod.stop_gradient(array) if frozen else array

    def neural_net(self, key, frozen=False):
        net = self.state.fields[key]
        if not isinstance(net, NeuralNet):
            raise TypeError("Expected NeuralNet, got type {} for key='{}'".format(type(net).__name__, key))
        if self.trace is None:
            raise NonAffineError("NeuralNet unknowns are evaluated by the general (graph) engine")
        return lambda *inputs: eval_neural_net(net, inputs, self.mod, frozen=frozen)


# --------------------------------------------------------------------------------------------------
# Problem (core.py:993-1386): the engine seam
# --------------------------------------------------------------------------------------------------
class Problem:

    def __init__(self, operator, domain, extra=None, tracers=None, jit=None):
        """
        operator: callable(ctx) returning a non-empty list of fields (or (name, field) tuples); each
            is one equation F_k = 0 and contributes mean(F_k^2) to the loss.
        """
        self.domain = domain
        self.operator = operator
        self.extra = extra
        if tracers is None:
            tracers = dict()
        if "epoch" not in tracers:
            tracers["epoch"] = 0
        self.tracers = tracers
        self.jit = jit
        self._cache_eval_loss_grad = dict()
        self._cache_eval_operator = dict()
        self._cache_eval_operator_grad = dict()
        mod = domain.mod
        if isinstance(mod, ModB200):
            self._eval_loss_grad = self._eval_loss_grad_b200
            self._eval_operator = self._eval_operator_b200
            self._eval_operator_grad = self._eval_operator_grad_b200
        else:
            raise NotImplementedError("Unsupported mod={:}".format(mod))

    def _engine(self, state):
        cache = self._cache_eval_loss_grad
        if "func" in cache and cache["func"].tracer_view.stale(self.tracers):
            # A tracer the operator read at trace time (e.g. tracers["epoch"] in an annealed weight) has changed:
            # the affine path holds its value inside the coefficient tables, so lower the operator again.
            old = cache.pop("func")
            cache["buffers"] = getattr(old, "_buffers", None)
        if "func" not in cache:
            from .engine import ResidualEngine

            cache["state"] = state
            try:
                cache["func"] = ResidualEngine(self, state)
            except NonAffineError:
                # Not an affine stencil with region-typed coefficients (products / functions of fields, neural nets,
                # location changes, per-cell coefficients, Raw terms ...): trace the operator again into the general
                # expression graph and generate its kernels (odil_b200.engine_graph).
                from .engine_graph import GraphEngine

                cache["func"] = GraphEngine(self, state)
            if cache.get("buffers"):
                cache["func"]._buffers = cache.pop("buffers")  # work arrays of the previous lowering
            cache["names"] = cache["func"].names
        return cache["func"]

    def _eval_loss_grad_b200(self, state):
        engine = self._engine(state)
        loss, grads, terms, norms = engine.loss_grad(self.domain.arrays_from_state(state))
        return loss, grads, terms, engine.names, norms

    def _eval_operator_b200(self, state):
        engine = self._engine(state)
        return engine.operator_values(self.domain.arrays_from_state(state)), engine.names

    def _eval_operator_grad_b200(self, state):
        """(values, [ {(key, shift, loc): coefficient array} per output ], names) like `_eval_operator_grad_tf`
        (core.py:1313-1361); affine operators only: the derivatives are the region-typed tables on the grid."""
        from .newton import StencilJacobian

        engine = self._engine(state)
        if hasattr(engine, "jacobian"):  # general engine: diagonals read off the generated Jacobian rows
            arrays = self.domain.arrays_from_state(state)
            return engine.operator_values(arrays), engine.operator_grad(arrays), engine.names
        values = engine.operator_values(self.domain.arrays_from_state(state))
        jac = StencilJacobian(engine)
        grads = []
        for d in jac.diagonals():
            loc = "c" * self.domain.ndim
            grads.append({(key, shift, loc): Known(torch.as_tensor(coef, dtype=engine.tdtype, device=engine.device))
                          for (key, shift), coef in d.items()})
        return values, grads, engine.names

    def eval_loss_grad(self, state):
        """
        Returns (loss, grads, terms, names, norms) like the reference (core.py:1219-1241).  `grads`
        are device tensors aligned with domain.arrays_from_state(state); loss / terms / norms are
        device-resident scalars that convert with np.array()/float() on first use, so a training
        loop that does not look at them never synchronises.
        """
        if not state.initialized:
            raise RuntimeError("Uninitialized state, use `state = domain.init_state(state)`")
        return self._eval_loss_grad(state)

    def eval_operator(self, state):
        if not state.initialized:
            raise RuntimeError("Uninitialized state, use `state = domain.init_state(state)`")
        return self._eval_operator(state)

    def eval_operator_grad(self, state):
        if not state.initialized:
            raise RuntimeError("Uninitialized state, use `state = domain.init_state(state)`")
        return self._eval_operator_grad(state)

    def linearize(self, state, modsp=None):
        """
        Returns (vector, matrix) with operator(packed + d) = vector + matrix.dot(d) (core.py:1113-1217).
        `vector` is a flat device tensor; `matrix` is a matrix-free `newton.StencilJacobian` (products on the
        device; `.tocsr()` gives the SciPy matrix the reference assembles).  Affine operators, multigrid off.
        """
        if not state.initialized:
            raise RuntimeError("Uninitialized state, use `state = domain.init_state(state)`")
        from .newton import StencilJacobian, residual_vector

        engine = self._engine(state)
        arrays = self.domain.arrays_from_state(state)
        vector = Known(residual_vector(engine, arrays))  # np.array(vector) and `-vector` both work (core.py:1127-1138)
        if hasattr(engine, "jacobian"):  # general (graph) engine: generated forward- / reverse-mode kernels
            return vector, engine.jacobian(arrays)
        return vector, StencilJacobian(engine)


# --------------------------------------------------------------------------------------------------
# Checkpoints (core.py:1389-1436): byte-compatible pickle {fields: {key: [ndarray, ...]}}
# --------------------------------------------------------------------------------------------------
def _to_numpy(a):
    if torch.is_tensor(a):
        return a.detach().cpu().numpy()
    return np.array(a)


def _global_numpy(domain, a):
    """Host copy of one state array in the reference's (global, halo-free) layout; collective in slab runs."""
    slab = getattr(domain, "slab", None)
    if slab is not None and torch.is_tensor(a) and a.dim() >= 1 and a.dim() == domain.ndim:
        return slab.gather(a).detach().cpu().numpy()
    return _to_numpy(a)


def checkpoint_save(domain, state, path):
    """Writes {fields: {key: [ndarray, ...]}} with GLOBAL arrays.  In slab-decomposed runs every rank takes part
    in the gather of the owned planes and rank 0 alone writes the file."""
    fields = {key: [_global_numpy(domain, a) for a in domain.arrays_from_field(f)] for key, f in state.fields.items()}
    slab = getattr(domain, "slab", None)
    if slab is not None and slab.rank != 0:
        return
    with open(path, "wb") as f:
        pickle.dump({"fields": fields}, f)


def checkpoint_load(domain, state, path, skip_missing=True, keys=None):
    with open(path, "rb") as f:
        data = pickle.load(f).get("fields", dict())
    slab = getattr(domain, "slab", None)
    for key in keys or state.fields.keys():
        if key not in data:
            if not skip_missing:
                raise RuntimeError(f"Field {key} not found in {path}")
            continue
        arrays = data[key]
        if not isinstance(arrays, list):
            arrays = [arrays]
        field = state.fields[key]
        if state.initialized:
            current = domain.arrays_from_field(field)
            assert_equal(len(arrays), len(current), f" arrays of field '{key}' in {path}")
            loaded = []
            for a, cur in zip(arrays, current):
                t = domain.mod.variable(a, dtype=domain.dtype)
                if slab is not None and isinstance(field, (Field, MultigridField)):
                    t = slab.scatter(t)  # checkpoints hold global arrays; cut this rank's slab (halos filled)
                if cur is not None and hasattr(cur, "shape"):
                    assert_equal(tuple(t.shape), tuple(cur.shape), f" for field '{key}' in {path}")
                loaded.append(t)
            arrays = loaded
        domain.arrays_to_field(arrays, field)


# --------------------------------------------------------------------------------------------------
# Pointwise helpers usable inside operators (core.py:1439-1457) -- pure arithmetic, trace-safe
# --------------------------------------------------------------------------------------------------
def extrap_quadh(u0, u1, u1p):
    """Quadratic extrapolation from points 0, 1, 1.5 to point 2."""
    return (u0 - 6 * u1 + 8 * u1p) / 3


def extrap_quad(u0, u1, u2):
    """Quadratic extrapolation from points 0, 1, 2 to point 3."""
    return u0 - 3 * u1 + 3 * u2


def extrap_linear(u0, u1):
    """Linear extrapolation from points 0, 1 to point 2."""
    return 2 * u1 - u0


def latin_hypercube(ndim, size, dtype):
    cut = np.linspace(0, 1, size + 1, dtype=dtype)
    u = np.random.rand(size, ndim).astype(dtype)
    pts = u * (cut[1:] - cut[:-1])[:, None] + cut[:-1, None]
    res = np.zeros_like(pts)
    for j in range(ndim):
        res[:, j] = pts[np.random.permutation(size), j]
    return res


def struct_to_numpy(mod, d):
    if isinstance(d, dict):
        return {k: struct_to_numpy(mod, v) for k, v in d.items()}
    if isinstance(d, (list, tuple)):
        return type(d)(struct_to_numpy(mod, v) for v in d)
    if torch.is_tensor(d) or isinstance(d, Lazy):
        return _to_numpy(d) if torch.is_tensor(d) else np.array(d)
    return d
