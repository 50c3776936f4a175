"""
Expression graph of operators that are NOT affine stencils (reference op seam backend.py:12-279).

The affine tracer (backend.Affine) keeps `sum coef * U[x + s] + const` in closed form and lowers it to the hand-written
stencil kernels.  Everything else -- products of fields, exp / sigmoid / tanh of fields, `where` on field values,
neural-network unknowns, location changes (pad / trim), slicing and concatenation of fields, per-cell coefficient
arrays, `Context.Raw` terms, epoch-dependent weights -- is traced into the small DAG below, exactly once, like the
reference's `jax.jit` trace (core.py:1106-1107).  `odil_b200.codegen` then emits ONE CUDA kernel per output shape
that evaluates the residuals, the loss partial sums and the reverse-mode adjoint per cell, compiled with NVRTC for
sm_100a (csrc/jit.cu).

Node kinds (`Expr.op`):
  input      an unknown array: a state array (Field / Array / NeuralNet weight) or the synthesised regular field of a
             MultigridField; attrs: slot
  const      a Known array (independent of the unknowns), kept in compact broadcast form; attrs: known
  lit        a scalar literal (with a broadcastable shape); attrs: value
  param      a run-time scalar (ctx.tracers[...]); attrs: name
  <ufunc>    neg exp log sin cos tanh sigmoid relu abs sqrt square floor cast_f logical_not
  <binary>   add sub mul div pow minimum maximum eq ne lt le gt ge logical_and logical_or
  where      where(cond, a, b)
  roll       out[i] = x[(i - shift) mod n] per axis; attrs: shifts (one per axis)
  index      basic indexing: attrs: spec = per input axis ('int', k) | ('slice', start, step, outaxis), out shape
  pad        zero padding; attrs: widths
  concat / stack   attrs: axis
  reshape / transpose (attrs: perm) / broadcast
  stopgrad   identity whose adjoint is dropped (mod.stop_gradient, ctx.field(..., frozen=True))
"""
import math

import numpy as np
import torch

from . import backend
from .backend import Known, Lazy, as_known

UNARY = {"neg", "exp", "log", "sin", "cos", "tanh", "sigmoid", "relu", "abs", "sqrt", "square", "floor", "cast_f",
         "logical_not", "stopgrad"}
BINARY = {"add", "sub", "mul", "div", "pow", "minimum", "maximum"}
COMPARE = {"eq", "ne", "lt", "le", "gt", "ge"}
LOGICAL = {"logical_and", "logical_or"}


class GraphError(NotImplementedError):
    pass


class Expr(Lazy):
    """One node of the traced expression graph.  Immutable; `kind` is 'f' (floating, the domain dtype), 'b' (bool)
    or 'i' (integer constant data)."""
    _count = 0

    def __init__(self, op, args=(), shape=(), dtype=None, kind="f", **attrs):
        self.op = op
        self.args = tuple(args)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype) if dtype is not None else np.dtype(np.float64)
        self.kind = kind
        self.attrs = attrs
        self.id = Expr._count
        Expr._count += 1

    # -- numpy-like surface ---------------------------------------------------------------------------
    def __getitem__(self, idx):
        return g_index(self, idx)

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        for i in range(self.shape[0]):
            yield self[i]

    def __abs__(self):
        return g_unary("abs", self)

    def __bool__(self):
        raise GraphError("the truth value of a traced expression is not known at trace time (use mod.where)")

    def __array__(self, dtype=None, copy=None):
        raise GraphError("a traced expression of the unknowns has no concrete value at trace time")

    def __and__(self, o):
        return g_binary("logical_and", self, o)

    __rand__ = __and__

    def __or__(self, o):
        return g_binary("logical_or", self, o)

    __ror__ = __or__

    def __invert__(self):
        return g_unary("logical_not", self)

    def astype(self, dtype):
        return g_cast(self, dtype)

    def flatten(self):
        return g_reshape(self, (-1,))

    def reshape(self, *shape):
        return g_reshape(self, shape[0] if len(shape) == 1 and np.ndim(shape[0]) else shape)

    @property
    def size(self):
        return math.prod(self.shape)

    def __repr__(self):
        return f"Expr#{self.id}({self.op}, shape={self.shape}, kind={self.kind})"


def is_expr(x):
    return isinstance(x, Expr)


def any_expr(*xs):
    return any(isinstance(x, Expr) for x in xs)


# --------------------------------------------------------------------------------------------------
# Leaves
# --------------------------------------------------------------------------------------------------
def g_input(slot, shape, dtype):
    return Expr("input", (), shape, dtype, "f", slot=slot)


def g_param(name, dtype):
    return Expr("param", (), (), dtype, "f", name=name)


def _graph_dtype(*xs):
    for x in xs:
        if isinstance(x, Expr) and x.kind == "f":
            return x.dtype
    for x in xs:
        if isinstance(x, Expr):
            return x.dtype
    return np.dtype(np.float64)


def node(x, dtype):
    """Expr for any operand the operator may combine with a traced value."""
    if isinstance(x, Expr):
        return x
    if isinstance(x, backend.Affine):
        raise GraphError("internal: Affine symbol met a graph expression (mixed tracing modes)")
    if isinstance(x, (bool, np.bool_)):
        return Expr("lit", (), (), np.bool_, "b", value=bool(x))
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        return Expr("lit", (), (), dtype, "f", value=int(x))
    if isinstance(x, (float, np.floating)):
        # NumPy scalars carry their own precision (a float32 step size stays a float32 value); Python floats are weak
        v = float(x)
        return Expr("lit", (), (), dtype, "f", value=v)
    k = as_known(x)
    kd = k.t.dtype
    kind = "b" if kd == torch.bool else ("f" if kd.is_floating_point else "i")
    if k.t.numel() == 1:
        v = k.t.reshape(-1)[0].item()
        return Expr("lit", (), k.shape, np.bool_ if kind == "b" else dtype, "b" if kind == "b" else "f", value=v)
    return Expr("const", (), k.shape, np.bool_ if kind == "b" else dtype, kind, known=k)


def _bshape(*shapes):
    return tuple(np.broadcast_shapes(*shapes))


# --------------------------------------------------------------------------------------------------
# Elementwise
# --------------------------------------------------------------------------------------------------
def g_unary(op, x):
    x = node(x, _graph_dtype(x))
    if op == "logical_not":
        return Expr(op, (x,), x.shape, np.bool_, "b")
    return Expr(op, (x,), x.shape, x.dtype if x.kind == "f" else _graph_dtype(x), "f")


def g_binary(op, a, b):
    dt = _graph_dtype(a, b)
    a, b = node(a, dt), node(b, dt)
    shape = _bshape(a.shape, b.shape)
    if op in COMPARE or op in LOGICAL:
        return Expr(op, (a, b), shape, np.bool_, "b")
    if op not in BINARY:
        raise GraphError(f"operation '{op}' is not supported on traced expressions")
    return Expr(op, (a, b), shape, dt, "f")


def g_where(c, a, b):
    dt = _graph_dtype(a, b, c)
    c, a, b = node(c, dt), node(a, dt), node(b, dt)
    return Expr("where", (c, a, b), _bshape(c.shape, a.shape, b.shape), dt, "f")


def g_cast(x, dtype):
    x = node(x, _graph_dtype(x))
    nd = backend.numpy_dtype(dtype)
    if np.issubdtype(nd, np.floating):
        # one floating dtype per graph (the domain dtype): a cast to another float width is the identity here
        return x if x.kind == "f" else Expr("cast_f", (x,), x.shape, x.dtype, "f")
    raise GraphError(f"cast of a traced expression to {nd}")


def g_stop_gradient(x):
    if not isinstance(x, Expr):
        return x
    return Expr("stopgrad", (x,), x.shape, x.dtype, x.kind)


# --------------------------------------------------------------------------------------------------
# Structural
# --------------------------------------------------------------------------------------------------
def g_roll(x, shifts, axes):
    x = node(x, _graph_dtype(x))
    nd = len(x.shape)
    per = [0] * nd
    for s, a in zip(shifts, axes):
        per[int(a) % nd] += int(s)
    per = [s % n if n > 0 else 0 for s, n in zip(per, x.shape)]
    if not any(per):
        return x
    if x.op == "roll":  # roll of a roll composes
        inner = x.attrs["shifts"]
        per = [(p + q) % n if n > 0 else 0 for p, q, n in zip(per, inner, x.shape)]
        x = x.args[0]
        if not any(per):
            return x
    return Expr("roll", (x,), x.shape, x.dtype, x.kind, shifts=tuple(per))


def g_index(x, idx):
    """Basic indexing (ints, slices, None, Ellipsis) of a traced expression."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    if any(isinstance(i, (list, np.ndarray, torch.Tensor, Known, Expr)) for i in idx):
        raise GraphError("advanced (array) indexing of a traced expression is not supported")
    nd = len(x.shape)
    n_given = sum(1 for i in idx if i is not None and i is not Ellipsis)
    if Ellipsis in idx:
        k = idx.index(Ellipsis)
        idx = idx[:k] + (slice(None),) * (nd - n_given) + idx[k + 1:]
    else:
        idx = idx + (slice(None),) * (nd - n_given)
    spec, out_shape = [], []
    axis = 0
    for i in idx:
        if i is None:
            out_shape.append(1)
            continue
        n = x.shape[axis]
        if isinstance(i, (int, np.integer)):
            k = int(i)
            if k < 0:
                k += n
            if not 0 <= k < n:
                raise IndexError(f"index {i} is out of bounds for axis {axis} with size {n}")
            spec.append(("int", k))
        elif isinstance(i, slice):
            start, stop, step = i.indices(n)
            count = len(range(start, stop, step))
            spec.append(("slice", start, step, len(out_shape)))
            out_shape.append(count)
        else:
            raise GraphError(f"unsupported index {i!r} on a traced expression")
        axis += 1
    if axis != nd:
        raise IndexError("too many indices for a traced expression")
    if tuple(out_shape) == x.shape and all(s[0] == "slice" and s[1] == 0 and s[2] == 1 and s[3] == a
                                           for a, s in enumerate(spec)):
        return x
    return Expr("index", (x,), out_shape, x.dtype, x.kind, spec=tuple(spec))


def g_pad(x, pad_width, mode="constant"):
    if mode != "constant":
        raise GraphError(f"pad mode '{mode}' of a traced expression")
    x = node(x, _graph_dtype(x))
    widths = tuple((int(lo), int(hi)) for lo, hi in pad_width)
    if len(widths) != len(x.shape):
        raise ValueError("pad_width must have one (before, after) pair per axis")
    if not any(lo or hi for lo, hi in widths):
        return x
    shape = tuple(n + lo + hi for n, (lo, hi) in zip(x.shape, widths))
    return Expr("pad", (x,), shape, x.dtype, x.kind, widths=widths)


def g_concat(xs, axis=0):
    dt = _graph_dtype(*xs)
    xs = [node(x, dt) for x in xs]
    nd = len(xs[0].shape)
    axis = int(axis) % nd
    for x in xs:
        if len(x.shape) != nd or any(x.shape[a] != xs[0].shape[a] for a in range(nd) if a != axis):
            raise ValueError(f"concatenate: incompatible shapes {[x.shape for x in xs]}")
    shape = list(xs[0].shape)
    shape[axis] = sum(x.shape[axis] for x in xs)
    if len(xs) == 1:
        return xs[0]
    return Expr("concat", xs, shape, dt, "f", axis=axis)


def g_stack(xs, axis=0):
    dt = _graph_dtype(*xs)
    xs = [node(x, dt) for x in xs]
    shape = _bshape(*[x.shape for x in xs])
    xs = [g_broadcast(x, shape) for x in xs]
    nd = len(shape) + 1
    axis = int(axis) % nd
    parts = [g_index(x, (slice(None),) * axis + (None,)) for x in xs]
    return g_concat(parts, axis)


def g_broadcast(x, shape):
    x = node(x, _graph_dtype(x))
    shape = _bshape(x.shape, tuple(shape))
    if shape == x.shape:
        return x
    return Expr("broadcast", (x,), shape, x.dtype, x.kind)


def g_reshape(x, shape):
    x = node(x, _graph_dtype(x))
    shape = [int(s) for s in (shape if np.ndim(shape) else [shape])]
    n = math.prod(x.shape)
    if -1 in shape:
        known = -math.prod(shape)
        shape[shape.index(-1)] = n // known if known else 0
    if math.prod(shape) != n:
        raise ValueError(f"cannot reshape {x.shape} into {tuple(shape)}")
    if tuple(shape) == x.shape:
        return x
    return Expr("reshape", (x,), shape, x.dtype, x.kind)


def g_transpose(x, perm=None):
    x = node(x, _graph_dtype(x))
    nd = len(x.shape)
    perm = tuple(int(p) % nd for p in perm) if perm is not None else tuple(reversed(range(nd)))
    if perm == tuple(range(nd)):
        return x
    return Expr("transpose", (x,), tuple(x.shape[p] for p in perm), x.dtype, x.kind, perm=perm)


def g_restrict(x, loc):
    """restrict_to_coarser of a traced field (core.py:703-755): loc 'c' = mean of the two children per axis, loc 'n'
    = [1,2,1]/4 at stride 2 on the array padded by linear extrapolation, '.' = unchanged -- written with slices."""
    x = node(x, _graph_dtype(x))
    for a, l in enumerate(loc):
        n = x.shape[a]
        sl = lambda s: (slice(None),) * a + (s,)
        if l == "c":
            x = (g_index(x, sl(slice(0, None, 2))) + g_index(x, sl(slice(1, None, 2)))) * 0.5
        elif l == "n":
            first = g_index(x, sl(slice(0, 1)))
            second = g_index(x, sl(slice(1, 2)))
            last = g_index(x, sl(slice(n - 1, n)))
            prev = g_index(x, sl(slice(n - 2, n - 1)))
            p = g_concat([2 * first - second, x, 2 * last - prev], axis=a)
            x = (g_index(p, sl(slice(0, n - 1, 2))) + 2 * g_index(p, sl(slice(1, n, 2)))
                 + g_index(p, sl(slice(2, n + 1, 2)))) * 0.25
        elif l != ".":
            raise ValueError("Invalid loc=" + loc)
    return x


def g_mlp(weights, biases, inputs, activation):
    """Fully connected network applied per element (core.py:807-862): inputs are arrays of one common shape, the
    outputs have the same shape.  Weights w[i] have shape (n_out, n_in); written with scalar-weight products so that
    the code generator sees plain elementwise arithmetic (weight loads are uniform: their adjoints are block-reduced)."""
    dt = _graph_dtype(*inputs, *weights, *biases)
    act = {"tanh": lambda v: g_unary("tanh", v), "relu": lambda v: g_unary("relu", v), "none": lambda v: v}[activation]
    h = [node(x, dt) for x in inputs]
    for li, (w, b) in enumerate(zip(weights, biases)):
        w, b = node(w, dt), node(b, dt)
        no, ni = w.shape
        assert ni == len(h), f"layer {li}: {ni} inputs expected, got {len(h)}"
        nh = []
        for j in range(no):
            acc = None
            for i in range(ni):
                t = g_binary("mul", g_index(w, (j, i)), h[i])
                acc = t if acc is None else g_binary("add", acc, t)
            acc = g_binary("add", acc, g_index(b, (j,)))
            nh.append(act(acc) if li < len(weights) - 1 else acc)
        h = nh
    return h


def toposort(outputs):
    """Nodes reachable from `outputs`, children before parents."""
    order, seen = [], set()
    stack = [(o, False) for o in reversed(outputs)]
    while stack:
        n, done = stack.pop()
        if done:
            order.append(n)
            continue
        if n.id in seen:
            continue
        seen.add(n.id)
        stack.append((n, True))
        for a in reversed(n.args):
            if a.id not in seen:
                stack.append((a, False))
    return order
