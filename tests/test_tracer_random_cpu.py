"""
Randomised check of the tracer and the lowering (odil_b200/backend.py, engine.py): operators are generated from
the calls user scripts make -- ctx.field with shifts, + - and scalar factors, mod.where on index masks (boundary
rows), mod.roll of expressions, constant fields, stop_gradient, two unknown fields -- traced once, lowered to
region-typed plans, and the plans applied by the oracle must reproduce the SAME operator function evaluated
directly with NumPy arrays (a NumPy `ctx` below: ctx.field = np.roll(U, -shift), core.py:963).  No GPU involved.
"""
import argparse

import numpy as np
import pytest

import odil
from odil_b200.engine import ResidualEngine
from oracle import odil_oracle as orc


class NumpyMod:
    """The handful of `mod` calls the generated operators use, on plain arrays."""
    where = staticmethod(np.where)

    @staticmethod
    def roll(a, shift, axis):
        return np.roll(a, shift, axis)

    @staticmethod
    def stop_gradient(a):
        return a

    @staticmethod
    def cast(v, dtype=None):
        return np.asarray(v, dtype=dtype)


class NumpyCtx:
    def __init__(self, shape, fields, extra):
        self.shape, self.fields, self.extra, self.mod = shape, fields, extra, NumpyMod()

    def field(self, key, *shift, frozen=False):
        shift = shift or (0,) * len(self.shape)
        return np.roll(self.fields[key], tuple(-s for s in shift), tuple(range(len(self.shape))))

    def indices(self):
        return tuple(np.meshgrid(*[np.arange(n) for n in self.shape], indexing="ij"))  # a tuple also in 1-D

    def size(self):
        return list(self.shape)


def make_operator(rng, ndim, keys, nconst):
    """Returns operator(ctx) built from a random expression tree (fixed by the draws made here)."""

    def draw_shift():
        return tuple(int(s) for s in rng.integers(-2, 3, size=ndim))

    def draw_mask():
        axis = int(rng.integers(ndim))
        kind = int(rng.integers(4))
        return axis, kind, int(rng.integers(0, 2))

    def build(depth):
        roll = rng.random()
        if depth == 0 or roll < 0.25:
            leaf = rng.random()
            if leaf < 0.75:
                return ("field", keys[int(rng.integers(len(keys)))], draw_shift(), rng.random() < 0.15)
            return ("const", int(rng.integers(nconst)))
        if roll < 0.55:
            return ("lin", float(rng.normal()), build(depth - 1), float(rng.normal()), build(depth - 1))
        if roll < 0.8:
            return ("where", draw_mask(), build(depth - 1), build(depth - 1))
        if roll < 0.9:
            return ("roll", int(rng.integers(-1, 2)), int(rng.integers(ndim)), build(depth - 1))
        return ("scale", float(rng.normal()), build(depth - 1))

    trees = [build(3) for _ in range(int(rng.integers(1, 3)))]

    def evaluate(ctx, node):
        mod = ctx.mod
        tag = node[0]
        if tag == "field":
            _, key, shift, frozen = node
            f = ctx.field(key, *shift)
            return mod.stop_gradient(f) if frozen else f
        if tag == "const":
            return ctx.extra.consts[node[1]]
        if tag == "lin":
            _, a, x, b, y = node
            return evaluate(ctx, x) * a - evaluate(ctx, y) * b
        if tag == "scale":
            return evaluate(ctx, node[2]) * node[1] / 3
        if tag == "roll":
            return mod.roll(evaluate(ctx, node[3]), node[1], node[2])
        _, (axis, kind, k), x, y = node
        idx, n = ctx.indices(), ctx.size()
        i, m = idx[axis], n[axis]
        cond = [i == k, i == m - 1 - k, i < k + 1, i >= m - 1 - k][kind]
        return mod.where(cond, evaluate(ctx, x), evaluate(ctx, y))

    def operator(ctx):
        return [("f%d" % t, evaluate(ctx, tree)) for t, tree in enumerate(trees)]

    return operator


def plan_values(engine, fields):
    """Outputs of the lowered plans: sum over blocks of the oracle's stencil applied to each unknown + constant."""
    res = []
    for out in engine.outputs:
        F = np.zeros(out.shape) if out.const is None else out.const.cpu().numpy().astype(np.float64).copy()
        for blk in out.blocks:
            spec = blk.spec
            tshape = tuple(2 * r + 1 for r in spec["rwidth"]) + (len(spec["offsets"]),)
            F = F + orc.stencil_forward(fields[blk.key], [tuple(int(v) for v in o) for o in spec["offsets"]],
                                        np.asarray(spec["table"], dtype=np.float64).reshape(tshape),
                                        spec["rwidth"], None)
        res.append(F)
    return res


@pytest.mark.parametrize("seed", range(60))
def test_random_affine_operator_lowers_to_equivalent_plans(seed):
    rng = np.random.default_rng(1000 + seed)
    ndim = int(rng.integers(1, 4))
    shape = tuple(int(n) for n in rng.integers(9, 13, size=ndim))
    keys = ["u", "v"][: int(rng.integers(1, 3))]
    domain = odil.Domain(cshape=shape, dimnames=["x", "y", "z"][:ndim], dtype=np.float64, multigrid=False)
    extra = argparse.Namespace(consts=[rng.standard_normal(shape) for _ in range(2)])
    operator = make_operator(rng, ndim, keys, len(extra.consts))
    state = odil.State()
    for key in keys:
        state.fields[key] = np.zeros(shape)
    state = domain.init_state(state)
    fields = {key: rng.standard_normal(shape) for key in keys}
    direct = [np.broadcast_to(v, shape) for _, v in operator(NumpyCtx(shape, fields, extra))]
    engine = ResidualEngine(odil.Problem(operator, domain, extra), state, trace_only=True)
    lowered = plan_values(engine, fields)
    assert engine.names == ["f%d" % t for t in range(len(direct))]
    for a, b in zip(lowered, direct):
        assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(b)))
