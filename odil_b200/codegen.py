"""
CUDA code generation for traced expression graphs (odil_b200.graph) -- the stand-in for XLA in the reference's
`jax.jit(value_and_grad(eval_loss))` (core.py:1100-1107) for operators that are not affine stencils.

One *group* = the operator outputs of one shape.  For every group and mode one kernel is generated; a thread walks
cells in a grid-stride loop and, per cell, executes a straight-line SSA program:

  values    F_k[cell]                                         (Problem.eval_operator, core.py:1298-1311)
  lossgrad  sum F_k^2 partials (fp64) + reverse-mode adjoint with seed 2 w_k F_k: every load of an unknown scatters its
            adjoint with atomicAdd (the transposed stencil); loads whose address does not depend on the cell (neural-net
            weights, elements of Array unknowns) accumulate in registers and are block-reduced once per block
  jvp       forward-mode tangent  (J v)_k[cell]               (Newton: matrix-free Jacobian products)
  vjp       reverse mode with a given cotangent  J^T w
  jac       per cell and per load: column index and dF/d(load) (COO rows of the Jacobian, core.py:1144-1171)
  jacd      per cell: dF_k/d(load) of every structurally connected (output, load) pair, one array per pair -- the
            reference's per-(key, shift, loc) Jacobian "diagonals" (core.py:1341-1350) kept on the device
  jvpd/vjpd the Jacobian products of a Newton step's CG iterations from those stored diagonals: J v = sum_l D_l * v[col_l],
            J^T w scattered likewise -- no re-evaluation of the operator's arithmetic (exp, divisions) per product
  vjpg      J^T w GATHERED per input cell (no atomics) when every connected load is a pure roll: ctx.field(key, *shift)
            without a location change -- the common case

Index arithmetic (roll, slicing, pad, concatenate, reshape, transpose, broadcasting) is resolved symbolically per cell:
`emit(node, index tuple, guard)` returns the SSA value of `node` at that index; integer literals are folded at
generation time so a scalar output such as u[-1, ixc] costs no index arithmetic at all.

The same source compiles for the host when ODIL_HOST is defined (plain loops, `+=` for atomicAdd): tests use that to
check the generated programs against reference-generated goldens without a GPU.  The product path never does.
"""
import math
import os
import struct

import numpy as np
import torch

from .graph import BINARY, COMPARE, LOGICAL, GraphError, toposort

BLOCK = 256
MAX_GRID = 148 * 4
EXACT_DIV = os.environ.get("ODIL_B200_EXACT_DIV", "0") not in ("", "0")


def _lit(v, real):
    if isinstance(v, bool):
        return "true" if v else "false"
    v = float(v)
    if math.isnan(v):
        return "T(NAN)"
    if math.isinf(v):
        return "T(INFINITY)" if v > 0 else "T(-INFINITY)"
    if v == int(v) and abs(v) < 2 ** 31:
        return f"T({int(v)})"
    return f"T({v!r})"


class Val:
    __slots__ = ("name", "kind", "active")

    def __init__(self, name, kind, active=False):
        self.name, self.kind, self.active = name, kind, active


class GroupProgram:
    """Straight-line per-cell program of one output group."""

    def __init__(self, gen, gid, shape, outputs):
        self.gen, self.gid, self.shape = gen, gid, tuple(shape)
        self.outputs = outputs  # list of (global output index, Expr, is_raw)
        self.lines = []         # forward statements
        self.tape = []          # ('load', z, slot, lin, guard, uniform) | ('op', z, [(arg, contrib, tcontrib)])
        self.memo = {}
        self.defs = {}          # expression string -> SSA name (ints / bools / loads)
        self.fdefs = {}         # floating expression string -> SSA name (value numbering)
        self.tape_index = {}    # SSA name -> its ('op', ...) tape entry
        self.n = 0
        self.uniform = {}       # (slot, lin) -> index of the register accumulator
        self.shift = {}         # index variable -> (axis, k): its value is (c_axis + k) mod shape[axis]
        self.load_shift = {}    # load variable -> per-axis k if the load reads cell (c + k) mod shape of an array of the
                                # group's shape (a pure roll), else None
        self.ncell = math.prod(self.shape)
        nd = len(self.shape)
        self.I0 = tuple(f"c{a}" if self.shape[a] > 1 else 0 for a in range(nd))
        self.results = []
        for k, e, raw in outputs:
            v = self.emit(e, self.bidx(e.shape, self.I0, self.shape), None)
            self.results.append((k, v, raw))

    # -- SSA helpers ----------------------------------------------------------------------------------
    def new(self, prefix):
        self.n += 1
        return f"{prefix}{self.n}"

    def define(self, ctype, expr, prefix):
        key = (ctype, expr)
        if key not in self.defs:
            name = self.new(prefix)
            self.lines.append(f"const {ctype} {name} = {expr};")
            self.defs[key] = name
        return self.defs[key]

    def ivar(self, expr):
        return self.define("int", expr, "i")

    def bvar(self, expr):
        return self.define("bool", expr, "b")

    def iadd(self, i, c):
        if isinstance(i, int):
            return i + c
        if c == 0:
            return i
        return self.ivar(f"{i} + {c}" if c > 0 else f"{i} - {-c}")

    def imul_add(self, i, m, c):
        """m * i + c"""
        if isinstance(i, int):
            return m * i + c
        if m == 1:
            return self.iadd(i, c)
        return self.ivar(f"{m} * {i}" + (f" + {c}" if c > 0 else (f" - {-c}" if c < 0 else "")))

    def band(self, *gs):
        gs = [g for g in gs if g is not None and g is not True]
        if any(g is False for g in gs):
            return False
        gs = sorted(set(gs))
        if not gs:
            return None
        if len(gs) == 1:
            return gs[0]
        return self.bvar(" && ".join(gs))

    def in_range(self, i, lo, hi):
        """lo <= i < hi as a guard (None = always true, False = never)."""
        if isinstance(i, int):
            return None if lo <= i < hi else False
        return self.bvar(f"{i} >= {lo} && {i} < {hi}")

    def bidx(self, child_shape, I, shape):
        off = len(shape) - len(child_shape)
        return tuple(0 if child_shape[a] == 1 else I[a + off] for a in range(len(child_shape)))

    def lin(self, I, strides, numel):
        const, terms = 0, []
        for i, s in zip(I, strides):
            if s == 0:
                continue
            if isinstance(i, int):
                const += i * s
            else:
                terms.append(i if s == 1 else f"{i} * {s}")
        if not terms:
            return const
        big = numel >= 2 ** 31
        if big:
            terms = [f"(long long){t}" for t in terms]
        expr = " + ".join(terms) + (f" + {const}" if const else "")
        return self.define("long long" if big else "int", expr, "l")

    # -- node emission --------------------------------------------------------------------------------
    def emit(self, n, I, G):
        if G is False:
            return Val("false" if n.kind == "b" else "T(0)", n.kind, False)
        key = (n.id, I, G)
        v = self.memo.get(key)
        if v is None:
            v = self._emit(n, I, G)
            self.memo[key] = v
        return v

    def fvar(self, expr):
        """SSA variable of a floating expression; identical expressions share one variable (value numbering: a frozen
        copy of a stencil -- ctx.field(..., frozen=True) -- re-uses the arithmetic of the live one, only the adjoint
        flow differs)."""
        name = self.fdefs.get(expr)
        if name is None:
            name = self.new("v")
            self.lines.append(f"const T {name} = {expr};")
            self.fdefs[expr] = name
        return name

    def record(self, z, parts):
        """Adjoint rules of z: parts = [(argument position, argument variable, rule)].  One tape entry per variable,
        created at its first ACTIVE use; later uses add the positions that were inactive before."""
        entry = self.tape_index.get(z)
        if entry is None:
            entry = ("op", z, [])
            self.tape.append(entry)
            self.tape_index[z] = entry
        have = {p[0] for p in entry[2]}
        for part in parts:
            if part[0] not in have:
                entry[2].append(part)

    def as_float(self, v):
        if v.kind == "b":
            return f"({v.name} ? T(1) : T(0))"
        return v.name

    def as_bool(self, v):
        return v.name if v.kind == "b" else f"({v.name} != T(0))"

    def _emit(self, n, I, G):
        op, gen = n.op, self.gen
        if op == "lit":
            return Val(_lit(n.attrs["value"], gen.real), n.kind, False)
        if op == "param":
            return Val(self.define("T", f"T(a.prm[{gen.param_index(n.attrs['name'])}])", "p"), "f", False)
        if op == "const":
            j, strides, numel = gen.const_slot(n)
            l = self.lin(I, strides, numel)
            guard = "" if G is None else f"!{G} ? {'false' if n.kind == 'b' else 'T(0)'} : "
            if n.kind == "b":
                return Val(self.define("bool", f"{guard}((const unsigned char*)a.kc[{j}])[{l}] != 0", "k"), "b", False)
            return Val(self.define("T", f"{guard}((const T*)a.kc[{j}])[{l}]", "k"), "f", False)
        if op == "input":
            slot = n.attrs["slot"]
            strides = gen.c_strides(n.shape)
            l = self.lin(I, strides, math.prod(n.shape))
            key = ("ld", slot, l, G)
            if key not in self.defs:
                name = self.new("u")
                src = f"a.in[{slot}][{l}]"
                self.lines.append(f"const T {name} = {src};" if G is None else f"const T {name} = {G} ? {src} : T(0);")
                self.defs[key] = name
                uniform = isinstance(l, int)
                self.tape.append(("load", name, slot, l, G, uniform))
                ks = None
                if G is None and tuple(n.shape) == self.shape:
                    ks = []
                    for a, i in enumerate(I):
                        if isinstance(i, int):
                            ks.append(0 if self.shape[a] == 1 and i == 0 else None)
                        elif i == f"c{a}":
                            ks.append(0)
                        else:
                            sh = self.shift.get(i)
                            ks.append(sh[1] if sh is not None and sh[0] == a else None)
                    ks = None if any(k is None for k in ks) else tuple(ks)
                self.load_shift[name] = ks
            return Val(self.defs[key], "f", True)
        if op == "stopgrad":
            x = self.emit(n.args[0], I, G)
            return x if not x.active else Val(x.name, x.kind, False)  # same value, adjoint flow cut
        if op in ("broadcast",):
            return self.emit(n.args[0], self.bidx(n.args[0].shape, I, n.shape), G)
        if op == "roll":
            x = n.args[0]
            J = []
            for i, s, m in zip(I, n.attrs["shifts"], x.shape):
                if s == 0 or m == 1:
                    J.append(i)
                elif isinstance(i, int):
                    J.append((i - s) % m)
                else:
                    v = self.ivar(f"{i} >= {s} ? {i} - {s} : {i} + {m - s}")
                    base = self.shift.get(i)
                    if base is None and i.startswith("c") and i[1:].isdigit():
                        base = (int(i[1:]), 0)
                    if base is not None and m == self.shape[base[0]]:
                        self.shift[v] = (base[0], (base[1] - s) % m)
                    J.append(v)
            return self.emit(x, tuple(J), G)
        if op == "index":
            J = []
            for sp in n.attrs["spec"]:
                if sp[0] == "int":
                    J.append(sp[1])
                else:
                    _, start, step, oa = sp
                    J.append(self.imul_add(I[oa], step, start))
            return self.emit(n.args[0], tuple(J), G)
        if op == "transpose":
            perm = n.attrs["perm"]
            J = [None] * len(perm)
            for a, p in enumerate(perm):
                J[p] = I[a]
            return self.emit(n.args[0], tuple(J), G)
        if op == "reshape":
            x = n.args[0]
            l = self.lin(I, gen.c_strides(n.shape), math.prod(n.shape))
            J = []
            rem = l
            for a, (m, s) in enumerate(zip(x.shape, gen.c_strides(x.shape))):
                if isinstance(rem, int):
                    J.append(rem // s if a == 0 else (rem // s) % m)
                elif m == 1:
                    J.append(0)
                elif a == 0:
                    J.append(self.ivar(f"(int)({rem} / {s})"))
                else:
                    J.append(self.ivar(f"(int)(({rem} / {s}) % {m})" if s != 1 else f"(int)({rem} % {m})"))
            return self.emit(x, tuple(J), G)
        if op == "pad":
            x = n.args[0]
            J, inr = [], []
            for i, (lo, hi), m in zip(I, n.attrs["widths"], x.shape):
                J.append(self.iadd(i, -lo))
                if lo or hi:
                    inr.append(self.in_range(i, lo, lo + m))
            inside = self.band(*inr)
            if inside is False:
                return Val("false" if n.kind == "b" else "T(0)", n.kind, False)
            v = self.emit(x, tuple(J), self.band(G, inside))
            if inside is None:
                return v
            return self.select(inside, v, Val("T(0)", "f", False))
        if op == "concat":
            ax = n.attrs["axis"]
            pieces, start = [], 0
            for x in n.args:
                m = x.shape[ax]
                inside = self.in_range(I[ax], start, start + m)
                J = tuple(self.iadd(i, -start) if a == ax else i for a, i in enumerate(I))
                if inside is not False:
                    pieces.append((inside, self.emit(x, J, self.band(G, inside))))
                start += m
            res = None
            for inside, v in reversed(pieces):
                if inside is None or res is None:
                    res = v
                else:
                    res = self.select(inside, v, res)
            return res
        if op == "where":
            c, x, y = n.args
            cv = self.emit(c, self.bidx(c.shape, I, n.shape), G)
            xv = self.emit(x, self.bidx(x.shape, I, n.shape), G)
            yv = self.emit(y, self.bidx(y.shape, I, n.shape), G)
            if cv.name in ("true", "false"):
                return xv if cv.name == "true" else yv
            return self.select(self.as_bool(cv), xv, yv)
        args = [self.emit(x, self.bidx(x.shape, I, n.shape), G) for x in n.args]
        if op in COMPARE:
            sym = {"eq": "==", "ne": "!=", "lt": "<", "le": "<=", "gt": ">", "ge": ">="}[op]
            return Val(self.bvar(f"{self.as_float(args[0])} {sym} {self.as_float(args[1])}"), "b", False)
        if op in LOGICAL:
            sym = "&&" if op == "logical_and" else "||"
            return Val(self.bvar(f"{self.as_bool(args[0])} {sym} {self.as_bool(args[1])}"), "b", False)
        if op == "logical_not":
            return Val(self.bvar(f"!{self.as_bool(args[0])}"), "b", False)
        return self.arith(op, n, args)

    def select(self, cond, x, y):
        xa, ya = self.as_float(x), self.as_float(y)
        if xa == ya:
            return Val(xa, "f", x.active or y.active)
        z = self.fvar(f"{cond} ? {xa} : {ya}")
        parts = []
        if x.active:
            parts.append((0, x.name, lambda d, c=cond: f"({c} ? {d} : T(0))"))
        if y.active:
            parts.append((1, y.name, lambda d, c=cond: f"({c} ? T(0) : {d})"))
        if parts:
            self.record(z, parts)
        return Val(z, "f", bool(parts))

    @staticmethod
    def _literal(text):
        """Value of a `T(...)` literal operand, else None."""
        if text.startswith("T(") and text.endswith(")"):
            try:
                return float(text[2:-1])
            except ValueError:
                return None
        return None

    def arith(self, op, n, args):
        a = self.as_float(args[0])
        b = self.as_float(args[1]) if len(args) > 1 else None
        lin = lambda p: (lambda d, p=p: f"{p} * {d}")
        one = lambda d: d
        neg = lambda d: f"-{d}"
        # exact algebraic identities (x + 0, x * 1, x / 1) and division by a literal whose reciprocal is exact (a
        # power of two): the quotient is bit-identical and costs one multiplication instead of a division sequence
        la, lb = self._literal(a), (self._literal(b) if b is not None else None)
        if op in ("add", "sub") and lb == 0.0:
            return args[0] if args[0].kind == "f" else Val(a, "f", False)
        if op == "add" and la == 0.0:
            return args[1] if args[1].kind == "f" else Val(b, "f", False)
        if op in ("mul", "div") and lb == 1.0:
            return args[0] if args[0].kind == "f" else Val(a, "f", False)
        if op == "mul" and la == 1.0:
            return args[1] if args[1].kind == "f" else Val(b, "f", False)
        if op == "div" and lb is not None and lb != 0.0 and math.isfinite(lb) and math.isfinite(1.0 / lb):
            mant, _ = math.frexp(lb)
            # Other literal divisors (the /3 of extrap_quadh, evaluated for every cell under a boundary `where`): the
            # product with the rounded reciprocal differs from the quotient by at most one ulp -- far inside the parity
            # bars -- and replaces a ~25-instruction fp64 division sequence; ODIL_B200_EXACT_DIV=1 keeps the division.
            if abs(mant) == 0.5 or not EXACT_DIV:
                op, b = "mul", _lit(1.0 / lb, self.gen.real)
        if op == "add":
            expr, parts = f"{a} + {b}", [one, one]
        elif op == "sub":
            expr, parts = f"{a} - {b}", [one, neg]
        elif op == "mul":
            expr, parts = f"{a} * {b}", [lin(b), lin(a)]
        elif op == "div":
            expr, parts = f"{a} / {b}", None
        elif op == "neg":
            expr, parts = f"-{a}", [neg]
        elif op == "cast_f":
            expr, parts = a, [one]
        elif op == "square":
            expr, parts = f"{a} * {a}", [lin(f"(T(2) * {a})")]
        elif op == "pow":
            e = n.args[1]
            if e.op == "lit" and float(e.attrs["value"]) == 2.0:
                expr, parts = f"{a} * {a}", [lin(f"(T(2) * {a})"), None]
            elif e.op == "lit" and float(e.attrs["value"]) == 1.0:
                expr, parts = a, [one, None]
            elif e.op == "lit" and float(e.attrs["value"]) == 3.0:
                expr, parts = f"{a} * {a} * {a}", [lin(f"(T(3) * {a} * {a})"), None]
            else:
                expr, parts = f"POW({a}, {b})", None
        elif op in ("exp", "log", "sin", "cos", "tanh", "sqrt", "floor", "abs"):
            expr, parts = f"{op.upper()}({a})", None
        elif op == "sigmoid":
            expr, parts = f"T(1) / (T(1) + EXP(-{a}))", None
        elif op == "relu":
            expr, parts = f"{a} > T(0) ? {a} : T(0)", [lambda d, a=a: f"({a} > T(0) ? {d} : T(0))"]
        elif op == "minimum":
            expr, parts = f"{a} < {b} ? {a} : {b}", [lambda d: f"({a} < {b} ? {d} : T(0))",
                                                      lambda d: f"({a} < {b} ? T(0) : {d})"]
        elif op == "maximum":
            expr, parts = f"{a} > {b} ? {a} : {b}", [lambda d: f"({a} > {b} ? {d} : T(0))",
                                                      lambda d: f"({a} > {b} ? T(0) : {d})"]
        else:
            raise GraphError(f"no code generator for operation '{op}'")
        z = self.fvar(expr)
        if parts is None:  # partials that need the result
            parts = {
                "div": lambda: [lin(f"(T(1) / {b})"), lin(f"(-{z} / {b})")],
                "pow": lambda: [lin(f"({b} * POW({a}, {b} - T(1)))"), lin(f"({z} * LOG({a}))")],
                "exp": lambda: [lin(z)],
                "log": lambda: [lin(f"(T(1) / {a})")],
                "sin": lambda: [lin(f"COS({a})")],
                "cos": lambda: [lin(f"(-SIN({a}))")],
                "tanh": lambda: [lin(f"(T(1) - {z} * {z})")],
                "sqrt": lambda: [lin(f"(T(0.5) / {z})")],
                "sigmoid": lambda: [lin(f"({z} * (T(1) - {z}))")],
                "floor": lambda: [None],
                "abs": lambda: [lambda d: f"({a} > T(0) ? {d} : ({a} < T(0) ? -{d} : T(0)))"],
            }[op]()
        active = [(i, x.name, p) for i, (x, p) in enumerate(zip(args, parts)) if x.active and p is not None]
        if active:
            self.record(z, active)
        return Val(z, "f", bool(active))

    # -- per-cell bodies ------------------------------------------------------------------------------
    def coords(self):
        out, rem = [], "cell"
        if self.ncell < 2 ** 31:  # 32-bit divisions by constants (multiply-shift) instead of 64-bit ones
            out.append("const unsigned cell32 = (unsigned)cell;")
            rem = "cell32"
        nd = len(self.shape)
        strides = self.gen.c_strides(self.shape)
        for a in range(nd):
            if self.shape[a] == 1:
                continue
            if strides[a] == 1:
                out.append(f"const int c{a} = (int)({rem} % {self.shape[a]});")
            elif all(self.shape[b] == 1 for b in range(a)):
                out.append(f"const int c{a} = (int)({rem} / {strides[a]});")
            else:
                out.append(f"const int c{a} = (int)(({rem} / {strides[a]}) % {self.shape[a]});")
        return out

    def uniform_index(self, slot, lin):
        key = (slot, lin)
        if key not in self.uniform:
            self.uniform[key] = len(self.uniform)
        return self.uniform[key]

    def backward(self, seeds, mode):
        """Reverse sweep over the tape.  seeds: {var: expression}.  mode 'grad': scatter adjoints of loads;
        mode 'jac': write (column, value) per load."""
        out = []
        adj = {}

        def add(var, expr):
            if var in adj:
                out.append(f"d_{var} += {expr};")
            else:
                out.append(f"T d_{var} = {expr};")
                adj[var] = True

        for var, expr in seeds.items():
            add(var, expr)
        loads = [e for e in self.tape if e[0] == "load"]
        for e in reversed(self.tape):
            z = e[1]
            if e[0] == "op":
                if z not in adj:
                    continue
                for _, arg, contrib in e[2]:
                    add(arg, contrib(f"d_{z}"))
            else:
                _, z, slot, lin, G, uniform = e
                if mode == "jac":
                    j = loads.index(e)
                    col = f"a.colbase[{slot}] + (long long){lin}"
                    val = f"d_{z}" if z in adj else "T(0)"
                    g = "" if G is None else f"!{G} ? T(0) : "
                    out.append(f"a.jcol[cell * {len(loads)} + {j}] = {col};")
                    out.append(f"a.jval[cell * {len(loads)} + {j}] = {g}{val};")
                    continue
                if z not in adj or not self.gen.slot_has_grad(slot):
                    continue
                if uniform:
                    out.append(f"ua[{self.uniform_index(slot, lin)}] += d_{z};")
                elif G is None:
                    out.append(f"ATOMIC_ADD(&a.gin[{slot}][{lin}], d_{z});")
                else:
                    out.append(f"if ({G}) ATOMIC_ADD(&a.gin[{slot}][{lin}], d_{z});")
        return out

    def loads(self):
        return [e for e in self.tape if e[0] == "load"]

    def pairs(self):
        """[(result position j, load position l)] whose derivative d result_j / d load_l is not structurally zero, for
        loads of slots that receive gradients; None if such a load has a cell-independent address (network weights,
        elements of Array unknowns: those take the plain jvp / vjp kernels)."""
        if not hasattr(self, "_pairs"):
            loads = self.loads()
            pairs = []
            for j, (k, v, raw) in enumerate(self.results):
                if not v.active:
                    continue
                reach = {v.name}
                for e in reversed(self.tape):
                    if e[0] == "op" and e[1] in reach:
                        reach.update(arg for _, arg, _c in e[2])
                for l, e in enumerate(loads):
                    if e[1] in reach and self.gen.slot_has_grad(e[2]):
                        if e[5]:
                            pairs = None
                            break
                        pairs.append((j, l))
                if pairs is None:
                    break
            self._pairs = pairs
        return self._pairs

    def diagonal_store(self):
        """Reverse sweep per result with seed 1; the adjoint of every connected load goes to its pair's array."""
        out = []
        loads = self.loads()
        pairs = self.pairs()
        for j, (k, v, raw) in enumerate(self.results):
            mine = {l: pi for pi, (jj, l) in enumerate(pairs) if jj == j}
            if not mine:
                continue
            out.append("{")
            adj = {}

            def add(var, expr):
                if var in adj:
                    out.append(f"d_{var} += {expr};")
                else:
                    out.append(f"T d_{var} = {expr};")
                    adj[var] = True

            add(v.name, "T(1)")
            for e in reversed(self.tape):
                z = e[1]
                if e[0] == "op":
                    if z in adj:
                        for _, arg, contrib in e[2]:
                            add(arg, contrib(f"d_{z}"))
                else:
                    l = loads.index(e)
                    if l in mine:
                        G = e[4]
                        val = f"d_{z}" if z in adj else "T(0)"
                        g = "" if G is None else f"!{G} ? T(0) : "
                        out.append(f"a.jval[{mine[l]}ll * a.ncell + cell] = {g}{val};")
            out.append("}")
        return out

    def diagonal_products(self, mode):
        out = []
        loads = self.loads()
        pairs = self.pairs()
        used = sorted({l for _, l in pairs})
        D = lambda pi: f"a.jval[{pi}ll * a.ncell + cell]"
        if mode == "jvpd":
            for l in used:
                _, z, slot, lin, G, _u = loads[l]
                src = f"a.tin[{slot}][{lin}]"
                out.append(f"const T t_{z} = {src};" if G is None else f"const T t_{z} = {G} ? {src} : T(0);")
            for j, (k, v, raw) in enumerate(self.results):
                terms = [f"{D(pi)} * t_{loads[l][1]}" for pi, (jj, l) in enumerate(pairs) if jj == j]
                out.append(f"a.out[{k}][cell] = {' + '.join(terms) if terms else 'T(0)'};")
        else:
            for j in sorted({j for j, _ in pairs}):
                out.append(f"const T s_{j} = a.seed[{self.results[j][0]}][cell];")
            for l in used:
                _, z, slot, lin, G, _u = loads[l]
                terms = [f"{D(pi)} * s_{jj}" for pi, (jj, ll) in enumerate(pairs) if ll == l]
                stmt = f"ATOMIC_ADD(&a.gin[{slot}][{lin}], {' + '.join(terms)});"
                out.append(stmt if G is None else f"if ({G}) {stmt}")
        return out

    def gather_ok(self):
        """Every connected load is a pure roll of an array of the group's shape: J^T w can be GATHERED per input cell
        (no atomics, no zero fill of the result beyond the caller's)."""
        pairs = self.pairs()
        if pairs is None:
            return False
        loads = self.loads()
        return all(self.load_shift.get(loads[l][1]) is not None for _, l in pairs)

    def diagonal_gather(self):
        """vjpg: g[slot][cell] += sum over loads l of that slot and results j of D[j,l][src] * seed_j[src], where src is
        the cell whose load l reads `cell`: src_a = (c_a - k_a) mod n_a."""
        out = []
        loads = self.loads()
        pairs = self.pairs()
        strides = self.gen.c_strides(self.shape)
        srcs = {}
        accs = {}
        for l in sorted({l for _, l in pairs}):
            _, z, slot, lin, G, _u = loads[l]
            ks = self.load_shift[z]
            if ks not in srcs:
                q = len(srcs)
                terms = []
                for a, k in enumerate(ks):
                    n = self.shape[a]
                    if n == 1:
                        continue
                    if k == 0:
                        v = f"c{a}"
                    else:
                        v = f"q{q}_{a}"
                        out.append(f"const int {v} = c{a} >= {k} ? c{a} - {k} : c{a} + {n - k};")
                    big = self.ncell >= 2 ** 31
                    terms.append(v if strides[a] == 1 else (f"(long long){v} * {strides[a]}" if big else f"{v} * {strides[a]}"))
                out.append(f"const {'long long' if self.ncell >= 2 ** 31 else 'int'} s{q} = "
                           f"{' + '.join(terms) if terms else '0'};")
                srcs[ks] = f"s{q}"
            src = srcs[ks]
            terms = [f"a.jval[{pi}ll * a.ncell + {src}] * a.seed[{self.results[jj][0]}][{src}]"
                     for pi, (jj, ll) in enumerate(pairs) if ll == l]
            if slot not in accs:
                accs[slot] = f"g{slot}"
                out.append(f"T g{slot} = {' + '.join(terms)};")
            else:
                out.append(f"g{slot} += {' + '.join(terms)};")
        for slot, name in accs.items():
            out.append(f"a.gin[{slot}][cell] += {name};")
        return out

    def forward_tangent(self):
        out, tan = [], {}
        for e in self.tape:
            z = e[1]
            if e[0] == "load":
                _, z, slot, lin, G, uniform = e
                if not self.gen.slot_has_grad(slot):
                    continue
                src = f"a.tin[{slot}][{lin}]"
                out.append(f"const T t_{z} = {src};" if G is None else f"const T t_{z} = {G} ? {src} : T(0);")
                tan[z] = True
            else:
                terms = [contrib(f"t_{arg}") for _, arg, contrib in e[2] if arg in tan]
                if terms:
                    out.append(f"const T t_{z} = {' + '.join(terms)};")
                    tan[z] = True
        return out, tan

    def body(self, mode, which=None):
        """Statements of one cell for `mode`; `which` selects one result of the group (mode 'jac')."""
        lines = list(self.coords()) + list(self.lines)
        nloads = len([e for e in self.tape if e[0] == "load"])
        if mode == "values":
            for k, v, raw in self.results:
                lines.append(f"a.out[{k}][cell] = {self.as_float(v)};")
        elif mode == "lossgrad":
            seeds = {}
            for j, (k, v, raw) in enumerate(self.results):
                f = self.as_float(v)
                w = f"T(a.prm[{self.gen.weight_index(k)}])"
                lines.append(f"lacc[{j}] += (double){f};" if raw else f"lacc[{j}] += (double){f} * (double){f};")
                if v.active:
                    seed = w if raw else f"T(2) * {w} * {f}"
                    seeds[v.name] = seed if v.name not in seeds else f"{seeds[v.name]} + {seed}"
            lines += self.backward(seeds, "grad")
        elif mode == "vjp":
            seeds = {}
            for k, v, raw in self.results:
                if v.active:
                    seed = f"a.seed[{k}][cell]"
                    seeds[v.name] = seed if v.name not in seeds else f"{seeds[v.name]} + {seed}"
            lines += self.backward(seeds, "grad")
        elif mode == "jvp":
            tl, tan = self.forward_tangent()
            lines += tl
            for k, v, raw in self.results:
                lines.append(f"a.out[{k}][cell] = {'t_' + v.name if v.name in tan else 'T(0)'};")
        elif mode == "jac":
            k, v, raw = self.results[which]
            lines += self.backward({v.name: "T(1)"} if v.active else {}, "jac")
        elif mode == "jacd":
            lines += self.diagonal_store()
        elif mode == "vjpg":
            return list(self.coords()) + self.diagonal_gather(), nloads
        elif mode in ("jvpd", "vjpd"):
            # only the index arithmetic of the forward statements is live here; the compiler drops the rest
            lines += self.diagonal_products(mode)
        else:
            raise ValueError(mode)
        return lines, nloads


PRELUDE = r"""
// Generated by odil_b200.codegen -- do not edit.  One translation unit per traced operator and mode.
#ifdef ODIL_HOST
#include <cmath>
#define ODIL_DEVICE static inline
#define ATOMIC_ADD(p, v) (*(p) += (v))
#else
#define ODIL_DEVICE __device__ __forceinline__
#define ATOMIC_ADD(p, v) atomicAdd((p), (v))
#ifndef NAN
#define NAN __int_as_float(0x7fffffff)
#define INFINITY __int_as_float(0x7f800000)
#endif
#endif
typedef REAL T;
#if REAL_IS_FLOAT
#define EXP(x) expf(x)
#define LOG(x) logf(x)
#define SIN(x) sinf(x)
#define COS(x) cosf(x)
#define TANH(x) tanhf(x)
#define SQRT(x) sqrtf(x)
#define FLOOR(x) floorf(x)
#define ABS(x) fabsf(x)
#define POW(x, y) powf(x, y)
#else
#define EXP(x) exp(x)
#define LOG(x) log(x)
#define SIN(x) sin(x)
#define COS(x) cos(x)
#define TANH(x) tanh(x)
#define SQRT(x) sqrt(x)
#define FLOOR(x) floor(x)
#define ABS(x) fabs(x)
#define POW(x, y) pow(x, y)
#endif
"""

BLOCK_SUM = r"""
#ifndef ODIL_HOST
__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    if (warp == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;  // valid in thread 0
}
#endif
"""

REDUCE = r"""
#ifndef ODIL_HOST
// sums[k] = sum of the per-block partials of output k, in a fixed order (deterministic loss)
extern "C" __global__ void __launch_bounds__(256) k_reduce(const P a) {
    __shared__ double red[32];
    const int k = blockIdx.x;
    double v = 0.0;
    for (int i = threadIdx.x; i < a.nblk; i += blockDim.x) v += a.partials[(long long)k * a.nblk + i];
    v = block_sum(v, red);
    if (threadIdx.x == 0) a.sums[k] = v;
}
#endif
"""


class Generator:
    """Source of all kernels of one traced operator."""

    def __init__(self, outputs, real, input_shapes, grad_slots, param_names):
        """outputs: list of (Expr, is_raw); real: 'float' | 'double'; input_shapes: shape per input slot;
        grad_slots: slots that receive gradients / tangents; param_names: run-time scalars (tracers)."""
        self.real = real
        self.outputs = outputs
        self.input_shapes = [tuple(s) for s in input_shapes]
        self.grad_slots = set(grad_slots)
        self.params = list(param_names)
        self.consts = []        # (Known, strides, is_bool)
        self._const_of = {}
        self.NO = len(outputs)
        groups = {}
        for k, (e, raw) in enumerate(outputs):
            groups.setdefault(tuple(e.shape), []).append((k, e, raw))
        self.groups = [GroupProgram(self, g, shape, outs) for g, (shape, outs) in enumerate(groups.items())]

    @staticmethod
    def c_strides(shape):
        s, out = 1, []
        for n in reversed(shape):
            out.append(s)
            s *= n
        return tuple(reversed(out))

    def slot_has_grad(self, slot):
        return slot in self.grad_slots

    def param_index(self, name):
        """prm = [w_0 .. w_{NO-1}, run-time scalars...]"""
        if name not in self.params:
            self.params.append(name)
        return self.NO + self.params.index(name)

    def weight_index(self, k):
        return k

    def const_slot(self, n):
        k = n.attrs["known"]
        if k.t.numel() <= (1 << 20):  # equal masks / coordinate arrays built twice by the operator share one slot
            import hashlib

            digest = hashlib.sha1(k.t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()
        else:
            digest = k.t.data_ptr()
        key = (digest, tuple(k.t.shape), tuple(k.shape), str(k.t.dtype))
        if key not in self._const_of:
            t = k.t
            nd = len(k.shape)
            t = t.reshape((1,) * (nd - t.dim()) + tuple(t.shape)) if t.dim() < nd else t
            cs = self.c_strides(tuple(t.shape))
            strides = tuple(0 if t.shape[a] == 1 else cs[a] for a in range(nd))
            self._const_of[key] = len(self.consts)
            self.consts.append((t, strides, t.dtype == torch.bool))
        j = self._const_of[key]
        return j, self.consts[j][1], self.consts[j][0].numel()

    # -- parameter block ------------------------------------------------------------------------------
    def layout(self):
        NI, NK, NO, NP = max(1, len(self.input_shapes)), max(1, len(self.consts)), max(1, self.NO), len(self.params)
        return NI, NK, NO, NP + self.NO

    def struct_source(self):
        NI, NK, NO, NPR = self.layout()
        return (f"struct P {{\n    long long ncell;\n    long long nblk;\n    const T* in[{NI}];\n    T* gin[{NI}];\n"
                f"    const T* tin[{NI}];\n    long long colbase[{NI}];\n    const void* kc[{NK}];\n    T* out[{NO}];\n"
                f"    const T* seed[{NO}];\n    double* partials;\n    double* sums;\n    long long* jcol;\n    T* jval;\n"
                f"    double prm[{max(1, NPR)}];\n}};\n")

    def pack(self, ncell, nblk, inp, gin, tin, colbase, kc, out, seed, partials, sums, jcol, jval, prm):
        NI, NK, NO, NPR = self.layout()

        def fill(xs, n):
            xs = list(xs) + [0] * (n - len(xs))
            return [int(x or 0) for x in xs]

        data = struct.pack("qq", int(ncell), int(nblk))
        data += struct.pack(f"{NI}Q", *fill(inp, NI)) + struct.pack(f"{NI}Q", *fill(gin, NI))
        data += struct.pack(f"{NI}Q", *fill(tin, NI)) + struct.pack(f"{NI}q", *fill(colbase, NI))
        data += struct.pack(f"{NK}Q", *fill(kc, NK)) + struct.pack(f"{NO}Q", *fill(out, NO))
        data += struct.pack(f"{NO}Q", *fill(seed, NO))
        data += struct.pack("QQQQ", int(partials or 0), int(sums or 0), int(jcol or 0), int(jval or 0))
        prm = list(prm) + [0.0] * (max(1, NPR) - len(prm))
        data += struct.pack(f"{max(1, NPR)}d", *[float(p) for p in prm])
        return data

    # -- source ---------------------------------------------------------------------------------------
    def source(self, mode):
        """CUDA source of every group's kernel for `mode` (+ the partial-sum reduction for 'lossgrad')."""
        units = []
        for g in self.groups:
            if mode == "jac":
                for j in range(len(g.results)):
                    units.append((g, f"g{g.gid}_jac{j}", *g.body(mode, j)))
            else:
                units.append((g, f"g{g.gid}_{mode}", *g.body(mode)))
        src = [PRELUDE.replace("REAL_IS_FLOAT", "1" if self.real == "float" else "0").replace("REAL", self.real),
               self.struct_source(), BLOCK_SUM if mode in ("lossgrad", "vjp") else "",
               REDUCE if mode == "lossgrad" else ""]
        for g, name, lines, nloads in units:
            text = "\n        ".join(lines)
            KG, NU = len(g.results), len(g.uniform)
            has_l, has_u = mode == "lossgrad", mode in ("lossgrad", "vjp")
            accs = ("double* lacc, " if has_l else "") + ("T* ua, " if has_u else "")
            call = ("lacc, " if has_l else "") + ("ua, " if has_u else "")
            src.append(f"ODIL_DEVICE void body_{name}(const P& a, {accs}const long long cell) {{\n        {text}\n}}\n")
            decl = ""
            if has_l:
                decl += f"    double lacc[{KG}];\n    for (int j = 0; j < {KG}; ++j) lacc[j] = 0.0;\n"
            if has_u:
                decl += f"    T ua[{max(1, NU)}];\n    for (int j = 0; j < {max(1, NU)}; ++j) ua[j] = T(0);\n"
            uni = sorted(g.uniform.items(), key=lambda kv: kv[1])
            dev = [f"extern \"C\" __global__ void __launch_bounds__({BLOCK}) k_{name}(const P a) {{", decl,
                   f"    for (long long cell = (long long)blockIdx.x * {BLOCK} + threadIdx.x; cell < a.ncell; "
                   f"cell += (long long)gridDim.x * {BLOCK})",
                   f"        body_{name}(a, {call}cell);"]
            if has_l or (has_u and NU):
                dev.append("    __shared__ double red[32];")
            if has_l:
                for j, (k, v, raw) in enumerate(g.results):
                    dev.append(f"    {{ const double s = block_sum(lacc[{j}], red); if (threadIdx.x == 0) "
                               f"a.partials[{k}ll * a.nblk + blockIdx.x] = s; }}")
            if has_u:
                for (slot, lin), j in uni:
                    dev.append(f"    {{ const double s = block_sum((double)ua[{j}], red); if (threadIdx.x == 0) "
                               f"ATOMIC_ADD(&a.gin[{slot}][{lin}], (T)s); }}")
            dev.append("}")
            host = [f"extern \"C\" void h_{name}(const P* ap) {{", "    const P& a = *ap;", decl,
                    f"    for (long long cell = 0; cell < a.ncell; ++cell) body_{name}(a, {call}cell);"]
            if has_l:
                for j, (k, v, raw) in enumerate(g.results):
                    host.append(f"    a.sums[{k}] = lacc[{j}];")
            if has_u:
                for (slot, lin), j in uni:
                    host.append(f"    a.gin[{slot}][{lin}] += ua[{j}];")
            host.append("}")
            src.append("#ifndef ODIL_HOST\n" + "\n".join(dev) + "\n#else\n" + "\n".join(host) + "\n#endif\n")
        return "\n".join(src)

    def kernels(self, mode):
        """[(group, kernel-name suffix, result index or None)] in launch order."""
        out = []
        for g in self.groups:
            if mode == "jac":
                out += [(g, f"g{g.gid}_jac{j}", j) for j in range(len(g.results))]
            else:
                out.append((g, f"g{g.gid}_{mode}", None))
        return out

    def nloads(self, g):
        return len([e for e in g.tape if e[0] == "load"])

    def gather_ok(self):
        """J^T w of every group can be gathered per input cell from the stored diagonals (mode 'vjpg')."""
        return self.dia_ok() and all(g.gather_ok() for g in self.groups)

    def dia_ok(self):
        """Every group can keep its Jacobian as per-cell diagonals (no cell-independent load receives a gradient)."""
        return all(g.pairs() is not None for g in self.groups)
