"""
Newton's method for AFFINE operators: matrix-free Jacobian on the device + conjugate gradients
on the normal equations.

Reference path: `Problem.linearize` (src/odil/core.py:1113-1217) differentiates the operator per (key, shift)
with a TF GradientTape (`_eval_operator_grad_tf`, :1313-1361; the JAX version raises NotImplementedError,
:1363-1364), assembles a SciPy CSR matrix from `rolled(arange)` column indices (:1144-1171) and
`util.optimize_newton` (src/odil/util.py:152-187) solves `M^T M delta = M^T (-F)` with SuperLU
(src/odil/linsolver.py:18-26).

Here the traced operator is already F(u) = sum_blocks A_blk U_key + c with region-typed coefficient tables
(engine.lower_block), so the Jacobian IS the set of stencil plans: J x and J^T y are the forward / adjoint stencil
kernels (odil_b200_stencil_forward / _adjoint), the per-(key, shift) "diagonals" the reference extracts by AD are
the table rows expanded to the grid, and the normal equations are solved by CG whose vector algebra runs in
odil_b200_dot / odil_b200_cg_update_xr / odil_b200_cg_update_p with the step scalars kept on the device.
`StencilJacobian.tocsr()` assembles the same matrix as the reference's `field_to_matrix` for small grids (parity
tests, and the SciPy solvers of linsolver.solve).
"""
import math

import numpy as np
import torch

from . import native


class StencilJacobian:
    """J = dF/d(packed state) of an affine operator.  Rows: the operator outputs in order, flattened; columns:
    `domain.pack_state` order.  Frozen fields (ctx.field(..., frozen=True)) do not contribute (stop_gradient,
    core.py:973-974)."""

    def __init__(self, engine):
        if engine.slab is not None:
            raise NotImplementedError("Newton on slab-decomposed grids")
        self.engine = engine
        self.dtype = engine.tdtype
        self.device = engine.device
        self.col_start, self.col_shape = {}, {}
        ncols = 0
        self.unk_order = []
        for key, unk in engine.unknowns.items():
            if unk.kind == "MultigridField" and unk.narrays > 1:
                raise NotImplementedError("Newton needs multigrid off (as in the reference, examples/wave/README.md:27)")
            if unk.kind == "NeuralNet":
                raise NotImplementedError("NeuralNet unknowns are not on the Newton path yet (SURVEY.md 8f-3)")
            self.col_start[key] = ncols
            self.col_shape[key] = tuple(unk.shapes[0])
            self.unk_order.append(key)
            ncols += math.prod(unk.shapes[0])
        self.row_start = []
        nrows = 0
        for out in engine.outputs:
            self.row_start.append(nrows)
            nrows += out.n
        self.shape = (nrows, ncols)

    # ---- products -----------------------------------------------------------------------------------
    def _cols(self, x, key):
        s = self.col_start[key]
        return x[s: s + math.prod(self.col_shape[key])].view(self.col_shape[key])

    def matvec(self, x):
        """y = J x (device tensor of length ncols -> length nrows)."""
        x = x.to(self.dtype).contiguous()
        y = torch.zeros(self.shape[0], dtype=self.dtype, device=self.device)
        for k, out in enumerate(self.engine.outputs):
            yk = y[self.row_start[k]: self.row_start[k] + out.n].view(out.shape)
            first = True
            for blk in out.blocks:
                if blk.frozen:
                    continue
                blk.plan.forward(self._cols(x, blk.key), None if first else yk, yk)
                first = False
        return y

    def rmatvec(self, y):
        """x = J^T y."""
        y = y.to(self.dtype).contiguous()
        x = torch.zeros(self.shape[1], dtype=self.dtype, device=self.device)
        touched = set()
        for k, out in enumerate(self.engine.outputs):
            yk = y[self.row_start[k]: self.row_start[k] + out.n].view(out.shape)
            for blk in out.blocks:
                if blk.frozen:
                    continue
                xk = self._cols(x, blk.key)
                blk.plan.adjoint(yk, 1.0, xk if blk.key in touched else None, xk)
                touched.add(blk.key)
        return x

    def dot(self, x):
        return self.matvec(x)

    # ---- explicit forms (small grids: parity with the reference's CSR assembly) --------------------
    def diagonals(self):
        """Per output: {(key, shift): coefficient array on the grid} -- what `_eval_operator_grad_tf` returns
        (core.py:1341-1350), here expanded from the region-typed tables."""
        res = []
        for out in self.engine.outputs:
            d = {}
            for blk in out.blocks:
                if blk.frozen:
                    continue
                sp = blk.spec
                shape, rr = sp["shape"], sp["rwidth"]
                cls = np.zeros(shape, dtype=np.int64)
                for a, n in enumerate(shape):
                    i = np.arange(n)
                    r = rr[a]
                    c = np.where(i < r, i, np.where(n - 1 - i < r, 2 * r - (n - 1 - i), r))
                    cls = cls * (2 * r + 1) + c.reshape([-1 if b == a else 1 for b in range(len(shape))])
                for o, off in enumerate(sp["offsets"]):
                    key = (blk.key, tuple(int(v) for v in off))
                    d[key] = d.get(key, 0) + sp["table"][cls, o]
            res.append(d)
        return res

    def tocsr(self):
        """SciPy CSR of J (host).  Column of (cell x, shift s) = linear index of (x + s) mod N, like the reference's
        rolled `arange` (core.py:1144-1171)."""
        import scipy.sparse

        nrows, ncols = self.shape
        if nrows * 8 > 2 ** 31:
            raise MemoryError("tocsr() is meant for small grids; use the matrix-free products")
        rows, cols, vals = [], [], []
        for k, (out, diag) in enumerate(zip(self.engine.outputs, self.diagonals())):
            lin = np.arange(out.n).reshape(out.shape)
            for (key, shift), coef in diag.items():
                col = lin
                for a, s in enumerate(shift):
                    col = np.roll(col, -s, axis=a)
                mask = coef != 0
                rows.append(self.row_start[k] + lin[mask])
                cols.append(self.col_start[key] + col[mask])
                vals.append(coef[mask])
        if not rows:
            return scipy.sparse.csr_matrix((nrows, ncols))
        return scipy.sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                       shape=(nrows, ncols))


def residual_vector(engine, arrays):
    """F(u) of every output, flattened and concatenated (the `vector` of Problem.linearize, core.py:1127-1138)."""
    vals = engine.operator_values(arrays)
    return torch.cat([v.full().reshape(-1) if hasattr(v, "full") else torch.as_tensor(v).reshape(-1) for v in vals])


def cg_normal(jac, rhs, tol=1e-6, maxiter=None, damp=0.0, status=None, check_every=10):
    """
    Solves (J^T J + damp^2 I) x = J^T rhs by conjugate gradients on the device.  Stops when the RMS of the
    normal-equation residual drops below `tol` (the same measure as the reference's bicgstab callback,
    linsolver.py:76-80) or after `maxiter` iterations.  One host synchronisation every `check_every` iterations.
    """
    dtype, dev = jac.dtype, jac.device
    n = jac.shape[1]
    maxiter = maxiter if maxiter is not None else 10 * n
    b = jac.rmatvec(rhs.to(dtype))
    x = torch.zeros_like(b)
    r = b.clone()
    p = b.clone()
    rs = torch.zeros(1, dtype=torch.float64, device=dev)
    rs_new = torch.zeros(1, dtype=torch.float64, device=dev)
    pq = torch.zeros(1, dtype=torch.float64, device=dev)
    native.dot(r, r, rs)
    res0 = math.sqrt(float(rs) / n)
    res, it = res0, 0
    while it < maxiter and res > tol:
        for _ in range(min(check_every, maxiter - it)):
            q = jac.rmatvec(jac.matvec(p))
            if damp:
                native.axpby(damp ** 2, p, 1.0, q)
            native.dot(p, q, pq)
            native.cg_update_xr(rs, pq, p, q, x, r)
            native.dot(r, r, rs_new)
            native.cg_update_p(rs_new, rs, r, p)
            rs, rs_new = rs_new, rs
            it += 1
        res = math.sqrt(float(rs) / n)
        if not math.isfinite(res):
            break
    if status is not None:
        status.update(residual=res, residual0=res0, niter=it)
    return x
