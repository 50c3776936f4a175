"""
CPU port of the reference's per-epoch arithmetic on torch-CPU  --  TEST/BENCH INFRASTRUCTURE, NOT PRODUCT.

Used only by bench.py (`cpu_baseline` leg and `--impl reference`) to time, on the GPU box's host cores,
the same sequence of array operations the reference executes per epoch on its JAX-CPU backend (which
cannot be installed here: no jax/tensorflow wheels, no network -- SURVEY.md 8c):

  multigrid_to_regular via interp_to_finer "stack" (pad 2*sym-reflect, 2^d weighted rolls, stack /
  reshape interleave; core.py:245-263, :606-700)  ->  ctx.field = roll (core.py:963)  ->  Poisson
  operator with where() boundary rows (examples/poisson/poisson.py:57-68, :100-113)  ->
  mean(square(F)) (core.py:1093)  ->  reverse-mode gradient (torch.autograd standing in for
  jax.value_and_grad, core.py:1100)  ->  Adam update (optimizer.py:311-319).

Checked against the NumPy oracle in tests/test_oracle_golden.py::test_torch_port_matches_oracle.
"""
import itertools

import numpy as np
import torch


def _pad_lin_extrap(u):
    """upad = 2*symmetric - reflect jointly on every axis (core.py:640-643)."""
    sym, ref = u, u
    for ax in range(u.dim()):
        n = u.shape[ax]
        idx_s = torch.tensor([0] + list(range(n)) + [n - 1])
        idx_r = torch.tensor([1] + list(range(n)) + [n - 2])
        sym = sym.index_select(ax, idx_s)
        ref = ref.index_select(ax, idx_r)
    return 2 * sym - ref


def interp_to_finer_cells(u):
    """'stack' interpolation for loc 'c'*ndim (core.py:668-696)."""
    d = u.dim()
    up = _pad_lin_extrap(u)
    corners = list(itertools.product([0, 1], repeat=d))
    dims = list(range(d))
    outs = []
    for tgt in corners:
        acc = 0
        wsum = 0
        for src in corners:
            w = 3 ** sum(1 - abs(a - b) for a, b in zip(src, tgt))
            acc = acc + w * torch.roll(up, [-s for s in src], dims)
            wsum += w
        outs.append(acc / wsum)
    res = torch.stack(outs).reshape((2,) * d + tuple(up.shape))
    # interleave: (p0, p1, .., n0, n1, ..) -> (n0, p0, n1, p1, ..)
    perm = []
    for a in range(d):
        perm += [d + a, a]
    res = res.permute(*perm).reshape([2 * s for s in up.shape])
    sl = tuple(slice(1, 2 * s - 3) for s in up.shape)
    return res[sl]


def multigrid_to_regular(terms):
    res = terms[-1]
    for t in reversed(terms[:-1]):
        res = t + interp_to_finer_cells(res)
    return res


def poisson_operator(U, rhs, steps):
    nd = U.dim()
    F = -rhs
    for ax in range(nd):
        n = U.shape[ax]
        um = torch.roll(U, 1, ax)
        up = torch.roll(U, -1, ax)
        idx = torch.arange(n).reshape([-1 if a == ax else 1 for a in range(nd)])
        qm = torch.where(idx == 0, (up - 6 * U) / 3, um)
        qp = torch.where(idx == n - 1, (um - 6 * U) / 3, up)
        F = F + (qp - 2 * U + qm) / steps[ax] ** 2
    return F


class PoissonAdamEpoch:
    """State + one epoch = loss/grad evaluation and Adam update, as optimizer.py:331-336 loops it."""

    def __init__(self, cshape, nlvl, dtype=torch.float32, lr=0.005, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.cshape = tuple(cshape)
        self.steps = [1.0 / n for n in cshape]
        shapes = [tuple(n >> l for n in cshape) for l in range(nlvl)]
        self.x = [torch.zeros(s, dtype=dtype) for s in shapes]
        self.m = [torch.zeros_like(a) for a in self.x]
        self.v = [torch.zeros_like(a) for a in self.x]
        self.rhs = torch.randn(self.cshape, generator=g, dtype=dtype)
        self.lr, self.t, self.dtype = lr, 0, dtype

    def loss_grad(self):
        leaves = [a.detach().requires_grad_(True) for a in self.x]
        U = multigrid_to_regular(leaves) if len(leaves) > 1 else leaves[0]
        F = poisson_operator(U, self.rhs, self.steps)
        loss = torch.mean(torch.square(F))
        grads = torch.autograd.grad(loss, leaves)
        return loss.detach(), grads

    def step(self):
        self.t += 1
        loss, grads = self.loss_grad()
        npdt = np.float32 if self.dtype == torch.float32 else np.float64
        lr, b1, b2, t = npdt(self.lr), npdt(0.9), npdt(0.999), npdt(self.t)
        alpha = float(lr * np.sqrt(npdt(1) - b2 ** t) / (npdt(1) - b1 ** t))
        omb1, omb2 = float(npdt(1) - b1), float(npdt(1) - b2)
        for i, g in enumerate(grads):
            self.m[i] = self.m[i] + (g - self.m[i]) * omb1
            self.v[i] = self.v[i] + (torch.square(g) - self.v[i]) * omb2
            self.x[i] = self.x[i] - (self.m[i] * alpha) / (torch.sqrt(self.v[i]) + 1e-7)
        return loss
