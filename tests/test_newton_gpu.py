"""
Newton path for affine operators: Problem.linearize / eval_operator_grad / linsolver.solve /
util.optimize_newton on the device, checked the way the reference checks it (tests/test_newton.py:115-148: one
Newton step solves a linear problem) and against an explicit Jacobian obtained by probing the operator.
"""
import argparse

import numpy as np
import pytest
import torch

import odil
from odil_b200 import linsolver, newton
from tests import operators as ops

pytestmark = pytest.mark.gpu


def two_field_operator(ctx):
    u, v = ctx.field("u"), ctx.field("v")
    um, vp = ctx.field("u", -1, 0), ctx.field("v", 0, 1)
    f1 = 2 * u - um + 0.5 * vp - ctx.extra.r1
    f2 = v + 0.25 * u - ctx.extra.r2
    return [("f1", f1), ("f2", f2)]


def make_two_field(cshape=(6, 5), dtype=np.float64):
    domain = odil.Domain(cshape=list(cshape), dimnames=["x", "y"], multigrid=False, dtype=dtype)
    rng = np.random.default_rng(5)
    extra = argparse.Namespace(r1=rng.standard_normal(cshape).astype(dtype), r2=rng.standard_normal(cshape).astype(dtype))
    state = odil.State()
    state.fields["u"] = None
    state.fields["v"] = None
    state = domain.init_state(state)
    return odil.Problem(two_field_operator, domain, extra), state


def probe_jacobian(problem, state):
    """Dense Jacobian by evaluating the (affine) operator on unit vectors."""
    domain = problem.domain
    packed0 = domain.pack_state(state).full().clone()
    n = packed0.numel()

    def F(p):
        domain.unpack_state(p, state)
        vals, _ = problem.eval_operator(state)
        return np.concatenate([np.asarray(v).reshape(-1) for v in vals])

    f0 = F(torch.zeros_like(packed0))
    J = np.zeros((f0.size, n))
    for j in range(n):
        e = torch.zeros_like(packed0)
        e[j] = 1
        J[:, j] = F(e) - f0
    domain.unpack_state(packed0, state)
    return J, f0


@pytest.mark.parametrize("which", ["poisson2d", "poisson3d", "two_field"])
def test_linearize_matches_probed_jacobian(which):
    if which == "poisson2d":
        problem, state = ops.make_poisson((7, 6), 0, np.float64)
    elif which == "poisson3d":
        problem, state = ops.make_poisson((5, 4, 6), 0, np.float64)
    else:
        problem, state = make_two_field()
    J, f0 = probe_jacobian(problem, state)
    vector, matrix = problem.linearize(state)
    assert matrix.shape == J.shape
    assert np.allclose(vector.cpu().numpy(), f0, rtol=0, atol=1e-12 * max(1, np.abs(f0).max()))
    assert np.allclose(matrix.tocsr().toarray(), J, rtol=0, atol=1e-10 * np.abs(J).max())
    rng = np.random.default_rng(0)
    x = rng.standard_normal(J.shape[1])
    y = rng.standard_normal(J.shape[0])
    tx, ty = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")
    assert np.allclose(matrix.matvec(tx).cpu().numpy(), J @ x, rtol=0, atol=1e-10 * np.abs(J).max())
    assert np.allclose(matrix.rmatvec(ty).cpu().numpy(), J.T @ y, rtol=0, atol=1e-10 * np.abs(J).max())
    # per-(key, shift) coefficient arrays, as eval_operator_grad returns them (core.py:1341-1350)
    values, grads, names = problem.eval_operator_grad(state)
    assert len(grads) == len(values) == len(names)
    for d in grads:
        for (key, shift, loc), coef in d.items():
            assert tuple(np.asarray(coef).shape) == tuple(problem.domain.cshape)


@pytest.mark.parametrize("solver", ["direct", "cg_b200"])
def test_one_newton_step_solves_linear_problem(solver):
    problem, state = ops.make_poisson((12, 10), 0, np.float64)
    args = argparse.Namespace(linsolver=solver, linsolver_tol=1e-13, linsolver_maxiter=4000, linsolver_damp=0,
                              linsolver_dampdiag=0, linsolver_verbose=0, epochs=1, epoch_start=0)
    arrays, info = odil.util.optimize_newton(args, problem, state)
    u = np.asarray(problem.domain.field(state, "u"))
    ref = problem.extra.ref_u
    assert np.sqrt(np.mean((u - ref) ** 2)) < 1e-6  # the reference's bar (tests/test_newton.py:142)
    loss = float(problem.eval_loss_grad(state)[0])
    assert loss < 1e-12


def test_cg_status_and_two_fields():
    problem, state = make_two_field((8, 8))
    vector, matrix = problem.linearize(state)
    status = {}
    args = argparse.Namespace(linsolver_tol=1e-12, linsolver_maxiter=500, linsolver_damp=0)
    delta = linsolver.solve(matrix, -vector, args, status, "cg_b200")
    A = matrix.tocsr()
    import scipy.sparse.linalg

    ref = scipy.sparse.linalg.spsolve((A.T @ A).tocsc(), A.T @ (-vector.cpu().numpy()))
    assert np.allclose(delta.cpu().numpy(), ref, rtol=0, atol=1e-8 * np.abs(ref).max())
    assert status["niter"] > 0 and status["residual"] <= 1e-12 < status["residual0"]


def test_newton_rejects_multigrid():
    problem, state = ops.make_poisson((16, 16), 2, np.float64)
    with pytest.raises(NotImplementedError):
        problem.linearize(state)
