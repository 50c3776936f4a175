"""
GPU parity of the general (non-affine) path: the cases of tests/nonaffine_cases.py -- UNMODIFIED reference operators
(heat with k(u) and with a neural net, velocity-from-tracer, heat_tmax, infer_constant, the operators of the
reference's tests/test_optimize.py and tests/test_newton.py, Poisson with mgloss, a Context.Raw term) -- evaluated
through the public API (Problem.eval_loss_grad / eval_operator / linearize) by the NVRTC-compiled kernels, against
goldens produced by the unmodified reference (tests/golden/make_nonaffine_goldens.py).
"""
import numpy as np
import pytest
import torch

import odil
from odil_b200 import linsolver
from odil_b200.engine_graph import GraphEngine
from tests import nonaffine_cases as cases
from tests import parity, refsrc
from tests.test_graph_cpu import build_case, golden_arrays, relerr

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refsrc.available(), reason="reference scripts absent")]


def load_state(problem, state, g, key, dt):
    domain = problem.domain
    x = [domain.mod.variable(a, dtype=dt) for a in golden_arrays(g, key, "x")]
    domain.arrays_to_state(x, state)
    return x


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", cases.CASES)
def test_eval_loss_grad_matches_reference_golden(golden, case, prec):
    g = golden("nonaffine")
    key = f"{case}_{prec}"
    problem, state, dt = build_case(case, prec, device="cuda")
    load_state(problem, state, g, key, dt)
    loss, grads, terms, names, norms = problem.eval_loss_grad(state)
    assert isinstance(problem._engine(state), GraphEngine)
    assert names == [str(n) for n in g[key + "_names"]]
    f64 = prec == "f64"
    parity.check(f"graph/{key}/loss", abs(float(loss) - float(g[key + "_loss"])) / abs(float(g[key + "_loss"])),
                 1e-11 if f64 else 2e-6)
    parity.check(f"graph/{key}/terms", relerr([float(t) for t in terms], g[key + "_terms"]), 1e-11 if f64 else 2e-6)
    raws = [bool(r) for r in g[key + "_raws"]]
    for t, n, r in zip(g[key + "_terms"], norms, raws):
        ref = t if r else np.sqrt(t)
        assert abs(float(n) - ref) <= (1e-10 if f64 else 1e-5) * max(abs(ref), 1e-30)
    for i, (gi, gr) in enumerate(zip(grads, golden_arrays(g, key, "g"))):
        assert tuple(gi.shape) == gr.shape
        parity.check(f"graph/{key}/grad{i}", relerr(gi.cpu().numpy(), gr), 1e-11 if f64 else 1e-5)
    values, names2 = problem.eval_operator(state)
    for i, (F, Fr) in enumerate(zip(values, golden_arrays(g, key, "F"))):
        parity.check(f"graph/{key}/F{i}", relerr(np.asarray(F), Fr), 1e-11 if f64 else 5e-6)
    # a second evaluation (buffers reused, gradients re-zeroed) gives the same numbers up to the atomic order
    loss2, grads2, _, _, _ = problem.eval_loss_grad(state)
    assert abs(float(loss2) - float(loss)) <= 1e-12 * abs(float(loss))
    for a, b in zip(grads, grads2):
        assert relerr(a.cpu().numpy(), b.cpu().numpy()) < (1e-13 if f64 else 1e-5)


@pytest.mark.parametrize("case", cases.NEWTON_CASES)
@pytest.mark.parametrize("dia", ["0", "1"])
def test_linearize_matches_reference_golden(golden, case, dia, monkeypatch):
    """dia = 1 (default): the Jacobian products come from per-cell diagonals stored once per linearisation (generated
    kernels 'jacd' / 'jvpd' / 'vjpd') where the operator allows it; dia = 0: from the forward- / reverse-mode kernels."""
    monkeypatch.setenv("ODIL_B200_NEWTON_DIA", dia)
    g = golden("nonaffine")
    key = f"{case}_f64"
    problem, state, dt = build_case(case, "f64", device="cuda")
    load_state(problem, state, g, key, dt)
    vector, matrix = problem.linearize(state)
    Jr = g[key + "_jac"]
    F = np.concatenate([f.reshape(-1) for f in golden_arrays(g, key, "F")])
    parity.check(f"graph/{key}/linearize_vector", relerr(vector.cpu().numpy(), F), 1e-11)
    J = matrix.tocsr().toarray()
    assert J.shape == Jr.shape == matrix.shape
    parity.check(f"graph/{key}/linearize_csr", np.max(np.abs(J - Jr)) / np.max(np.abs(Jr)), 1e-11)
    rng = np.random.default_rng(0)
    v = torch.as_tensor(rng.standard_normal(J.shape[1]), device="cuda")
    w = torch.as_tensor(rng.standard_normal(J.shape[0]), device="cuda")
    tag = "d" if dia == "1" and matrix.engine.gen.dia_ok() else ""
    parity.check(f"graph/{key}/jvp{tag}", relerr(matrix.matvec(v).cpu().numpy(), Jr @ v.cpu().numpy()), 1e-11)
    parity.check(f"graph/{key}/vjp{tag}", relerr(matrix.rmatvec(w).cpu().numpy(), Jr.T @ w.cpu().numpy()), 1e-11)
    assert bool(matrix._dia) == (tag == "d")
    if case == "heat3":
        assert tag == ("d" if dia == "1" else "")
    # SciPy-matrix surface the reference's scripts rely on (tests/test_newton.py:117-120)
    normal = matrix.T @ matrix
    assert np.allclose(normal.toarray(), Jr.T @ Jr, rtol=1e-10, atol=1e-12 * np.max(np.abs(Jr)) ** 2)
    assert np.allclose(matrix.T @ F, Jr.T @ F, rtol=1e-10, atol=1e-10)


def test_newton_step_cg_matches_direct_solve(golden):
    """One Newton step of the heat problem with k(u): matrix-free CG on the device (linsolver cg_b200) against the
    reference's direct solve of the normal equations on the assembled matrix."""
    import argparse

    g = golden("nonaffine")
    problem, state, dt = build_case("heat_k", "f64", device="cuda")
    load_state(problem, state, g, "heat_k_f64", dt)
    vector, matrix = problem.linearize(state)
    args = argparse.Namespace(linsolver_tol=1e-13, linsolver_maxiter=2000, linsolver_damp=0, linsolver_dampdiag=0)
    status = {}
    d_cg = linsolver.solve(matrix, -vector, args, status, "cg_b200").cpu().numpy()
    d_direct = linsolver.solve(matrix, -vector, args, {}, "direct").cpu().numpy()
    parity.check("graph/heat_k_f64/newton_step_cg_vs_direct", relerr(d_cg, d_direct), 1e-6)


def test_epoch_tracer_is_a_runtime_parameter(golden):
    """tracers['epoch'] feeds annealed weights (heat.py:43,118): changing it changes the loss WITHOUT re-tracing."""
    g = golden("nonaffine")
    problem, state, dt = build_case("heat_k", "f64", device="cuda")
    load_state(problem, state, g, "heat_k_f64", dt)
    loss7 = float(problem.eval_loss_grad(state)[0])
    engine = problem._engine(state)
    problem.tracers["epoch"] = 0
    loss0 = float(problem.eval_loss_grad(state)[0])
    assert problem._engine(state) is engine
    terms7 = g["heat_k_f64_terms"]
    # xreg term: weight kxreg * 0.5 ** (epoch / 5) -> term scales by 0.5 ** (-2 * 7 / 5)
    expect = float(g["heat_k_f64_loss"]) + terms7[2] * (0.5 ** (-14 / 5) - 1)
    assert abs(loss7 - float(g["heat_k_f64_loss"])) < 1e-11 * loss7
    assert abs(loss0 - expect) < 1e-10 * abs(expect)


@pytest.mark.parametrize("opt,epochs", [("adam", 60), ("lbfgsb", 20)])
def test_optimizers_drive_a_nonaffine_problem(opt, epochs, golden):
    """The optimizers run on a non-affine problem (heat with a neural-net conductivity): the loss goes down."""
    from tests.test_api_gpu import run_args, run_optimizer

    g = golden("nonaffine")
    problem, state, dt = build_case("heat_knet", "f64", device="cuda")
    load_state(problem, state, g, "heat_knet_f64", dt)
    losses = run_optimizer(problem, state, opt, run_args(epochs=epochs, lr=0.01))
    assert losses is not None and losses[-1] < 0.7 * losses[0]


@pytest.mark.parametrize("case", ["newton", "heat_k", "heat3"])
def test_eval_operator_grad_diagonals_match_reference_jacobian(golden, case):
    """Problem.eval_operator_grad for non-affine operators: per output {(key, shift, loc): dF/d(shifted field)} like
    `_eval_operator_grad_tf` (core.py:1313-1361) -- every entry of the reference-generated dense Jacobian that the
    diagonals claim is reproduced, and together with the Array / NeuralNet columns they account for all of it."""
    g = golden("nonaffine")
    key = f"{case}_f64"
    problem, state, dt = build_case(case, "f64", device="cuda")
    load_state(problem, state, g, key, dt)
    values, grads, names = problem.eval_operator_grad(state)
    engine = problem._engine(state)
    Jr = g[key + "_jac"].copy()
    sizes = [a.numel() for a in problem.domain.arrays_from_state(state)]
    col0 = np.concatenate([[0], np.cumsum(sizes)])
    row0 = np.concatenate([[0], np.cumsum([o.n for o in engine.outputs])])
    assert len(grads) == len(values) == len(names)
    nfield = 0
    for k, d in enumerate(grads):
        for desc, coef in d.items():
            coef = np.asarray(coef)
            if desc[1] is None:  # Array unknown: dense block
                unk = engine.unknowns[desc[0]]
                lo = int(col0[unk.first])
                blk = coef.reshape(engine.outputs[k].n, -1)
                assert np.allclose(blk, Jr[row0[k]:row0[k + 1], lo:lo + blk.shape[1]], rtol=1e-11, atol=1e-13)
                Jr[row0[k]:row0[k + 1], lo:lo + blk.shape[1]] = 0
                continue
            cmap = engine.trace.column_map(desc, col0).reshape(-1)
            cells = np.arange(engine.outputs[k].n)
            ok = cmap >= 0
            ref = np.zeros(engine.outputs[k].n)
            ref[ok] = Jr[row0[k] + cells[ok], cmap[ok]]
            assert np.allclose(coef.reshape(-1), ref, rtol=1e-11, atol=1e-13), (case, k, desc)
            Jr[row0[k] + cells[ok], cmap[ok]] = 0
            nfield += 1
    assert nfield > 0
    # what is left of the Jacobian belongs to NeuralNet weights only
    left = np.abs(Jr).sum(axis=0) > 0
    for key_, unk in engine.unknowns.items():
        if unk.kind != "NeuralNet":
            for i in range(unk.first, unk.first + unk.narrays):
                assert not left[col0[i]:col0[i + 1]].any(), key_
