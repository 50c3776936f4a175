"""
BASELINE configs[2] as a parity case: the wave operator in two space dimensions on a (t, x, y) grid through the
public API (7 offsets, 45 region classes; not a star, so the sweep runs in the general per-cell stencil kernel).
Pinned to tests/golden/wave2.npz: loss, residual field and gradient produced by the UNMODIFIED reference core.py
running this very operator (tests/golden/make_goldens.py::gen_wave2), plus the oracle's directly written residual
and the adjoint of the traced plan; the device L-BFGS follows the reference's SciPy L-BFGS-B trajectory.  (File name sorts last on purpose: this
case was added after the round's last GPU session.)
"""
import numpy as np
import pytest
import torch

from oracle import odil_oracle as orc
from tests import operators as ops
from tests import parity
from tests.test_api_gpu import relerr, run_args, run_optimizer, set_terms

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cshape", [(10, 8, 6), (20, 16, 24)])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_wave2_eval_loss_grad(cshape, prec, golden):
    dt = np.float64 if prec == "f64" else np.float32
    problem, state = ops.make_wave2(cshape, dt)
    e = problem.extra
    U = np.random.default_rng(4).standard_normal(cshape).astype(dt)
    set_terms(problem.domain, state, [U])
    loss, grads, terms, names, norms = problem.eval_loss_grad(state)
    # reference-generated golden (unmodified reference core.py + this operator)
    g = golden("wave2")
    tag = "w2_{}_{}".format("x".join(map(str, cshape)), prec)
    assert np.array_equal(g[tag + "_U"], U)
    e_loss = abs(float(loss) - float(g[tag + "_loss"])) / float(g[tag + "_loss"])
    e_grad = relerr(grads[0].cpu().numpy(), g[tag + "_grad0"])
    e_F = relerr(problem.eval_operator(state)[0][0].numpy(), g[tag + "_F"])
    f64 = prec == "f64"
    parity.check(f"api/wave2/{tag}/loss", e_loss, 1e-11 if f64 else parity.F32_LOSS)
    parity.check(f"api/wave2/{tag}/grad", e_grad, 1e-11 if f64 else parity.F32_GRAD)
    parity.check(f"api/wave2/{tag}/F", e_F, 1e-11 if f64 else parity.F32_FIELD)
    nt, nx, ny = cshape
    bnd = {k: np.asarray(getattr(e, k), dtype=np.float64) for k in ("xlo", "xhi", "ylo", "yhi")}
    F_ref = orc.wave2_residual(U.astype(np.float64), 1.0 / nt, 2.0 / nx, 2.0 / ny, bnd,
                               np.asarray(e.init_u, dtype=np.float64), np.asarray(e.init_ut, dtype=np.float64), 1.0)
    assert names == ["fu"]
    tol = 1e-10 if prec == "f64" else 3e-4
    assert abs(float(loss) - np.mean(F_ref ** 2)) < tol * np.mean(F_ref ** 2)
    assert abs(float(norms[0]) - np.sqrt(np.mean(F_ref ** 2))) < tol * np.sqrt(np.mean(F_ref ** 2))
    # gradient: (2/n) A^T F with A probed from the directly written residual (small grid) or applied through the
    # oracle's adjoint of the traced plan (larger grid)
    from odil_b200.engine import ResidualEngine

    spec = ResidualEngine(problem, state, trace_only=True).outputs[0].blocks[0].spec
    tshape = tuple(2 * r + 1 for r in spec["rwidth"]) + (len(spec["offsets"]),)
    g_ref = orc.stencil_adjoint(F_ref, [tuple(o) for o in spec["offsets"]],
                                np.asarray(spec["table"], dtype=np.float64).reshape(tshape), spec["rwidth"],
                                2.0 / F_ref.size)
    if U.size <= 480:
        fn = lambda V: orc.wave2_residual(V, 1.0 / nt, 2.0 / nx, 2.0 / ny, bnd,
                                          np.asarray(e.init_u, dtype=np.float64),
                                          np.asarray(e.init_ut, dtype=np.float64), 1.0)
        g_probe = orc.numerical_jacobian_T(fn, U.astype(np.float64), F_ref) * (2.0 / F_ref.size)
        # (the traced table carries dt, dx, dy in the domain dtype: its coefficients are fp32-rounded for f32)
        assert relerr(g_ref, g_probe) < (1e-9 if prec == "f64" else 1e-6)
    assert relerr(grads[0].cpu().numpy(), g_ref) < tol


def test_wave2_lbfgs_converges(golden):
    problem, state = ops.make_wave2((12, 12, 12), np.float64)
    losses = run_optimizer(problem, state, "lbfgsb", run_args(epochs=60, bfgs_m=20))
    assert losses is not None and len(losses) >= 30
    # the reference's LbfgsbOptimizer (SciPy) on the same operator from the same zero start: first row of the
    # callback is the initial loss here, the loss after iteration 1 there
    ref = golden("wave2")["w2_lbfgsb_12_f64_losses"]
    mine = np.asarray(losses[1:1 + len(ref)], dtype=np.float64)
    n = min(len(mine), 10)
    err = np.max(np.abs(mine[:n] - ref[:n]) / ref[:n])
    parity.check("api/wave2/lbfgs_first10_losses", err, 1e-6)
    assert abs(mine[len(ref) - 2] - ref[-2]) < 0.5 * ref[-2]
    # SciPy's L-BFGS-B on the oracle's plan goes 65.0 -> 0.17 in 60 iterations (m = 20), rms error 0.106
    assert losses[-1] < 2e-2 * losses[0]
    u = problem.domain.arrays_from_state(state)[0].cpu().numpy()
    assert np.sqrt(np.mean((u - problem.extra.ref_u) ** 2)) < 0.2


# --------------------------------------------------------------------------------------------------
# k_tile3d (marching tile kernel for non-star 3-D plans; selected with plan_tune(variant=80))
# --------------------------------------------------------------------------------------------------
from odil_b200 import native  # noqa: E402

WAVE2 = [(0, 0, 0), (-1, 0, 0), (-2, 0, 0), (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1)]
TILE3D_CASES = [
    ((7, 10, 9), WAVE2, (2, 1, 1), 3),
    ((40, 18, 70), WAVE2, (2, 1, 1), 0),
    ((4, 5, 6), [(0, 0, 0), (1, 1, 1), (-2, 0, 2), (0, -1, 0)], (1, 1, 2), 2),
    ((33, 16, 64), [(0, 0, 0), (0, 1, -1), (1, 0, 0)], (0, 0, 0), 8),
    ((64, 48, 200), WAVE2, (2, 1, 1), 16),
]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", range(len(TILE3D_CASES)))
def test_tile3d_matches_oracle_and_generic(prec, case):
    nd, td = (np.float64, torch.float64) if prec == "f64" else (np.float32, torch.float32)
    shape, offsets, rr, zchunk = TILE3D_CASES[case]
    rng = np.random.default_rng(700 + case)
    tshape = tuple(2 * r + 1 for r in rr) + (len(offsets),)
    table = rng.standard_normal(tshape)
    U = rng.standard_normal(shape).astype(nd)
    c = rng.standard_normal(shape).astype(nd)
    scale = 2.0 / U.size
    F_ref = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, c.astype(np.float64))
    g_ref = orc.stencil_adjoint(F_ref, offsets, table, rr, scale)
    tol = (1e-11 if prec == "f64" else 2e-5) * 10
    res = {}
    for variant in (80, 81):  # 80: k_tile3d, 81: the per-cell generic kernel
        plan = native.StencilPlan(shape, td, offsets, rr, table.reshape(-1, len(offsets)))
        plan.tune(zchunk=zchunk, variant=variant)
        dU, dc = torch.as_tensor(U, device="cuda"), torch.as_tensor(c, device="cuda")
        G = torch.full_like(dU, float("nan"))
        F = torch.full_like(dU, float("nan"))
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, dc, scale, G, ss, F_out=F)
        G0 = torch.full_like(dU, float("nan"))
        ss0 = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, None, scale, G0, ss0)
        res[variant] = [t.cpu().numpy() for t in (F, G, ss, G0, ss0)]
        F, G, ss, G0, ss0 = res[variant]
        assert relerr(F, F_ref) < tol and relerr(G, g_ref) < tol
        assert abs(ss[0] - np.sum(F_ref ** 2)) < tol * np.sum(F_ref ** 2)
        F0 = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, None)
        assert abs(ss0[0] - np.sum(F0 ** 2)) < tol * np.sum(F0 ** 2)
        assert relerr(G0, orc.stencil_adjoint(F0, offsets, table, rr, scale)) < tol
    for a, b in zip(res[80][:2], res[81][:2]):
        assert relerr(a, b) < 64 * np.finfo(nd).eps


# --------------------------------------------------------------------------------------------------
# k_tile3t (TMA-fed version of k_tile3d: the default for wrap-free non-star 3-D plans with <= 8 offsets)
# --------------------------------------------------------------------------------------------------
def wrap_free_table(rng, offsets, rr):
    """Random region-typed table whose coefficients vanish wherever the neighbour would cross a face of the grid
    (what every non-periodic operator lowers to; needs rr >= the stencil radius per axis)."""
    tshape = tuple(2 * r + 1 for r in rr) + (len(offsets),)
    table = rng.standard_normal(tshape)
    for cls in np.ndindex(*tshape[:-1]):
        for o, off in enumerate(offsets):
            for a in range(3):
                ci, r, d = cls[a], rr[a], off[a]
                if (ci < r and ci + d < 0) or (ci > r and d > 2 * r - ci):
                    table[cls + (o,)] = 0.0
    return table


TILE3T_CASES = [
    ((7, 10, 12), WAVE2, (2, 1, 1), 3),
    ((40, 18, 72), WAVE2, (2, 1, 1), 0),
    ((9, 5, 8), [(0, 0, 0), (1, 1, 1), (-2, 0, 2), (0, -1, 0)], (2, 1, 2), 2),
    ((33, 16, 64), [(0, 0, 0), (0, 1, -1), (1, 0, 0)], (1, 1, 1), 8),
    ((64, 48, 200), WAVE2, (2, 1, 1), 16),
    ((70, 50, 132), WAVE2, (2, 2, 2), 0),
    ((20, 33, 260), [(0, 0, 0), (2, 0, 0), (-2, 0, 0), (0, 2, 0), (0, -2, 0), (0, 0, 2), (0, 0, -2), (1, 1, 1)],
     (2, 2, 2), 0),
]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", range(len(TILE3T_CASES)))
def test_tile3t_matches_oracle_tile3d_and_generic(prec, case, monkeypatch):
    nd, td = (np.float64, torch.float64) if prec == "f64" else (np.float32, torch.float32)
    shape, offsets, rr, zchunk = TILE3T_CASES[case]
    rng = np.random.default_rng(900 + case)
    table = wrap_free_table(rng, offsets, rr)
    U = rng.standard_normal(shape).astype(nd)
    c = rng.standard_normal(shape).astype(nd)
    scale = 2.0 / U.size
    F_ref = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, c.astype(np.float64))
    g_ref = orc.stencil_adjoint(F_ref, offsets, table, rr, scale)
    F0 = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, None)
    g0_ref = orc.stencil_adjoint(F0, offsets, table, rr, scale)
    tol = (1e-11 if prec == "f64" else 2e-5) * 10
    res = {}
    for name, env, variant in (("tile3t", "1", 80), ("tile3d", "0", 80), ("generic", "0", 81)):
        monkeypatch.setenv("ODIL_B200_TILE3T", env)
        plan = native.StencilPlan(shape, td, offsets, rr, table.reshape(-1, len(offsets)))
        plan.tune(zchunk=zchunk, variant=variant)
        dU, dc = torch.as_tensor(U, device="cuda"), torch.as_tensor(c, device="cuda")
        G = torch.full_like(dU, float("nan"))
        F = torch.full_like(dU, float("nan"))
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, dc, scale, G, ss, F_out=F)
        G0 = torch.full_like(dU, float("nan"))
        ss0 = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, None, scale, G0, ss0)
        res[name] = [t.cpu().numpy() for t in (F, G, G0, ss, ss0)]
        F, G, G0, ss, ss0 = res[name]
        assert relerr(F, F_ref) < tol and relerr(G, g_ref) < tol and relerr(G0, g0_ref) < tol, name
        assert abs(ss[0] - np.sum(F_ref ** 2)) < tol * np.sum(F_ref ** 2), name
        assert abs(ss0[0] - np.sum(F0 ** 2)) < tol * np.sum(F0 ** 2), name
    # same operations per cell in the same order: F and g agree bit for bit (up to the sign of zero)
    for a, b in zip(res["tile3t"][:3], res["tile3d"][:3]):
        assert np.array_equal(a, b)
    for a, b in zip(res["tile3t"][:3], res["generic"][:3]):
        assert relerr(a, b) < 64 * np.finfo(nd).eps
