"""
Pins the CPU oracle (oracle/odil_oracle.py) against the golden vectors generated from the
unmodified reference (tests/golden/make_goldens.py).  No GPU, no /root/reference.
"""
import numpy as np
import pytest

from oracle import odil_oracle as orc

EPS64 = np.finfo(np.float64).eps


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


LOCS = ["cccc", "nnnn", "cnnn", "nccc", "c.cn"]


@pytest.mark.parametrize("ndim", [1, 2, 3, 4])
@pytest.mark.parametrize("loc4", LOCS)
def test_interp(golden, ndim, loc4):
    g = golden("transfers")
    loc = loc4[:ndim]
    u = g[f"interp_{ndim}_{loc}_in"]
    ref = g[f"interp_{ndim}_{loc}_out"]
    out = orc.interp_to_finer(u, loc)
    assert out.shape == ref.shape
    assert relerr(out, ref) < 50 * EPS64


@pytest.mark.parametrize("ndim", [1, 2, 3, 4])
@pytest.mark.parametrize("loc4", LOCS)
def test_interp_adjoint_is_transpose(golden, ndim, loc4):
    g = golden("transfers")
    loc = loc4[:ndim]
    u = g[f"interp_{ndim}_{loc}_in"]
    fine = g[f"interp_{ndim}_{loc}_out"]
    rng = np.random.default_rng(3)
    w = rng.standard_normal(fine.shape)
    lhs = np.sum(orc.interp_to_finer(u, loc) * w)
    rhs = np.sum(u * orc.interp_adjoint(w, loc, u.shape))
    assert abs(lhs - rhs) < 1e-12 * max(abs(lhs), 1)


@pytest.mark.parametrize("ndim", [1, 2, 3, 4])
@pytest.mark.parametrize("loc4", LOCS[:4])
def test_restrict(golden, ndim, loc4):
    g = golden("transfers")
    loc = loc4[:ndim]
    out = orc.restrict_to_coarser(g[f"restrict_{ndim}_{loc}_in"], loc)
    ref = g[f"restrict_{ndim}_{loc}_out"]
    assert out.shape == ref.shape
    assert relerr(out, ref) < 50 * EPS64


def test_interp_restrict_linear_exact():
    """Property pinned by reference tests/test_mg_interp.py:31 and tests/test_mg_restrict.py:40."""
    for ndim in [1, 2, 3]:
        for loc4 in ["cccc", "nnnn", "cnnn", "nccc"]:
            loc = loc4[:ndim]
            csh = 3 + np.arange(ndim)
            cs = csh * 2

            def pts(c):
                xs = [np.arange(n + (l == "n")) / n + (0.5 / n if l == "c" else 0) for n, l in zip(c, loc)]
                return np.meshgrid(*xs, indexing="ij")

            f = lambda xx: sum(x * np.sqrt(i + 1) for i, x in enumerate(xx))
            assert np.max(np.abs(orc.interp_to_finer(f(pts(csh)), loc) - f(pts(cs)))) < 100 * EPS64

            def fj(xx):
                r = 0
                for i, x in enumerate(xx):
                    r = r + x * (i + 1) + 10.0 * (x == 0) + 10.0 * (x == 1)
                return r

            assert np.max(np.abs(orc.restrict_to_coarser(fj(pts(cs)), loc) - fj(pts(csh)))) < 100 * EPS64


POISSON_CASES = {
    "p1d_16_L0": ((16,), 0), "p1d_16_L3": ((16,), 3), "p2d_16_L3": ((16, 16), 3), "p2d_12x8_L0": ((12, 8), 0),
    "p3d_8_L3": ((8, 8, 8), 3), "p3d_16x8x12_L2": ((16, 8, 12), 2), "p3d_12_L0": ((12, 12, 12), 0),
}


def _terms(g, tag):
    terms, grads, i = [], [], 0
    while f"{tag}_term{i}" in g.files:
        terms.append(g[f"{tag}_term{i}"])
        grads.append(g[f"{tag}_grad{i}"])
        i += 1
    return terms, grads


@pytest.mark.parametrize("name", list(POISSON_CASES))
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_poisson_loss_grad(golden, name, prec):
    g = golden("poisson")
    cshape, nlvl = POISSON_CASES[name]
    tag = f"{name}_{prec}"
    dt = np.float64 if prec == "f64" else np.float32
    tol = 1e-11 if prec == "f64" else 2e-4
    terms, grads = _terms(g, tag)
    ndim = len(cshape)
    steps = [dt(1) / dt(n) for n in cshape]
    loc = "c" * ndim
    U = orc.mg_synthesize(terms, loc) if len(terms) > 1 else terms[0]
    assert relerr(U, g[tag + "_U"]) < (1e-13 if prec == "f64" else 1e-5)
    F = orc.poisson_residual(U, g[tag + "_rhs"], steps)
    assert relerr(F, g[tag + "_F"]) < tol
    offsets, table, rr = orc.poisson_plan(ndim, steps)
    loss, gr, F2, _ = orc.eval_loss_grad_plan(terms, loc, offsets, table, rr, -g[tag + "_rhs"])
    assert relerr(F2, g[tag + "_F"]) < tol
    assert abs(loss - g[tag + "_loss"]) < tol * abs(g[tag + "_loss"])
    for a, b in zip(gr, grads):
        assert a.shape == b.shape
        assert relerr(a, b) < tol


@pytest.mark.parametrize("name,cshape", [("w_16x12_L0", (16, 12)), ("w_16x8_L2", (16, 8))])
def test_wave_loss_grad(golden, name, cshape):
    g = golden("wave")
    tag = name + "_f64"
    terms, grads = _terms(g, tag)
    U = orc.mg_synthesize(terms, "cc") if len(terms) > 1 else terms[0]
    dt_, dx = 1.0 / cshape[0], 2.0 / cshape[1]
    res = lambda V: orc.wave_residual(V, dt_, dx, g[tag + "_left_u"], g[tag + "_right_u"], g[tag + "_init_u"],
                                      g[tag + "_init_ut"], 1.0)
    F = res(U)
    assert relerr(F, g[tag + "_F"]) < 1e-12
    assert abs(np.mean(F ** 2) - g[tag + "_loss"]) < 1e-12 * g[tag + "_loss"]
    gU = orc.numerical_jacobian_T(res, U, F) * (2.0 / F.size)
    gr = orc.mg_adjoint(gU, [t.shape for t in terms], "cc") if len(terms) > 1 else [gU]
    for a, b in zip(gr, grads):
        assert relerr(a, b) < 1e-10


@pytest.mark.parametrize("cshape", [(10, 8, 6), (20, 16, 24)])
def test_wave2_oracle_matches_reference_golden(golden, cshape):
    """BASELINE configs[2]: the oracle's directly written (t, x, y) wave residual equals what the unmodified
    reference core.py computes for tests/operators.py::wave2_operator (tests/golden/make_goldens.py::gen_wave2)."""
    import argparse

    g = golden("wave2")
    tag = "w2_{}_f64".format("x".join(map(str, cshape)))
    U = g[tag + "_U"]
    nt, nx, ny = cshape
    # boundary data exactly as tests/operators.py::make_wave2 builds it (NumPy only)
    from tests import operators as ops

    t1 = (np.arange(nt) + 0.5) / nt
    x1 = -1 + (np.arange(nx) + 0.5) * 2 / nx
    y1 = -1 + (np.arange(ny) + 0.5) * 2 / ny
    T, X, Y = np.meshgrid(t1, x1, y1, indexing="ij")
    bnd = dict(xlo=ops.wave2_exact(T[:, 0, :], -1.0, Y[:, 0, :])[0], xhi=ops.wave2_exact(T[:, 0, :], 1.0, Y[:, 0, :])[0],
               ylo=ops.wave2_exact(T[:, :, 0], X[:, :, 0], -1.0)[0], yhi=ops.wave2_exact(T[:, :, 0], X[:, :, 0], 1.0)[0])
    u0, ut0 = ops.wave2_exact(0.0, X[0], Y[0])
    F = orc.wave2_residual(U, 1.0 / nt, 2.0 / nx, 2.0 / ny, bnd, u0, ut0, 1.0)
    assert np.max(np.abs(F - g[tag + "_F"])) < 1e-10 * np.max(np.abs(F))
    assert abs(np.mean(F ** 2) - g[tag + "_loss"]) < 1e-11 * g[tag + "_loss"]


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_adam_trajectory(golden, prec):
    g = golden("optim")
    dt = np.float64 if prec == "f64" else np.float32
    tag = f"adam_p2d_16_L3_{prec}"
    rhs = g[tag + "_rhs"]
    steps = [dt(1) / dt(16)] * 2
    offsets, table, rr = orc.poisson_plan(2, steps)
    x = [np.zeros(s, dt) for s in [(16, 16), (8, 8), (4, 4)]]
    m = [np.zeros_like(a) for a in x]
    v = [np.zeros_like(a) for a in x]
    losses = []
    for t in range(1, 21):
        loss, grads, _, _ = orc.eval_loss_grad_plan(x, "cc", offsets, table.astype(dt), rr, -rhs)
        losses.append(loss)
        for i in range(3):
            x[i], m[i], v[i] = orc.adam_step(x[i], m[i], v[i], grads[i].astype(dt), 0.005, t)
    tol = 1e-10 if prec == "f64" else 1e-4
    assert np.max(np.abs(np.array(losses) / g[tag + "_losses"] - 1)) < tol
    for i in range(3):
        assert relerr(x[i], g[f"{tag}_x{i}"]) < (1e-9 if prec == "f64" else 2e-3)


def test_gd(golden):
    g = golden("optim")
    tag = "gd_p1d_16_L3_f64"
    steps = [1.0 / 16]
    offsets, table, rr = orc.poisson_plan(1, steps)
    x = [g[f"{tag}_x0_{i}"] for i in range(3)]
    losses = []
    for _ in range(10):
        loss, grads, _, _ = orc.eval_loss_grad_plan(x, "c", offsets, table, rr, -g[tag + "_rhs"])
        losses.append(loss)
        x = [orc.gd_step(a, b, 1e-6) for a, b in zip(x, grads)]
    assert np.max(np.abs(np.array(losses) / g[tag + "_losses"] - 1)) < 1e-11
    for i in range(3):
        assert relerr(x[i], g[f"{tag}_x{i}"]) < 1e-12


def test_torch_port_matches_oracle():
    """The torch-CPU port timed by bench.py computes the same epoch as the NumPy oracle."""
    import torch

    from oracle import ref_port_torch as port

    for cshape, nlvl in [((16, 16), 3), ((8, 8, 8), 3)]:
        ep = port.PoissonAdamEpoch(cshape, nlvl, dtype=torch.float64, lr=0.005, seed=1)
        rng = np.random.default_rng(0)
        for a in ep.x:
            a.copy_(torch.from_numpy(rng.standard_normal(tuple(a.shape))))
        x0 = [a.numpy().copy() for a in ep.x]
        nd = len(cshape)
        offsets, table, rr = orc.poisson_plan(nd, [1.0 / n for n in cshape])
        loss_ref, grads_ref, _, _ = orc.eval_loss_grad_plan(x0, "c" * nd, offsets, table, rr, -ep.rhs.numpy())
        loss, grads = ep.loss_grad()
        assert abs(float(loss) - loss_ref) < 1e-12 * loss_ref
        for a, b in zip(grads, grads_ref):
            assert relerr(a.numpy(), b) < 1e-11
        ep.step()
        for i in range(nlvl):
            xr, _, _ = orc.adam_step(x0[i], np.zeros_like(x0[i]), np.zeros_like(x0[i]), grads_ref[i], 0.005, 1)
            assert relerr(ep.x[i].numpy(), xr) < 1e-11
