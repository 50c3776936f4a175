"""
GPU parity tests of the CUDA kernels, called through the C ABI (odil_b200.native -> libodil_b200.so),
against the CPU oracle and the golden vectors generated from the unmodified reference.
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import odil_oracle as orc
from tests import parity

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from odil_b200 import native
    native.load()

DT = {"f32": (np.float32, torch.float32), "f64": (np.float64, torch.float64)}
TOL = {"f32": 2e-5, "f64": 1e-12}


def dev(a, td=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t.to(td) if td is not None else t


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


LOCS = ["cccc", "nnnn", "cnnn", "nccc", "c.cn"]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("ndim", [1, 2, 3, 4])
@pytest.mark.parametrize("loc4", LOCS)
def test_interp_golden(golden, prec, ndim, loc4):
    nd, td = DT[prec]
    g = golden("transfers")
    loc = loc4[:ndim]
    u = g[f"interp_{ndim}_{loc}_in"].astype(nd)
    ref = g[f"interp_{ndim}_{loc}_out"]
    out = torch.empty(ref.shape, dtype=td, device="cuda")
    native.mg_interp_add(u.shape, loc, dev(u), 1.0, None, 0.0, out)
    assert relerr(out.cpu().numpy(), ref) < (100 * np.finfo(nd).eps)
    # fused synthesis step: out = 0.5*t + 2*I(u)
    t = np.random.default_rng(0).standard_normal(ref.shape).astype(nd)
    native.mg_interp_add(u.shape, loc, dev(u), 2.0, dev(t), 0.5, out)
    assert relerr(out.cpu().numpy(), 0.5 * t.astype(np.float64) + 2 * ref) < (100 * np.finfo(nd).eps)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("ndim", [1, 2, 3, 4])
@pytest.mark.parametrize("loc4", LOCS)
def test_interp_adjoint(golden, prec, ndim, loc4):
    nd, td = DT[prec]
    g = golden("transfers")
    loc = loc4[:ndim]
    cshape = g[f"interp_{ndim}_{loc}_in"].shape
    fshape = g[f"interp_{ndim}_{loc}_out"].shape
    w = np.random.default_rng(5).standard_normal(fshape).astype(nd)
    ref = orc.interp_adjoint(w.astype(np.float64), loc, cshape)
    out = torch.empty(cshape, dtype=td, device="cuda")
    native.mg_interp_adjoint(cshape, loc, dev(w), 1.0, out)
    assert relerr(out.cpu().numpy(), ref) < (200 * np.finfo(nd).eps)


def test_interp_adjoint_small_levels():
    """Coarse sizes 2 and 3 (deepest multigrid levels): every cell is a boundary cell."""
    for cshape in [(2,), (3,), (2, 2), (2, 3, 2), (3, 2, 2)]:
        loc = "c" * len(cshape)
        fshape = tuple(2 * n for n in cshape)
        w = np.random.default_rng(1).standard_normal(fshape)
        ref = orc.interp_adjoint(w, loc, cshape)
        out = torch.empty(cshape, dtype=torch.float64, device="cuda")
        native.mg_interp_adjoint(cshape, loc, dev(w), 1.0, out)
        assert relerr(out.cpu().numpy(), ref) < 1e-13
        u = np.random.default_rng(2).standard_normal(cshape)
        o2 = torch.empty(fshape, dtype=torch.float64, device="cuda")
        native.mg_interp_add(cshape, loc, dev(u), 1.0, None, 0.0, o2)
        assert relerr(o2.cpu().numpy(), orc.interp_to_finer(u, loc)) < 1e-13


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("ndim", [1, 2, 3, 4])
@pytest.mark.parametrize("loc4", LOCS[:4])
def test_restrict_golden(golden, prec, ndim, loc4):
    nd, td = DT[prec]
    g = golden("transfers")
    loc = loc4[:ndim]
    u = g[f"restrict_{ndim}_{loc}_in"].astype(nd)
    ref = g[f"restrict_{ndim}_{loc}_out"]
    out = torch.empty(ref.shape, dtype=td, device="cuda")
    native.mg_restrict(u.shape, loc, dev(u), out)
    assert relerr(out.cpu().numpy(), ref) < (100 * np.finfo(nd).eps)


def random_plan(rng, shape, offsets, rr):
    ncls = int(np.prod([2 * r + 1 for r in rr]))
    return rng.standard_normal((ncls, len(offsets)))


STENCIL_CASES = [
    # shape, offsets, rwidth
    ((37,), [(0,), (-1,), (1,), (2,)], (2,)),
    ((16, 12), [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)], (2, 1)),       # wave footprint
    ((9, 10), [(0, 0), (1, 1), (-1, 1)], (1, 1)),                            # non-star -> generic
    ((12, 12, 12), None, (1, 1, 1)),                                         # 3-D star -> tiled
    ((16, 8, 12), None, (1, 1, 1)),
    ((10, 7, 20), None, (1, 0, 1)),                                          # periodic axis 1
    ((24, 40), None, (1, 1)),                                                # 2-D star -> tiled
    ((33, 130), None, (2, 1)),                                               # 2-D star, scalar g path, r=2
    ((5, 6, 4, 7), None, (1, 1, 0, 1)),                                      # 4-D star -> generic
]


def star_offsets(ndim):
    offs = [(0,) * ndim]
    for a in range(ndim):
        for s in (-1, 1):
            offs.append(tuple(s if b == a else 0 for b in range(ndim)))
    return offs


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", range(len(STENCIL_CASES)))
def test_stencil_random_tables(prec, case):
    nd, td = DT[prec]
    shape, offsets, rr = STENCIL_CASES[case]
    offsets = offsets or star_offsets(len(shape))
    rng = np.random.default_rng(100 + case)
    table = random_plan(rng, shape, offsets, rr)
    tshape = tuple(2 * r + 1 for r in rr) + (len(offsets),)
    U = rng.standard_normal(shape).astype(nd)
    c = rng.standard_normal(shape).astype(nd)
    plan = native.StencilPlan(shape, td, offsets, rr, table)
    F_ref = orc.stencil_forward(U.astype(np.float64), offsets, table.reshape(tshape), rr, c.astype(np.float64))
    scale = 2.0 / U.size
    g_ref = orc.stencil_adjoint(F_ref, offsets, table.reshape(tshape), rr, scale)
    dU, dc = dev(U), dev(c)
    # forward
    F = torch.empty_like(dU)
    plan.forward(dU, dc, F)
    tol = TOL[prec] * 10
    assert relerr(F.cpu().numpy(), F_ref) < tol
    # adjoint of the device F
    G = torch.empty_like(dU)
    plan.adjoint(F, scale, None, G)
    assert relerr(G.cpu().numpy(), g_ref) < tol
    # fused (tiled when eligible)
    G2 = torch.full_like(dU, float("nan"))
    F2 = torch.full_like(dU, float("nan"))
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    plan.fused(dU, dc, scale, G2, ss, F_out=F2)
    assert relerr(F2.cpu().numpy(), F_ref) < tol
    assert relerr(G2.cpu().numpy(), g_ref) < tol
    assert abs(ss.item() - np.sum(F_ref ** 2)) < tol * np.sum(F_ref ** 2)
    # fused without c and without F_out
    plan.fused(dU, None, scale, G2, ss)
    F0 = orc.stencil_forward(U.astype(np.float64), offsets, table.reshape(tshape), rr, None)
    assert abs(ss.item() - np.sum(F0 ** 2)) < tol * np.sum(F0 ** 2)
    assert relerr(G2.cpu().numpy(), orc.stencil_adjoint(F0, offsets, table.reshape(tshape), rr, scale)) < tol


def test_star_plans_use_tiled_kernel():
    kinds = []
    for shape, offsets, rr in STENCIL_CASES:
        offsets = offsets or star_offsets(len(shape))
        ncls = int(np.prod([2 * r + 1 for r in rr]))
        plan = native.StencilPlan(shape, torch.float32, offsets, rr, np.ones((ncls, len(offsets))))
        kinds.append(plan.kind)
    assert kinds == [0, 0, 0, 1, 1, 1, 1, 1, 0]


def tune_or_skip(plan, zchunk, variant):
    """The superseded generations of the sweep (k_star_tma 0-3, k_star3d 10-13, k_star7 30-42) are only compiled with
    ODIL_B200_LEGACY=1; the default library rejects those variants."""
    try:
        plan.tune(zchunk=zchunk, variant=variant)
    except native.NativeError as e:
        if "ODIL_B200_LEGACY" in str(e):
            pytest.skip("library built without ODIL_B200_LEGACY")
        raise


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 10, 11, 12, 13, 20, 21, 22, 23])
@pytest.mark.parametrize("zchunk", [0, 1, 5, 64])
@pytest.mark.parametrize("kind", ["random", "dirichlet"])
def test_star_variants(variant, zchunk, kind):
    """Tile variants of the three fused kernels: 0-3 TMA-fed (needs a wrap-free table, e.g. the Dirichlet
    Poisson rows; random tables fall back to the LDG kernel), 10-13 tile + shell, 20-23 LDG column groups."""
    shape, rr = (20, 18, 36), (1, 1, 1)
    offsets = star_offsets(3)
    rng = np.random.default_rng(7)
    if kind == "random":
        table = rng.standard_normal((27, 7))
    else:
        offsets, table, rr = orc.poisson_plan(3, [0.3, 0.2, 0.1])
        table = table.reshape(27, 7) * (1 + 0.1 * rng.standard_normal((27, 7)))  # keeps the zero pattern
    U = rng.standard_normal(shape)
    c = rng.standard_normal(shape)
    plan = native.StencilPlan(shape, torch.float64, offsets, rr, table)
    tune_or_skip(plan, zchunk, variant)
    F_ref = orc.stencil_forward(U, offsets, table.reshape(3, 3, 3, 7), rr, c)
    g_ref = orc.stencil_adjoint(F_ref, offsets, table.reshape(3, 3, 3, 7), rr, 0.5)
    G = torch.empty(shape, dtype=torch.float64, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    plan.fused(dev(U), dev(c), 0.5, G, ss)
    assert relerr(G.cpu().numpy(), g_ref) < 1e-12
    assert abs(ss.item() - np.sum(F_ref ** 2)) < 1e-12 * np.sum(F_ref ** 2)


POISSON_CASES = {
    "p1d_16_L0": ((16,), 0), "p1d_16_L3": ((16,), 3), "p2d_16_L3": ((16, 16), 3), "p2d_12x8_L0": ((12, 8), 0),
    "p3d_8_L3": ((8, 8, 8), 3), "p3d_16x8x12_L2": ((16, 8, 12), 2), "p3d_12_L0": ((12, 12, 12), 0),
}


def device_eval_loss_grad(terms, loc, plan, c, td):
    """Multigrid synthesis -> fused stencil -> multigrid adjoint, all on the device."""
    L = len(terms)
    dterms = [dev(t) for t in terms]
    V = dterms[-1]
    for l in range(L - 2, -1, -1):
        out = torch.empty_like(dterms[l])
        native.mg_interp_add(dterms[l + 1].shape, loc, V, 1.0, dterms[l], 1.0, out)
        V = out
    U = V
    G = torch.empty_like(U)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    n = U.numel()
    plan.fused(U, c, 2.0 / n, G, ss)
    grads = [G]
    for l in range(1, L):
        gc = torch.empty_like(dterms[l])
        native.mg_interp_adjoint(dterms[l].shape, loc, grads[-1], 1.0, gc)
        grads.append(gc)
    return ss.item() / n, [x.cpu().numpy() for x in grads], U.cpu().numpy()


@pytest.mark.parametrize("name", list(POISSON_CASES))
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_poisson_golden(golden, name, prec):
    """Loss and gradient w.r.t. every multigrid term vs the reference (core.py + poisson.py under autograd)."""
    nd, td = DT[prec]
    g = golden("poisson")
    cshape, nlvl = POISSON_CASES[name]
    tag = f"{name}_{prec}"
    terms, grads = [], []
    i = 0
    while f"{tag}_term{i}" in g.files:
        terms.append(g[f"{tag}_term{i}"])
        grads.append(g[f"{tag}_grad{i}"])
        i += 1
    ndim = len(cshape)
    steps = [nd(1) / nd(n) for n in cshape]
    offsets, table, rr = orc.poisson_plan(ndim, steps)
    plan = native.StencilPlan(cshape, td, offsets, rr, table)
    loss, gr, U = device_eval_loss_grad(terms, "c" * ndim, plan, dev(-g[tag + "_rhs"]), td)
    f64 = prec == "f64"
    parity.check(f"kernels/poisson/{tag}/U", relerr(U, g[tag + "_U"]), 1e-13 if f64 else 2e-6)
    parity.check(f"kernels/poisson/{tag}/loss", abs(loss - g[tag + "_loss"]) / abs(g[tag + "_loss"]),
                 parity.F64_LOSS if f64 else parity.F32_LOSS)
    for i, (a, b) in enumerate(zip(gr, grads)):
        assert a.shape == b.shape
        parity.check(f"kernels/poisson/{tag}/grad{i}", relerr(a, b), parity.F64_GRAD if f64 else parity.F32_GRAD)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_adam_gd(prec):
    nd, td = DT[prec]
    rng = np.random.default_rng(3)
    shapes = [(33, 7), (1000,), (5,), (16, 16, 16)]
    x = [rng.standard_normal(s).astype(nd) for s in shapes]
    m = [np.zeros(s, nd) for s in shapes]
    v = [np.zeros(s, nd) for s in shapes]
    dx, dm, dv = [dev(a) for a in x], [dev(a) for a in m], [dev(a) for a in v]
    for t in range(1, 6):
        g = [(rng.standard_normal(s) * 10.0 ** rng.integers(-6, 2)).astype(nd) for s in shapes]
        alpha, omb1, omb2 = orc.adam_scalars(0.01, 0.9, 0.999, t, nd)
        native.adam_step(dx, dm, dv, [dev(a) for a in g], alpha, omb1, omb2, 1e-7)
        for i in range(len(x)):
            x[i], m[i], v[i] = orc.adam_step(x[i], m[i], v[i], g[i], 0.01, t)
    tol = 1e-13 if prec == "f64" else 3e-6
    for i in range(len(x)):
        assert relerr(dx[i].cpu().numpy(), x[i]) < tol
        assert relerr(dm[i].cpu().numpy(), m[i]) < tol
        assert relerr(dv[i].cpu().numpy(), v[i]) < tol
    g = [rng.standard_normal(s).astype(nd) for s in shapes]
    native.gd_step(dx, [dev(a) for a in g], 0.125)
    for i in range(len(x)):
        assert relerr(dx[i].cpu().numpy(), orc.gd_step(x[i], g[i], 0.125)) < tol


def test_adam_trajectory_golden(golden):
    """20 Adam epochs of 2-D Poisson 16^2 / 3 levels: loss trajectory vs the reference's own optimizer."""
    for prec in ["f64", "f32"]:
        nd, td = DT[prec]
        g = golden("optim")
        tag = f"adam_p2d_16_L3_{prec}"
        steps = [nd(1) / nd(16)] * 2
        offsets, table, rr = orc.poisson_plan(2, steps)
        plan = native.StencilPlan((16, 16), td, offsets, rr, table)
        c = dev(-g[tag + "_rhs"])
        shapes = [(16, 16), (8, 8), (4, 4)]
        x = [torch.zeros(s, dtype=td, device="cuda") for s in shapes]
        m = [torch.zeros_like(a) for a in x]
        v = [torch.zeros_like(a) for a in x]
        losses = []
        for t in range(1, 21):
            V = x[2]
            for l in (1, 0):
                out = torch.empty_like(x[l])
                native.mg_interp_add(shapes[l + 1], "cc", V, 1.0, x[l], 1.0, out)
                V = out
            G = torch.empty_like(V)
            ss = torch.zeros(1, dtype=torch.float64, device="cuda")
            plan.fused(V, c, 2.0 / 256, G, ss)
            grads = [G]
            for l in (1, 2):
                gc = torch.empty_like(x[l])
                native.mg_interp_adjoint(shapes[l], "cc", grads[-1], 1.0, gc)
                grads.append(gc)
            losses.append(ss.item() / 256)
            alpha, omb1, omb2 = orc.adam_scalars(0.005, 0.9, 0.999, t, nd)
            native.adam_step(x, m, v, grads, alpha, omb1, omb2, 1e-7)
        parity.check(f"kernels/adam20/{tag}/losses", np.max(np.abs(np.array(losses) / g[tag + "_losses"] - 1)),
                     1e-9 if prec == "f64" else 2e-6)
        for i in range(3):
            assert relerr(x[i].cpu().numpy(), g[f"{tag}_x{i}"]) < (1e-8 if prec == "f64" else 5e-6)


def test_reductions():
    rng = np.random.default_rng(9)
    for n in [1, 31, 4097, 1 << 20]:
        for prec in ["f32", "f64"]:
            nd, td = DT[prec]
            a = rng.standard_normal(n).astype(nd)
            b = rng.standard_normal(n).astype(nd)
            out = torch.zeros(1, dtype=torch.float64, device="cuda")
            native.sum_squares(dev(a), out)
            assert abs(out.item() - np.sum(a.astype(np.float64) ** 2)) < 1e-12 * n
            native.dot(dev(a), dev(b), out)
            assert abs(out.item() - np.dot(a.astype(np.float64), b.astype(np.float64))) < 1e-10 * n
            y = dev(b)
            native.axpby(2.0, dev(a), -0.5, y)
            assert relerr(y.cpu().numpy(), 2.0 * a - 0.5 * b) < 1e-6


# ------------------------------------------------------------------------------------------------
# Full-size property checks (BASELINE.json sizes): no CPU oracle at this size, so use identities.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,prec", [((512, 512, 512), "f32"), ((256, 256, 256), "f64"), ((1024, 1024), "f32")])
def test_fullsize_properties(shape, prec):
    nd, td = DT[prec]
    ndim = len(shape)
    steps = [nd(1) / nd(n) for n in shape]
    offsets, table, rr = orc.poisson_plan(ndim, steps)
    plan = native.StencilPlan(shape, td, offsets, rr, table)
    assert plan.kind == 1
    gen = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(shape, dtype=td, device="cuda", generator=gen)
    c = torch.randn(shape, dtype=td, device="cuda", generator=gen)
    n = U.numel()
    G = torch.empty_like(U)
    F = torch.empty_like(U)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    plan.fused(U, c, 2.0 / n, G, ss, F_out=F)
    # (1) fused F equals the unfused generic forward kernel; loss equals sum of squares of F
    F2 = torch.empty_like(U)
    plan.forward(U, c, F2)
    scaleF = F2.abs().max().item()
    assert (F - F2).abs().max().item() < (1e-5 if prec == "f32" else 1e-12) * scaleF
    s2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    native.sum_squares(F2, s2)
    assert abs(ss.item() - s2.item()) < (1e-6 if prec == "f32" else 1e-12) * s2.item()
    # (2) fused G equals the unfused generic adjoint kernel
    G2 = torch.empty_like(U)
    plan.adjoint(F2, 2.0 / n, None, G2)
    assert (G - G2).abs().max().item() < (2e-5 if prec == "f32" else 1e-12) * G2.abs().max().item()
    # (3) adjoint identity  <A W, F> == <W, A^T F>
    W = torch.randn(shape, dtype=td, device="cuda", generator=gen)
    AW = torch.empty_like(U)
    plan.forward(W, None, AW)
    d1 = torch.zeros(1, dtype=torch.float64, device="cuda")
    d2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    native.dot(AW, F2, d1)
    ATF = torch.empty_like(U)
    plan.adjoint(F2, 1.0, None, ATF)
    native.dot(W, ATF, d2)
    assert abs(d1.item() - d2.item()) < (1e-4 if prec == "f32" else 1e-11) * max(abs(d1.item()), 1.0)
    # (4) boundary rows: a field that is exactly the zero-Dirichlet quadratic in x along the last
    #     axis has the analytic second derivative in the first/last cell (poisson.py:57-68)
    del G, G2, AW, ATF, W


@pytest.mark.parametrize("cshape", [(256, 256, 256), (512, 512)])
def test_fullsize_mg_transpose(cshape):
    """<I u, w> == <u, I^T w> at full size, fp64."""
    ndim = len(cshape)
    loc = "c" * ndim
    fshape = tuple(2 * n for n in cshape)
    gen = torch.Generator(device="cuda").manual_seed(2)
    u = torch.randn(cshape, dtype=torch.float64, device="cuda", generator=gen)
    w = torch.randn(fshape, dtype=torch.float64, device="cuda", generator=gen)
    Iu = torch.empty(fshape, dtype=torch.float64, device="cuda")
    native.mg_interp_add(cshape, loc, u, 1.0, None, 0.0, Iu)
    Itw = torch.empty(cshape, dtype=torch.float64, device="cuda")
    native.mg_interp_adjoint(cshape, loc, w, 1.0, Itw)
    d1 = torch.zeros(1, dtype=torch.float64, device="cuda")
    d2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    native.dot(Iu, w, d1)
    native.dot(u, Itw, d2)
    assert abs(d1.item() - d2.item()) < 1e-10 * max(abs(d1.item()), 1.0)
    # linear functions are reproduced exactly by I (reference tests/test_mg_interp.py)
    xs = [torch.arange(n, dtype=torch.float64, device="cuda") + 0.5 for n in cshape]
    lin = sum((i + 1.0) * x.reshape([-1 if a == i else 1 for a in range(ndim)]) / cshape[i] for i, x in enumerate(xs))
    lin = lin.expand(cshape).contiguous()
    native.mg_interp_add(cshape, loc, lin, 1.0, None, 0.0, Iu)
    xf = [torch.arange(2 * n, dtype=torch.float64, device="cuda") + 0.5 for n in cshape]
    linf = sum((i + 1.0) * x.reshape([-1 if a == i else 1 for a in range(ndim)]) / (2 * cshape[i])
               for i, x in enumerate(xf))
    assert (Iu - linf).abs().max().item() < 1e-12


# --------------------------------------------------------------------------------------------------
# k_star8 / k_star7 (the hot kernels) on awkward shapes, every tile variant, x-uniform and per-cell arms
# --------------------------------------------------------------------------------------------------
STAR8_SHAPES = [(20, 18, 36), (7, 33, 260), (5, 15, 8), (45, 132), (64, 16, 128)]


@pytest.mark.parametrize("variant", [-1, 50, 51, 52, 60, 61, 62, 30, 31, 32, 40, 42])
@pytest.mark.parametrize("zchunk", [0, 1, 5])
@pytest.mark.parametrize("shape", STAR8_SHAPES)
def test_star8_variants(variant, zchunk, shape):
    """Default dispatch (-1) and every tile variant of k_star8 (50-52 x-uniform arms, 60-62 per-cell arms) and k_star7
    (30-32, 40-42): g, sum F^2 and F against the oracle, with and without the constant term."""
    nd = len(shape)
    rng = np.random.default_rng(11)
    offsets, table, rr = orc.poisson_plan(nd, [0.3, 0.2, 0.1][:nd])
    ncls = 3 ** nd
    table = np.asarray(table).reshape(ncls, 2 * nd + 1)
    if variant >= 60 or (40 <= variant < 50):
        table = table * (1 + 0.1 * rng.standard_normal(table.shape))  # keeps the zero pattern; arms now depend on x
    tshape = (3,) * nd + (2 * nd + 1,)
    for prec in ("f64", "f32"):
        npd, td = DT[prec]
        U = rng.standard_normal(shape).astype(npd)
        c = rng.standard_normal(shape).astype(npd)
        plan = native.StencilPlan(shape, td, offsets, rr, table)
        assert plan.kind == 1
        if variant >= 0:
            tune_or_skip(plan, zchunk, variant)
        elif zchunk:
            plan.tune(zchunk=zchunk, variant=-1)
        F_ref = orc.stencil_forward(U.astype(np.float64), offsets, table.reshape(tshape), rr, c.astype(np.float64))
        g_ref = orc.stencil_adjoint(F_ref, offsets, table.reshape(tshape), rr, 0.5)
        tol = TOL[prec] * 10
        G = torch.full(shape, float("nan"), dtype=td, device="cuda")
        F = torch.full(shape, float("nan"), dtype=td, device="cuda")
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dev(U), dev(c), 0.5, G, ss)
        assert relerr(G.cpu().numpy(), g_ref) < tol
        assert abs(ss.item() - np.sum(F_ref ** 2)) < tol * np.sum(F_ref ** 2)
        G.fill_(float("nan"))
        plan.fused(dev(U), dev(c), 0.5, G, ss, F_out=F)
        assert relerr(F.cpu().numpy(), F_ref) < tol
        assert relerr(G.cpu().numpy(), g_ref) < tol
        F0 = orc.stencil_forward(U.astype(np.float64), offsets, table.reshape(tshape), rr, None)
        plan.fused(dev(U), None, 0.5, G, ss)
        assert abs(ss.item() - np.sum(F0 ** 2)) < tol * np.sum(F0 ** 2)
        assert relerr(G.cpu().numpy(), orc.stencil_adjoint(F0, offsets, table.reshape(tshape), rr, 0.5)) < tol


# --------------------------------------------------------------------------------------------------
# marching multigrid transfers: odd shapes, plane ranges (what a slab passes), against the oracle
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cshape", [(4, 4, 4), (5, 7, 6), (9, 4, 34), (33, 17, 64), (6, 66, 130)])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_marching_transfers(cshape, prec):
    npd, td = DT[prec]
    rng = np.random.default_rng(3)
    fshape = tuple(2 * s for s in cshape)
    coarse = rng.standard_normal(cshape).astype(npd)
    term = rng.standard_normal(fshape).astype(npd)
    ref = 1.3 * term.astype(np.float64) + 0.7 * orc.interp_to_finer(coarse.astype(np.float64), "ccc")
    tol = 100 * np.finfo(npd).eps
    out = torch.full(fshape, float("nan"), dtype=td, device="cuda")
    native.mg_interp_add(cshape, "ccc", dev(coarse), 0.7, dev(term), 1.3, out)
    assert relerr(out.cpu().numpy(), ref) < tol
    native.mg_interp_add(cshape, "ccc", dev(coarse), 0.7, None, 0.0, out)
    assert relerr(out.cpu().numpy(), 0.7 * orc.interp_to_finer(coarse.astype(np.float64), "ccc")) < tol
    g_ref = 0.9 * orc.interp_adjoint(term.astype(np.float64), "ccc", cshape)
    gc = torch.full(cshape, float("nan"), dtype=td, device="cuda")
    native.mg_interp_adjoint(cshape, "ccc", dev(term), 0.9, gc)
    assert relerr(gc.cpu().numpy(), g_ref) < tol
    # plane ranges with shifted array origins, as the slab decomposition passes them
    n0 = cshape[0]
    if n0 >= 5:
        for (cb, ce) in [(0, 2), (1, n0 - 1), (n0 - 2, n0), (2, 3)]:
            fb, fe = 2 * cb, 2 * ce
            # interp: fine planes [fb, fe) from coarse planes cb-1 .. ce (clipped), arrays start at those planes
            c_lo, c_hi = max(cb - 1, 0), min(ce + 1, n0)
            o = torch.full((fe - fb,) + fshape[1:], float("nan"), dtype=td, device="cuda")
            native.mg_interp_add(cshape, "ccc", dev(coarse[c_lo:c_hi]), 0.7, dev(term[fb:fe]), 1.3, o, rng=(fb, fe, fb, c_lo))
            assert relerr(o.cpu().numpy(), ref[fb:fe]) < tol
            # transpose: coarse planes [cb, ce) from fine planes 2cb-2 .. 2ce+1 (clipped; the outermost two carry the
            # pad corrections of the coarse planes 1 and n0-2)
            f_lo, f_hi = max(2 * cb - 2, 0), min(2 * ce + 2, fshape[0])
            gg = torch.full((ce - cb,) + tuple(cshape[1:]), float("nan"), dtype=td, device="cuda")
            native.mg_interp_adjoint(cshape, "ccc", dev(term[f_lo:f_hi]), 0.9, gg, rng=(cb, ce, cb, f_lo))
            assert relerr(gg.cpu().numpy(), g_ref[cb:ce]) < tol


@pytest.mark.parametrize("cshape", [(4, 4, 4), (5, 7, 6), (9, 4, 34), (33, 17, 64), (6, 66, 130)])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_adam_synth_equals_adam_then_interp(cshape, prec):
    """odil_b200_adam_synth against odil_b200_adam_step followed by odil_b200_mg_interp_add: x, m, v and the
    synthesised field bit for bit (interior, faces, edges and corners of the joint pad)."""
    npd, td = DT[prec]
    rng = np.random.default_rng(5)
    fshape = tuple(2 * s for s in cshape)
    coarse = dev(rng.standard_normal(cshape).astype(npd))
    mk = lambda: dev(rng.standard_normal(fshape).astype(npd))
    x0, m0, g = mk(), mk(), mk()
    v0 = dev(rng.random(fshape).astype(npd))
    alpha, omb1, omb2, eps = 0.0123, 0.1, 0.001, 1e-7
    xa, ma, va = x0.clone(), m0.clone(), v0.clone()
    native.adam_step([xa], [ma], [va], [g], alpha, omb1, omb2, eps)
    ua = torch.full(fshape, float("nan"), dtype=td, device="cuda")
    native.mg_interp_add(cshape, "ccc", coarse, 0.7, xa, 1.3, ua)
    xb, mb, vb = x0.clone(), m0.clone(), v0.clone()
    ub = torch.full(fshape, float("nan"), dtype=td, device="cuda")
    assert native.adam_synth(cshape, "ccc", coarse, 0.7, 1.3, xb, mb, vb, g, ub, alpha, omb1, omb2, eps)
    for a, b in ((xa, xb), (ma, mb), (va, vb), (ua, ub)):
        assert torch.equal(a, b)
    # step size from device memory (graph replay)
    xc, mc, vc = x0.clone(), m0.clone(), v0.clone()
    uc = torch.full(fshape, float("nan"), dtype=td, device="cuda")
    alpha_dev = torch.tensor([float(npd(alpha))], dtype=torch.float64, device="cuda")
    assert native.adam_synth(cshape, "ccc", coarse, 0.7, 1.3, xc, mc, vc, g, uc, 0.0, omb1, omb2, eps, alpha_dev=alpha_dev)
    xd, md, vd = x0.clone(), m0.clone(), v0.clone()
    native.adam_step_dev([xd], [md], [vd], [g], alpha_dev, omb1, omb2, eps)
    assert torch.equal(xc, xd) and torch.equal(mc, md) and torch.equal(vc, vd)
    # unsupported geometry: nothing is touched
    assert not native.adam_synth((8, 8), "cc", dev(np.zeros((8, 8), npd)), 1.0, 1.0, *(dev(np.zeros((16, 16), npd)) for _ in range(5)),
                                 alpha, omb1, omb2, eps)


@pytest.mark.parametrize("cshape", [(4, 4, 4), (5, 7, 6), (9, 4, 34), (33, 17, 64), (6, 66, 130), (40, 72, 192)])
def test_adjoint_tma_equals_ldg(cshape, monkeypatch):
    """k_interp_adjoint3t (TMA-staged fine planes, the fp32 default) against k_interp_adjoint3m (LDG, ODIL_B200_ADJ_TMA=0):
    the same operations per coarse cell, so the results are bit-identical -- whole arrays and slab-style plane ranges."""
    rng = np.random.default_rng(11)
    fshape = tuple(2 * s for s in cshape)
    term = rng.standard_normal(fshape).astype(np.float32)
    n0 = cshape[0]
    ranges = [None] + ([(0, 2), (1, n0 - 1), (n0 - 2, n0), (2, 3)] if n0 >= 5 else [])
    for r in ranges:
        outs = []
        for flag in ("1", "0"):
            monkeypatch.setenv("ODIL_B200_ADJ_TMA", flag)
            if r is None:
                gc = torch.full(cshape, float("nan"), dtype=torch.float32, device="cuda")
                native.mg_interp_adjoint(cshape, "ccc", dev(term), 0.9, gc)
            else:
                cb, ce = r
                f_lo, f_hi = max(2 * cb - 2, 0), min(2 * ce + 2, fshape[0])
                gc = torch.full((ce - cb,) + tuple(cshape[1:]), float("nan"), dtype=torch.float32, device="cuda")
                native.mg_interp_adjoint(cshape, "ccc", dev(term[f_lo:f_hi]), 0.9, gc, rng=(cb, ce, cb, f_lo))
            outs.append(gc.cpu().numpy())
        assert np.array_equal(outs[0], outs[1]), (cshape, r, np.max(np.abs(outs[0] - outs[1])))
    if ranges[0] is None:
        g_ref = 0.9 * orc.interp_adjoint(term.astype(np.float64), "ccc", cshape)
        monkeypatch.setenv("ODIL_B200_ADJ_TMA", "1")
        gc = torch.full(cshape, float("nan"), dtype=torch.float32, device="cuda")
        native.mg_interp_adjoint(cshape, "ccc", dev(term), 0.9, gc)
        assert relerr(gc.cpu().numpy(), g_ref) < 100 * np.finfo(np.float32).eps


# --------------------------------------------------------------------------------------------------
# k_tile2d: any offset set of small radius on 2-D grids (the wave example's footprint), all three modes
# --------------------------------------------------------------------------------------------------
TILE2D_CASES = [
    # shape, offsets (None = star), rwidth
    ((70, 150), [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)], (2, 1)),        # wave footprint, 3 x 3 tiles
    ((32, 64), None, (1, 1)),                                                  # exactly one tile
    ((33, 65), None, (1, 1)),                                                  # one cell over in both axes
    ((3, 5), [(0, 0), (2, -2), (-1, 1)], (1, 2)),                              # grid smaller than the halo
    ((40, 66), [(0, 0), (3, 0), (0, -4), (-2, 2)], (0, 0)),                    # radius 3 / 4, fully periodic
    ((96, 200), [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)], (1, 1)),  # 9-point box
    ((1, 300), [(0, 0), (0, -1), (0, 1)], (0, 1)),                             # single row
    ((130, 7), [(0, 0), (1, 0), (-1, 0), (0, 1)], (3, 1)),                     # narrow, wide regions
]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", range(len(TILE2D_CASES)))
def test_tile2d_matches_oracle_and_generic(prec, case):
    nd, td = DT[prec]
    shape, offsets, rr = TILE2D_CASES[case]
    offsets = offsets or star_offsets(2)
    rng = np.random.default_rng(500 + case)
    tshape = tuple(2 * r + 1 for r in rr) + (len(offsets),)
    table = rng.standard_normal(tshape)
    U = rng.standard_normal(shape).astype(nd)
    c = rng.standard_normal(shape).astype(nd)
    Gin = rng.standard_normal(shape).astype(nd)
    scale = 2.0 / U.size
    F_ref = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, c.astype(np.float64))
    g_ref = orc.stencil_adjoint(F_ref, offsets, table, rr, scale)
    tol = TOL[prec] * 10
    res = {}
    for variant in (70, 71):  # 70: k_tile2d, 71: the per-cell generic kernel
        plan = native.StencilPlan(shape, td, offsets, rr, table.reshape(-1, len(offsets)))
        plan.tune(variant=variant)
        dU, dc = dev(U), dev(c)
        F = torch.full_like(dU, float("nan"))
        plan.forward(dU, dc, F)
        G = torch.full_like(dU, float("nan"))
        plan.adjoint(dev(F_ref.astype(nd)), scale, dev(Gin), G)
        G2 = torch.full_like(dU, float("nan"))
        F2 = torch.full_like(dU, float("nan"))
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, dc, scale, G2, ss, F_out=F2)
        G3 = torch.full_like(dU, float("nan"))
        ss0 = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, None, scale, G3, ss0)
        res[variant] = [t.cpu().numpy() for t in (F, G, G2, F2, ss, G3, ss0)]
        F, G, G2, F2, ss, G3, ss0 = res[variant]
        assert relerr(F, F_ref) < tol and relerr(F2, F_ref) < tol
        assert relerr(G, g_ref + Gin.astype(np.float64)) < tol
        assert relerr(G2, g_ref) < tol
        assert abs(ss[0] - np.sum(F_ref ** 2)) < tol * np.sum(F_ref ** 2)
        F0 = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, None)
        assert abs(ss0[0] - np.sum(F0 ** 2)) < tol * np.sum(F0 ** 2)
        assert relerr(G3, orc.stencil_adjoint(F0, offsets, table, rr, scale)) < tol
    # same summation order in both kernels: the arrays agree to rounding of the fused multiply-adds
    eps = np.finfo(nd).eps
    for a, b in zip(res[70][:4], res[71][:4]):
        assert relerr(a, b) < 8 * eps


def test_tile2d_large_wave_footprint_properties():
    """2048 x 4096 fp32 (BASELINE configs[2] scale): linearity of the fused sweep in (U, c) and the adjoint
    identity <A U, F> = <U, A^T F> through the forward / adjoint modes."""
    shape, rr = (2048, 4096), (2, 1)
    offsets = [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)]
    rng = np.random.default_rng(9)
    table = rng.standard_normal((15, 5))
    plan = native.StencilPlan(shape, torch.float64, offsets, rr, table)
    gen = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen)
    V = torch.randn(shape, dtype=torch.float64, device="cuda", generator=gen)
    AU, ATV = torch.empty_like(U), torch.empty_like(U)
    plan.forward(U, None, AU)
    plan.adjoint(V, 1.0, None, ATV)
    lhs, rhs = (AU * V).sum().item(), (U * ATV).sum().item()
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0) + 1e-6
    G, ss = torch.empty_like(U), torch.zeros(1, dtype=torch.float64, device="cuda")
    plan.fused(U, None, 0.5, G, ss)
    assert abs(ss.item() - (AU * AU).sum().item()) < 1e-10 * ss.item()
    ATAU = torch.empty_like(U)
    plan.adjoint(AU, 0.5, None, ATAU)
    assert (G - ATAU).abs().max().item() < 1e-9 * ATAU.abs().max().item()


# --------------------------------------------------------------------------------------------------
# k_tile2w: the warp-private marching version of the fused 2-D sweep (wrap-free plans, <= 8 offsets, radii <= 2)
# --------------------------------------------------------------------------------------------------
TILE2W_CASES = [
    # shape, offsets (None = star), rwidth
    ((70, 152), [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)], (2, 1)),        # wave footprint (wave.py:38-46), 2 strips
    ((32, 64), None, (1, 1)),
    ((33, 68), None, (1, 1)),
    ((5, 4), [(0, 0), (0, 1), (1, 0)], (1, 1)),                                # one lane of cells
    ((96, 200), [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dy, dx) != (1, 1)], (1, 1)),  # 8 offsets
    ((1, 300), [(0, 0), (0, -1), (0, 1)], (0, 1)),                             # single row
    ((130, 8), [(0, 0), (1, 0), (-1, 0), (0, 1)], (3, 1)),                     # narrow, wide regions
    ((257, 1024), None, (1, 1)),                                               # 9 strips x 33 chunks
    ((64, 244), [(0, 0), (2, -2), (-1, 1), (1, 2), (-2, 0), (0, -1)], (2, 2)), # radius 2 on both axes
    ((40, 120), [(0, 0), (1, 0), (0, -2), (-2, 2), (2, 1), (-1, -1), (0, 1), (1, -2)], (3, 2)),
]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", range(len(TILE2W_CASES)))
def test_tile2w_matches_oracle_and_tile2d(prec, case, monkeypatch):
    """k_tile2w (ODIL_B200_TILE2W=2: refuse to fall back) against the oracle and against k_tile2d: same operations per
    cell in the same order, so F and g are bit-identical; the loss sums differ by their summation order only."""
    from tests.test_tile_emulation_cpu import wrap_free_table

    nd, td = DT[prec]
    shape, offsets, rr = TILE2W_CASES[case]
    offsets = offsets or star_offsets(2)
    rng = np.random.default_rng(700 + case)
    table = wrap_free_table(rng.standard_normal(tuple(2 * r + 1 for r in rr) + (len(offsets),)), offsets, rr)
    U = rng.standard_normal(shape).astype(nd)
    c = rng.standard_normal(shape).astype(nd)
    scale = 2.0 / U.size
    F_ref = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, c.astype(np.float64))
    g_ref = orc.stencil_adjoint(F_ref, offsets, table, rr, scale)
    F0 = orc.stencil_forward(U.astype(np.float64), offsets, table, rr, None)
    g0 = orc.stencil_adjoint(F0, offsets, table, rr, scale)
    tol = TOL[prec] * 10
    res = {}
    for flag in ("0", "2"):
        monkeypatch.setenv("ODIL_B200_TILE2W", flag)
        plan = native.StencilPlan(shape, td, offsets, rr, table.reshape(-1, len(offsets)))
        dU, dc = dev(U), dev(c)
        G = torch.full_like(dU, float("nan"))
        F = torch.full_like(dU, float("nan"))
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, dc, scale, G, ss, F_out=F)
        G1 = torch.full_like(dU, float("nan"))
        ss1 = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, dc, scale, G1, ss1)
        G3 = torch.full_like(dU, float("nan"))
        ss0 = torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(dU, None, scale, G3, ss0)
        res[flag] = [t.cpu().numpy() for t in (F, G, G1, G3, ss, ss1, ss0)]
        F, G, G1, G3, ss, ss1, ss0 = res[flag]
        assert relerr(F, F_ref) < tol and relerr(G, g_ref) < tol and relerr(G3, g0) < tol
        assert np.array_equal(G, G1)
        assert abs(ss[0] - np.sum(F_ref ** 2)) < tol * np.sum(F_ref ** 2) and ss[0] == ss1[0]
        assert abs(ss0[0] - np.sum(F0 ** 2)) < tol * np.sum(F0 ** 2)
    for a, b in zip(res["0"][:4], res["2"][:4]):
        assert np.array_equal(a, b)


def test_tile2w_refuses_what_it_cannot_do(monkeypatch):
    """ODIL_B200_TILE2W=2 turns the silent choice of k_tile2d (periodic plan, odd row length, 9 offsets) into an error;
    ODIL_B200_TILE2W=1 falls back."""
    rng = np.random.default_rng(3)
    offsets, rr = star_offsets(2), (1, 1)
    table = rng.standard_normal((9, 5))  # not wrap-free: the boundary classes couple across the periodic boundary
    for shape, tab in (((16, 16), table), ((16, 18), None)):
        if tab is None:
            from tests.test_tile_emulation_cpu import wrap_free_table

            tab = wrap_free_table(table.reshape(3, 3, 5), offsets, rr).reshape(9, 5)
        U = torch.randn(shape, dtype=torch.float32, device="cuda")
        G, ss = torch.empty_like(U), torch.zeros(1, dtype=torch.float64, device="cuda")
        monkeypatch.setenv("ODIL_B200_TILE2W", "2")
        plan = native.StencilPlan(shape, torch.float32, offsets, rr, tab)
        with pytest.raises(native.NativeError):
            plan.fused(U, None, 1.0, G, ss)
        monkeypatch.setenv("ODIL_B200_TILE2W", "1")
        plan.fused(U, None, 1.0, G, ss)
        F = orc.stencil_forward(U.cpu().numpy().astype(np.float64), offsets, tab.reshape(3, 3, 5), rr, None)
        assert relerr(G.cpu().numpy(), orc.stencil_adjoint(F, offsets, tab.reshape(3, 3, 5), rr, 1.0)) < 1e-5


def test_tile2w_large_wave_footprint_equals_tile2d(monkeypatch):
    """2048 x 4096 fp32 with the wave footprint: k_tile2w and k_tile2d agree bit for bit on g, to rounding on the loss."""
    from tests.test_tile_emulation_cpu import wrap_free_table

    shape, rr = (2048, 4096), (2, 1)
    offsets = [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)]
    rng = np.random.default_rng(9)
    table = wrap_free_table(rng.standard_normal((5, 3, 5)), offsets, rr).reshape(15, 5)
    gen = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(shape, dtype=torch.float32, device="cuda", generator=gen)
    c = torch.randn(shape, dtype=torch.float32, device="cuda", generator=gen)
    out = {}
    for flag in ("0", "2"):
        monkeypatch.setenv("ODIL_B200_TILE2W", flag)
        plan = native.StencilPlan(shape, torch.float32, offsets, rr, table)
        G, ss = torch.full_like(U, float("nan")), torch.zeros(1, dtype=torch.float64, device="cuda")
        plan.fused(U, c, 0.5, G, ss)
        out[flag] = (G, ss.item())
    assert torch.equal(out["0"][0], out["2"][0])
    assert abs(out["0"][1] - out["2"][1]) < 1e-6 * out["0"][1]


# --------------------------------------------------------------------------------------------------
# 2-D cell-centred transfers through shared-memory tiles (k_interp_add2t / k_interp_adjoint2t)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cshape", [(2, 2), (3, 4), (16, 64), (17, 66), (5, 130), (40, 2), (33, 128), (64, 200),
                                    (3, 3)])  # the last one (odd width) stays on the per-cell kernels
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_tile2d_transfers(cshape, prec):
    npd, td = DT[prec]
    rng = np.random.default_rng(11)
    fshape = tuple(2 * s for s in cshape)
    coarse = rng.standard_normal(cshape).astype(npd)
    term = rng.standard_normal(fshape).astype(npd)
    I = orc.interp_to_finer(coarse.astype(np.float64), "cc")
    tol = 100 * np.finfo(npd).eps
    out = torch.full(fshape, float("nan"), dtype=td, device="cuda")
    native.mg_interp_add(cshape, "cc", dev(coarse), 0.7, dev(term), 1.3, out)
    assert relerr(out.cpu().numpy(), 1.3 * term.astype(np.float64) + 0.7 * I) < tol
    native.mg_interp_add(cshape, "cc", dev(coarse), 1.0, None, 0.0, out)
    assert relerr(out.cpu().numpy(), I) < tol
    gc = torch.full(cshape, float("nan"), dtype=td, device="cuda")
    native.mg_interp_adjoint(cshape, "cc", dev(term), 0.9, gc)
    assert relerr(gc.cpu().numpy(), 0.9 * orc.interp_adjoint(term.astype(np.float64), "cc", cshape)) < tol


def test_tile2d_transfers_transpose_identity_large():
    """1024^2 -> 2048^2 (configs[1] scale and above): <I u, v> = <u, I^T v>, and I reproduces linear functions."""
    cshape, fshape = (1024, 1024), (2048, 2048)
    gen = torch.Generator(device="cuda").manual_seed(2)
    u = torch.randn(cshape, dtype=torch.float64, device="cuda", generator=gen)
    v = torch.randn(fshape, dtype=torch.float64, device="cuda", generator=gen)
    Iu, ITv = torch.empty_like(v), torch.empty_like(u)
    native.mg_interp_add(cshape, "cc", u, 1.0, None, 0.0, Iu)
    native.mg_interp_adjoint(cshape, "cc", v, 1.0, ITv)
    lhs, rhs = (Iu * v).sum().item(), (u * ITv).sum().item()
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0) + 1e-7
    yc = (torch.arange(1024, dtype=torch.float64, device="cuda") + 0.5) / 1024
    yf = (torch.arange(2048, dtype=torch.float64, device="cuda") + 0.5) / 2048
    lin = (2 * yc[:, None] - 3 * yc[None, :] + 1).contiguous()
    native.mg_interp_add(cshape, "cc", lin, 1.0, None, 0.0, Iu)
    assert (Iu - (2 * yf[:, None] - 3 * yf[None, :] + 1)).abs().max().item() < 1e-12
