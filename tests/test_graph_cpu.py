"""
CPU tests of the general (non-affine) path: tracer -> expression graph -> generated program, executed through the host
twin of the generated kernels (tests/jit_host.py) and compared with goldens produced by the UNMODIFIED reference
(tests/golden/make_nonaffine_goldens.py).  Also: every generated source compiles with NVRTC for sm_100a (no device
needed for compilation).  The GPU suite runs the same cases through the real kernels (tests/test_graph_gpu.py).
"""
import numpy as np
import pytest
import torch

import odil
from odil_b200 import native
from tests import nonaffine_cases as cases
from tests import refsrc
from tests.jit_host import HostTwin

pytestmark = pytest.mark.skipif(not refsrc.available(), reason="reference scripts absent (oracle/_ref not staged)")


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def build_case(case, prec, device="cpu"):
    dt = np.float64 if prec == "f64" else np.float32
    mod = odil.backend.ModB200(device=device)
    scripts = {name: refsrc.load(cases.SCRIPTS[name]) for name in cases.scripts_for(case)}
    operator, domain, state, extra, tracers = cases.build(case, odil, mod, dt, scripts)
    problem = odil.Problem(operator, domain, extra, tracers=dict(tracers))
    return problem, state, dt


def golden_arrays(g, key, prefix):
    out, i = [], 0
    while f"{key}_{prefix}{i}" in g.files:
        out.append(g[f"{key}_{prefix}{i}"])
        i += 1
    return out


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", cases.CASES)
def test_host_twin_matches_reference_golden(golden, case, prec):
    g = golden("nonaffine")
    key = f"{case}_{prec}"
    problem, state, dt = build_case(case, prec)
    tw = HostTwin(problem, state)
    assert tw.engine.names == [str(n) for n in g[key + "_names"]]
    assert [o.raw for o in tw.engine.outputs] == [bool(r) for r in g[key + "_raws"]]
    x = [torch.as_tensor(a) for a in golden_arrays(g, key, "x")]
    assert [tuple(a.shape) for a in x] == [tuple(a.shape) for a in problem.domain.arrays_from_state(state)]
    loss, grads, terms = tw.loss_grad(x)
    tol = 1e-11 if prec == "f64" else 2e-5
    assert abs(loss - float(g[key + "_loss"])) < tol * abs(float(g[key + "_loss"]))
    assert relerr(terms, g[key + "_terms"]) < tol
    for gi, gr in zip(grads, golden_arrays(g, key, "g")):
        assert relerr(gi, gr) < tol, (case, relerr(gi, gr))
    for F, Fr in zip(tw.values(x), golden_arrays(g, key, "F")):
        assert F.shape == Fr.shape and relerr(F, Fr) < tol


@pytest.mark.parametrize("case", cases.NEWTON_CASES)
def test_host_twin_jacobian_matches_reference_golden(golden, case):
    """`jac` rows, forward-mode (J v) and reverse-mode (J^T w) programs against the dense Jacobian of the reference
    evaluation (torch.autograd.functional.jacobian over the unmodified reference operator)."""
    g = golden("nonaffine")
    key = f"{case}_f64"
    problem, state, dt = build_case(case, "f64")
    tw = HostTwin(problem, state)
    x = [torch.as_tensor(a) for a in golden_arrays(g, key, "x")]
    J, row0, col0 = tw.jacobian_dense(x)
    Jr = g[key + "_jac"]
    assert J.shape == Jr.shape
    assert np.max(np.abs(J - Jr)) < 1e-11 * np.max(np.abs(Jr))
    rng = np.random.default_rng(0)
    v, w = rng.standard_normal(J.shape[1]), rng.standard_normal(J.shape[0])
    assert relerr(tw.jvp(x, v), Jr @ v) < 1e-11
    assert relerr(tw.vjp(x, w), Jr.T @ w) < 1e-11
    # the same products from stored per-cell diagonals (modes jacd / jvpd / vjpd: what a Newton step's CG uses)
    if tw.gen.dia_ok():
        dia = tw.diagonals(x)
        assert relerr(tw.jvpd(x, v, dia), Jr @ v) < 1e-11
        assert relerr(tw.vjpd(x, w, dia), Jr.T @ w) < 1e-11
        if tw.gen.gather_ok():  # every load a pure roll: J^T w gathered per input cell
            assert relerr(tw.vjpg(x, w, dia), Jr.T @ w) < 1e-11
        else:
            assert case != "heat3"
    else:
        assert case in ("newton", "infer_constant", "heat_k")  # network weights / Array elements among the unknowns


@pytest.mark.parametrize("case", cases.CASES)
def test_generated_sources_compile_for_sm100a(case):
    """NVRTC turns every generated translation unit into an sm_100a cubin (compilation needs no device)."""
    problem, state, dt = build_case(case, "f32")
    tw = HostTwin(problem, state)
    modes = ["lossgrad", "values"] + (["jvp", "vjp", "jac"] if case in cases.NEWTON_CASES else [])
    if case in cases.NEWTON_CASES and tw.gen.dia_ok():
        modes += ["jacd", "jvpd", "vjpd"] + (["vjpg"] if tw.gen.gather_ok() else [])
    for mode in modes:
        m = native.JitModule(tw.engine.source(mode))
        assert len(m.cubin()) > 1000
        assert "error" not in m.log.lower()


@pytest.mark.parametrize("case", cases.CASES)
def test_stored_diagonals_reproduce_the_forward_and_reverse_products(golden, case):
    """For every example operator whose unknowns allow it: the Jacobian products from the stored per-cell diagonals
    (modes jacd + jvpd / vjpd, and the gathered vjpg where every load is a pure roll) equal the forward- / reverse-mode
    products (jvp / vjp) at the golden state -- pads, trims, strided slices (mgloss), Raw terms and where() included."""
    g = golden("nonaffine")
    key = f"{case}_f64"
    problem, state, dt = build_case(case, "f64")
    tw = HostTwin(problem, state)
    if not tw.gen.dia_ok():
        pytest.skip("cell-independent loads among the unknowns (network weights / Array elements)")
    x = [torch.as_tensor(a) for a in golden_arrays(g, key, "x")]
    if any(u.kind == "MultigridField" and u.narrays > 1 for u in tw.engine.unknowns.values()):
        pytest.skip("Newton needs multigrid off")
    ncol = sum(a.numel() for a in x)
    nrow = sum(o.n for o in tw.engine.outputs)
    rng = np.random.default_rng(1)
    v, w = rng.standard_normal(ncol), rng.standard_normal(nrow)
    dia = tw.diagonals(x)
    jv, jtw = tw.jvp(x, v), tw.vjp(x, w)
    assert relerr(tw.jvpd(x, v, dia), jv) < 1e-12
    assert relerr(tw.vjpd(x, w, dia), jtw) < 1e-12
    if tw.gen.gather_ok():
        assert relerr(tw.vjpg(x, w, dia), jtw) < 1e-12
