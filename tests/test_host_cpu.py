"""
CPU tests (no GPU): the tracing front-end and its lowering to stencil plans, the host-side API mirror,
and the C-ABI library surface.  The plans are checked by interpreting them with the CPU oracle.
"""
import argparse
import os
import sys
import re

import numpy as np
import pytest
import torch

import odil
from odil_b200.engine import ResidualEngine
from oracle import odil_oracle as orc
from tests import operators as ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan_apply(spec, U, const):
    shape = spec["shape"]
    tshape = tuple(2 * r + 1 for r in spec["rwidth"]) + (len(spec["offsets"]),)
    return orc.stencil_forward(U, [tuple(o) for o in spec["offsets"]], spec["table"].reshape(tshape),
                               spec["rwidth"], const)


@pytest.mark.parametrize("cshape", [(16,), (12, 8), (8, 6, 10), (4, 6, 4, 6)])
def test_trace_poisson_matches_direct_operator(cshape):
    problem, state = ops.make_poisson(cshape)
    eng = ResidualEngine(problem, state, trace_only=True)
    assert len(eng.outputs) == 1 and eng.outputs[0].fused
    spec = eng.outputs[0].blocks[0].spec
    assert spec["rwidth"] == (1,) * len(cshape)
    assert len(spec["offsets"]) == 2 * len(cshape) + 1
    rng = np.random.default_rng(0)
    U = rng.standard_normal(cshape)
    F = plan_apply(spec, U, eng.outputs[0].const.cpu().numpy())
    rhs = np.asarray(problem.extra.rhs)
    F_ref = orc.poisson_residual(U, rhs, [1.0 / n for n in cshape])
    assert np.max(np.abs(F - F_ref)) < 1e-9 * np.max(np.abs(F_ref))
    # the table is exactly the one written down in SURVEY.md Appendix A
    offsets, table, rr = orc.poisson_plan(len(cshape), [1.0 / n for n in cshape])
    mine = {tuple(o): spec["table"][:, i] for i, o in enumerate(spec["offsets"])}
    ref = {tuple(o): table.reshape(-1, len(offsets))[:, i] for i, o in enumerate(offsets)}
    assert set(mine) == set(ref)
    for k in ref:
        assert np.max(np.abs(mine[k] - ref[k])) < 1e-9 * np.max(np.abs(ref[k]))


def test_trace_poisson_rhs_matches_golden(golden):
    """discrete rhs computed through ModB200 (host side) equals the reference's (poisson.py:71-86)."""
    problem, state = ops.make_poisson((16, 16))
    g = golden("poisson")
    assert np.max(np.abs(np.asarray(problem.extra.rhs) - g["p2d_16_L3_f64_rhs"])) < 1e-11


@pytest.mark.parametrize("cshape", [(10, 8, 6), (7, 12, 9)])
def test_trace_wave2_matches_direct_operator(cshape):
    """BASELINE configs[2] footprint: the wave operator in two space dimensions lowers to 7 offsets
    (t, t-1, t-2 and the four space neighbours at t-1) x 45 classes (rows it == 0 and it == 1 distinct)."""
    problem, state = ops.make_wave2(cshape)
    eng = ResidualEngine(problem, state, trace_only=True)
    out = eng.outputs[0]
    assert eng.names == ["fu"] and out.fused
    spec = out.blocks[0].spec
    assert spec["rwidth"] == (2, 1, 1) and spec["table"].shape[0] == 45
    assert sorted(map(tuple, spec["offsets"])) == sorted(
        [(0, 0, 0), (-1, 0, 0), (-2, 0, 0), (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1)])
    e = problem.extra
    U = np.random.default_rng(2).standard_normal(cshape)
    F = plan_apply(spec, U, out.const.cpu().numpy())
    nt, nx, ny = cshape
    F_ref = orc.wave2_residual(U, 1.0 / nt, 2.0 / nx, 2.0 / ny, dict(xlo=e.xlo, xhi=e.xhi, ylo=e.ylo, yhi=e.yhi),
                               e.init_u, e.init_ut, 1.0)
    assert np.max(np.abs(F - F_ref)) < 1e-10 * np.max(np.abs(F_ref))
    # the discrete equations are consistent: the exact solution leaves a residual that shrinks with the grid
    res = []
    for n in (12, 24):
        p2, _ = ops.make_wave2((n, n, n))
        x = p2.extra
        R = orc.wave2_residual(x.ref_u, 1.0 / n, 2.0 / n, 2.0 / n, dict(xlo=x.xlo, xhi=x.xhi, ylo=x.ylo, yhi=x.yhi),
                               x.init_u, x.init_ut, 1.0)
        res.append(np.sqrt(np.mean(R[2:] ** 2)))
    assert res[1] < 0.5 * res[0]


@pytest.mark.parametrize("cshape,nlvl", [((16, 12), 0), ((16, 8), 2)])
def test_trace_wave_matches_direct_operator(golden, cshape, nlvl):
    problem, state = ops.make_wave(cshape, nlvl)
    eng = ResidualEngine(problem, state, trace_only=True)
    out = eng.outputs[0]
    assert eng.names == ["fu"] and out.fused
    spec = out.blocks[0].spec
    assert spec["rwidth"] == (2, 1)  # rows it==0 and it==1 differ from the interior (wave.py:63,71)
    assert sorted(map(tuple, spec["offsets"])) == sorted([(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)])
    rng = np.random.default_rng(1)
    U = rng.standard_normal(cshape)
    e = problem.extra
    F = plan_apply(spec, U, out.const.cpu().numpy())
    F_ref = orc.wave_residual(U, 1.0 / cshape[0], 2.0 / cshape[1], e.left_u, e.right_u, e.init_u, e.init_ut, 1.0)
    assert np.max(np.abs(F - F_ref)) < 1e-10 * np.max(np.abs(F_ref))
    # and against the reference itself
    g = golden("wave")
    tag = ("w_16x12_L0" if nlvl == 0 else "w_16x8_L2") + "_f64"
    terms = []
    i = 0
    while f"{tag}_term{i}" in g.files:
        terms.append(g[f"{tag}_term{i}"])
        i += 1
    Ug = orc.mg_synthesize(terms, "cc") if len(terms) > 1 else terms[0]
    Fg = plan_apply(spec, Ug, out.const.cpu().numpy())
    assert np.max(np.abs(Fg - g[tag + "_F"])) < 1e-10 * np.max(np.abs(g[tag + "_F"]))


def test_trace_multiple_fields_and_array():
    """Operator in the style of reference tests/test_optimize.py: several fields at c/n locations + Array."""
    domain = odil.Domain(cshape=(8, 4), dimnames=["x", "y"], lower=(0, 0), upper=(2, 1), multigrid=True, mg_nlvl=2,
                         dtype=np.float64)
    ref = {}
    for key, loc in [("uc", "cc"), ("un", "nn"), ("ufx", "nc"), ("ufy", "cn")]:
        x, y = domain.points(loc=loc)
        ref[key] = np.asarray(x) * 0.25 + np.asarray(y) * 0.5
    ref["a"] = np.arange(5, dtype=np.float64)

    def operator(ctx):
        res = [(key, ctx.field(key) - ctx.extra[key]) for key in ["uc", "un", "ufx", "ufy"]]
        res += [("a", ctx.field("a") - ctx.extra["a"])]
        return res

    state = odil.State(fields={
        "uc": odil.Field(np.zeros(domain.size(loc="cc")), loc="cc"),
        "un": odil.Field(np.zeros(domain.size(loc="nn")), loc="nn"),
        "ufx": odil.Field(np.zeros(domain.size(loc="nc")), loc="nc"),
        "ufy": odil.Field(np.zeros(domain.size(loc="cn")), loc="cn"),
        "a": odil.Array(np.zeros(5)),
    })
    state = domain.init_state(state)
    assert isinstance(state.fields["un"], odil.MultigridField)  # mg_convert_all
    assert [tuple(t.array.shape) for t in state.fields["ufx"].terms] == [(9, 4), (5, 2)]
    eng = ResidualEngine(odil.Problem(operator, domain, ref), state, trace_only=True)
    assert eng.names == ["uc", "un", "ufx", "ufy", "a"]
    assert all(o.fused for o in eng.outputs)
    for o, key in zip(eng.outputs, eng.names):
        spec = o.blocks[0].spec
        assert spec["rwidth"] == (0,) * len(o.shape) and spec["table"].tolist() == [[1.0]]
        assert np.allclose(o.const.cpu().numpy(), -ref[key])
    assert eng.narrays == 2 * 4 + 1


def test_nonaffine_operators_fail_loudly():
    problem, state = ops.make_poisson((32, 8))

    def op_square(ctx):
        u = ctx.field("u")
        return [u * u]

    def op_mask(ctx):
        u = ctx.field("u")
        return [ctx.mod.where(u > 0, u, 0 * u)]

    def op_varcoef(ctx):
        x, y = ctx.points()
        return [ctx.field("u", 1, 0) * x - ctx.field("u")]

    for op in (op_square, op_mask, op_varcoef):
        p = odil.Problem(op, problem.domain, problem.extra)
        with pytest.raises(odil.NonAffineError):
            ResidualEngine(p, state, trace_only=True)


def test_roll_and_stop_gradient_of_expressions():
    problem, state = ops.make_poisson((8, 6))

    def operator(ctx):
        mod = ctx.mod
        u = ctx.field("u")
        d = (mod.roll(u, -1, 0) - u) * 3.0          # forward difference via roll of an expression
        d2 = mod.roll(d, [1, 0], [0, 1])            # shift back: (u - roll(u, 1))*3
        return [d2 + mod.stop_gradient(ctx.field("u", 0, 1))]

    eng = ResidualEngine(odil.Problem(operator, problem.domain, problem.extra), state, trace_only=True)
    out = eng.outputs[0]
    blocks = {b.frozen: b.spec for b in out.blocks}
    live = {tuple(o): blocks[False]["table"][0, i] for i, o in enumerate(blocks[False]["offsets"])}
    assert live == {(0, 0): 3.0, (-1, 0): -3.0}
    assert [tuple(o) for o in blocks[True]["offsets"]] == [(0, 1)]
    assert not out.fused


def test_domain_geometry_and_state_roundtrip():
    domain = odil.Domain(cshape=(4, 6), dimnames=["x", "y"], lower=(0, -1), upper=(2, 1), dtype=np.float64,
                         multigrid=True, mg_convert_all=False)
    x, y = domain.points()
    assert x.shape == (4, 6) and np.allclose(np.asarray(x)[:, 0], [0.25, 0.75, 1.25, 1.75])
    xn = domain.points("x", loc="nn")
    assert xn.shape == (5, 7) and np.allclose(np.asarray(xn)[:, 0], [0, 0.5, 1, 1.5, 2])
    ix, iy = domain.indices()
    assert np.array_equal(np.asarray(iy)[0], np.arange(6))
    assert domain.size() == [4, 6] and domain.size("y", loc="cn") == 7
    assert np.allclose(domain.step(), (0.5, 1 / 3))
    assert domain.mg_cshapes == [(4, 6), (2, 3)]
    state = odil.State(fields={
        "field": np.random.rand(4, 6),
        "mgfield": domain.regular_to_multigrid(np.random.rand(4, 6)),
        "net": domain.make_neural_net([3, 3]),
        "array": [1, 2, 3],
    })
    state = domain.init_state(state)
    arrays = domain.arrays_from_state(state)
    assert [tuple(a.shape) for a in arrays] == [(4, 6), (4, 6), (2, 3), (3, 3), (3,), (3,)]
    packed = domain.pack_state(state)
    assert packed.shape == (24 + 24 + 6 + 9 + 3 + 3,)
    domain.unpack_state(packed + 1, state)
    after = domain.arrays_from_state(state)
    for a, b in zip(arrays, after):
        assert torch.equal(a + 1, b)
    with pytest.raises(ValueError):
        odil.Domain(cshape=(6, 6), multigrid=True, mg_nlvl=3)  # 6 -> 3 -> 1 is not a halving chain


def test_adam_scalars_follow_dtype():
    from odil_b200.optimizer import adam_scalars

    for dt in (np.float32, np.float64):
        for t in (1, 2, 17, 1000):
            a, o1, o2 = adam_scalars(0.005, 0.9, 0.999, t, dt)
            ra, ro1, ro2 = orc.adam_scalars(0.005, 0.9, 0.999, t, dt)
            assert (a, o1, o2) == (float(ra), float(ro1), float(ro2))
    assert adam_scalars(0.1, 0.9, 0.999, 1, np.float32)[1] == float(np.float32(1) - np.float32(0.9))


def test_history_csv(tmp_path):
    h = odil.History(csvpath=str(tmp_path / "train.csv"), warmup=1)
    for e in range(3):
        h.append("epoch", e)
        h.append("loss", np.float32(1.0 / (e + 1)))
        if e > 0:
            h.append("late", 2.0 * e)
        h.write()
    lines = open(tmp_path / "train.csv").read().strip().split("\n")
    assert lines[0] == "epoch,loss,late" and len(lines) == 4 and lines[1].startswith("0,1.0,0.0")


def test_cabi_exports_match_header():
    """Every function declared in include/odil_b200.h is exported by the built library."""
    from odil_b200 import build, native

    lib_path = build.build()
    header = open(os.path.join(ROOT, "include", "odil_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(odil_b200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 18
    import ctypes

    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(native.EXPORTS) == declared
    assert lib.odil_b200_version() == 100


def test_engine_refuses_cpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    problem, state = ops.make_poisson((8, 8))
    with pytest.raises(odil.native.NativeError):
        problem.eval_loss_grad(state)


def test_lbfgs_control_flow_matches_scipy(monkeypatch):
    """The device L-BFGS driver (compact-form direction + dcsrch transcription) makes the same step-length
    decisions as scipy.optimize.fmin_l_bfgs_b.  The four vector kernels are replaced by torch-CPU stand-ins
    here; the kernels themselves are checked on the GPU (tests/test_kernels_gpu.py)."""
    from scipy import optimize

    from odil_b200 import lbfgs, native

    monkeypatch.setattr(native, "multi_dot", lambda V, k, g, out: out.__setitem__(slice(0, k), V[:k] @ g))
    monkeypatch.setattr(native, "multi_axpy", lambda V, k, coef, a0, g, d: d.copy_(a0 * g + coef[:k] @ V[:k]))
    monkeypatch.setattr(native, "dot", lambda a, b, out: out.__setitem__(0, torch.dot(a, b)))
    monkeypatch.setattr(native, "axpby", lambda a, x, b, y: y.copy_(a * x + (b * y if b != 0 else 0)))
    rng = np.random.default_rng(0)
    n = 40
    A = rng.standard_normal((n, n))
    A = A @ A.T + np.eye(n) * 0.1
    b = rng.standard_normal(n)

    def fnp(x):
        r = A @ x - b
        return 0.5 * r @ r + 0.1 * np.sum(x ** 4), A.T @ r + 0.4 * x ** 3

    for m, iters in [(7, 60), (50, 40)]:
        ref = []
        _, _, info = optimize.fmin_l_bfgs_b(fnp, np.zeros(n), maxiter=iters, pgtol=1e-16, m=m, maxls=50, factr=0,
                                            callback=lambda x: ref.append(fnp(x)[0]))
        mine = []
        x, f, inf = lbfgs.minimize(lambda x: (float(fnp(x.numpy())[0]), torch.from_numpy(fnp(x.numpy())[1])),
                                   torch.zeros(n, dtype=torch.float64), m=m, maxiter=iters, maxls=50, pgtol=1e-16,
                                   factr=0, callback=lambda x: mine.append(fnp(x.numpy())[0]))
        assert inf["nit"] == info["nit"] and inf["funcalls"] == info["funcalls"]
        assert np.max(np.abs(np.array(mine) / np.array(ref) - 1)) < 1e-7


@pytest.mark.parametrize("cshape", [(9,), (6, 5), (4, 5, 6)])
def test_newton_jacobian_assembly_matches_oracle(cshape):
    """StencilJacobian.tocsr() / diagonals() (host assembly of the Newton matrix) equal the Jacobian of
    the oracle's stencil application; no GPU needed (trace-only engine)."""
    from odil_b200.newton import StencilJacobian

    problem, state = ops.make_poisson(cshape)
    eng = ResidualEngine(problem, state, trace_only=True)
    jac = StencilJacobian(eng)
    n = int(np.prod(cshape))
    assert jac.shape == (n, n)
    A = jac.tocsr().toarray()
    spec = eng.outputs[0].blocks[0].spec
    f0 = plan_apply(spec, np.zeros(cshape), None).reshape(-1)
    J = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1
        J[:, j] = plan_apply(spec, e.reshape(cshape), None).reshape(-1) - f0
    assert np.max(np.abs(A - J)) < 1e-12 * np.max(np.abs(J))
    diag = jac.diagonals()[0]
    assert all(tuple(v.shape) == tuple(cshape) for v in diag.values())
    assert {k[1] for k in diag} == {tuple(int(x) for x in o) for o in spec["offsets"]}


@pytest.mark.parametrize("dtype,shape,variant,zchunk", [
    (np.float32, (512, 512, 512), -1, 0), (np.float32, (64, 40, 136), -1, 0), (np.float32, (1, 1024, 1024), -1, 0),
    (np.float64, (256, 256, 512), -1, 0), (np.float32, (100, 45, 132), 51, 7), (np.float32, (512, 512, 512), 52, 64),
])
def test_star_worklist_tiles_the_slab_once(dtype, shape, variant, zchunk):
    """Launch plan of the fused star sweep (host logic of k_star8): the CTAs tile the slab exactly once, respect the
    rows-per-CTA limit, and the headline grid fills the 148 SMs in one balanced wave."""
    from odil_b200 import native

    n0, N1, N2 = shape
    work = native.star_worklist(dtype, n0, N1, N2, variant=variant, zchunk=zchunk)
    vw = 4 if dtype == np.float32 else 2
    tx = 32 * vw
    nr_max = {-1: 14, 51: 12, 52: 8}[variant]
    cover = np.zeros((n0, N1, (N2 + tx - 1) // tx), dtype=np.int32)
    for x0, y0, rows, zs, ze in work:
        assert x0 % tx == 0 and 0 <= x0 < N2
        assert 1 <= rows <= nr_max and 0 <= y0 and y0 + rows <= N1
        assert 0 <= zs < ze <= n0
        cover[zs:ze, y0:y0 + rows, x0 // tx] += 1
    assert cover.min() == 1 and cover.max() == 1
    if shape == (512, 512, 512) and variant == -1:
        assert len(work) == 148                       # one CTA per SM, one wave
        assert set(work[:, 2]) <= {13, 14}            # 512 rows x 4 column blocks over 148 CTAs
        assert np.all(work[:, 3] == 0) and np.all(work[:, 4] == 512)
    if zchunk:
        assert np.max(work[:, 4] - work[:, 3]) <= zchunk


def test_cli_flags_match_the_reference():
    """Every flag of the reference's `util.add_arguments` / `linsolver.add_arguments` exists here with the same
    type, default and choices (tests/golden/io/cli_flags.json is produced by executing those two functions from the
    reference sources); the only addition is the `cg_b200` linear solver."""
    import argparse
    import json

    from tests.golden.make_io_goldens import describe_flags

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "io", "cli_flags.json")))
    for module, add in (("util", odil.util.add_arguments), ("linsolver", odil.linsolver.add_arguments)):
        parser = argparse.ArgumentParser()
        add(parser)
        mine = json.loads(json.dumps(describe_flags(parser)))
        assert sorted(mine) == sorted(gold[module])
        for flag, (typ, default, choices) in gold[module].items():
            if flag == "linsolver":
                assert mine[flag][:2] == [typ, default] and mine[flag][2] == choices + ["cg_b200"]
            else:
                assert mine[flag] == [typ, default, choices], flag


def test_epoch_callback_schedule_log_and_history(tmp_path, monkeypatch):
    """make_callback: which epochs report / record / plot / checkpoint, the log lines downstream tools parse, the
    train.csv columns, the hooks' arguments, and that time spent in hooks is not counted as optimizer time."""
    import argparse
    import time

    monkeypatch.chdir(tmp_path)
    log = open(tmp_path / "train.log", "w")
    odil.set_log_file(log, echo=0)
    problem = argparse.Namespace(domain=argparse.Namespace(cshape=(8, 4)), tracers={})
    args = argparse.Namespace(report_every=2, history_every=2, history_full=2, plot_every=4, frames=1,
                              checkpoint_every=4, linsolver_history=1)
    seen = {"report": [], "history": [], "plot": [], "checkpoint": [], "epoch": []}

    def epoch_func(p, state, epoch, cb):
        seen["epoch"].append(epoch)
        time.sleep(0.02)

    def report_func(p, state, epoch, cb):
        seen["report"].append((epoch, cb.task_report, cb.pinfo["loss"], cb.args is args, cb.problem is problem))

    def history_func(p, state, epoch, history, cb):
        seen["history"].append(epoch)
        history.append("extra", 2.0 * epoch)
        assert history is cb.history

    def plot_func(p, state, epoch, frame, cb):
        seen["plot"].append((epoch, frame, cb.frame))

    def checkpoint_func(p, state, epoch, cb):
        seen["checkpoint"].append(epoch)

    cb = odil.make_callback(problem, args, epoch_func=epoch_func, report_func=report_func, history_func=history_func,
                            checkpoint_func=checkpoint_func, plot_func=plot_func)
    assert cb.cbinfo.history is not None and cb.cbinfo.frame == 0
    for epoch in range(0, 7):
        pinfo = {"norms": [np.float64(0.5) ** epoch, 3.0], "names": ["fu", ""], "loss": 1.0 / (epoch + 1),
                 "linsolver": {"niter": 3, "residual": 0.25, "skipped": [1, 2]}}
        cb("state", epoch, pinfo)
    odil.set_log_file(sys.stderr)
    log.close()
    assert seen["epoch"] == list(range(7)) and problem.tracers["epoch"] == 6
    assert [r[0] for r in seen["report"]] == [0, 2, 4, 6] and all(r[1] and r[3] and r[4] for r in seen["report"])
    assert seen["history"] == [0, 1, 2, 4, 6]            # every 2nd epoch, and every epoch below history_full
    assert seen["plot"] == [(0, 0, 0), (4, 1, 1)] and cb.frame == 2
    assert seen["checkpoint"] == [0, 4]
    text = open(tmp_path / "train.log").read()
    assert "\nepoch=00004\nresidual: fu:0.0625, 1:3\n" in text
    assert text.count("throughput: ") == 4 and "walltime/epoch: " in text and "gpu_pool: " in text
    assert "throughput: 0.000 Mcells/s" in text.split("epoch=00002")[0]   # first report has no interval yet
    rows = open(tmp_path / "train.csv").read().strip().split("\n")
    assert rows[0] == "epoch,frame,norm_fu,norm_1,loss,lin_niter,lin_residual,walltime,memory,gpu_used,gpu_pool,extra"
    assert [r.split(",")[0] for r in rows[1:]] == ["0", "1", "2", "4", "6"]
    assert rows[3].split(",")[1:7] == ["1", "0.25", "3.0", "0.3333333333333333", "3", "0.25"]
    # 7 x 20 ms were spent in the epoch hook: they are callback time, not optimizer time
    assert cb.time_callback >= 0.14 and cb.walltime < cb.time_callback
    assert cb.throughput > 0 and cb.epoch == 6


def test_epoch_callback_default_checkpoint_and_no_history(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    saved = []
    monkeypatch.setattr(odil.core, "checkpoint_save", lambda domain, state, path: saved.append(path))
    problem = argparse.Namespace(domain=argparse.Namespace(cshape=(4,)), tracers=None)
    args = argparse.Namespace(report_every=0, history_every=0, history_full=0, plot_every=5, frames=0,
                              checkpoint_every=3)
    cb = odil.make_callback(problem, args)
    assert cb.history is None
    for epoch in range(7):
        cb(None, epoch, None)
    assert saved == ["checkpoint_000000.pickle", "checkpoint_000003.pickle", "checkpoint_000006.pickle"]
    assert cb.frame == 1  # epoch 0 is not a frame when frames == 0; epoch 5 is
    assert not os.path.exists(tmp_path / "train.csv")


def test_optimize_grad_glue(monkeypatch):
    """util.optimize_grad around a stand-in optimizer: the initial state is reported as epoch `epoch_start` with its
    own loss, flags map to optimizer keywords, epochs count from `epoch_start`, the state object follows the
    optimizer's arrays, and `callback_update_state` hands arrays modified by the callback back to the optimizer."""
    from odil_b200 import util

    class Domain:
        dtype = np.float64
        mod = "mod"

        def arrays_from_state(self, state):
            return list(state["arrays"])

        def arrays_to_state(self, arrays, state):
            state["arrays"] = list(arrays)

    class Problem:
        domain = Domain()
        evals = 0

        def eval_loss_grad(self, state):
            Problem.evals += 1
            x = state["arrays"][0]
            return float(np.sum(x ** 2)), [2 * x], [float(np.sum(x ** 2))], ["f"], [float(np.sqrt(np.sum(x ** 2)))]

    made = {}

    class FakeOpt:
        displayname = "Fake"

        def run(self, x0, loss_grad, epochs, callback, epoch_start, lr, **kw):
            made["run"] = dict(epochs=epochs, epoch_start=epoch_start, lr=lr, kw=kw)
            x = list(x0)
            for epoch in range(epoch_start + 1, epoch_start + epochs + 1):
                loss, grads, pinfo = loss_grad(x)
                x[0] = x[0] - lr * grads[0]
                if callback:
                    callback(x, epoch, pinfo)
            return x, argparse.Namespace(epochs=epochs, evals=epochs)

    def fake_make(name, dtype=None, mod=None, **kw):
        made["make"] = dict(name=name, dtype=dtype, mod=mod, kw=kw)
        return FakeOpt()

    monkeypatch.setattr(util, "make_optimizer", fake_make)
    monkeypatch.setattr(util, "printlog", lambda *a: None)
    seen = []

    def callback(state, epoch, pinfo):
        seen.append((epoch, pinfo["loss"], float(state["arrays"][0][0])))
        if epoch == 12:
            state["arrays"] = [np.array([10.0])]  # the callback resets the unknown

    args = argparse.Namespace(epochs=14, epoch_start=10, lr=0.25, callback_update_state=1, bfgs_m=7, bfgs_pgtol=None,
                              bfgs_maxls=None, adam_epsilon=1e-3, adam_beta_1=None, adam_beta_2=None)
    state = {"arrays": [np.array([8.0])]}
    arrays, info = util.optimize(args, "fake", Problem(), state, callback)
    assert made["make"] == dict(name="fake", dtype=np.float64, mod="mod", kw={"m": 7, "epsilon": 1e-3})
    assert made["run"] == dict(epochs=4, epoch_start=10, lr=0.25, kw={"m": 7, "epsilon": 1e-3})
    # epoch 10 = initial state; x halves every epoch (x - 0.25 * 2x); the callback sees the updated x and the loss
    # evaluated before the update; after epoch 12 the optimizer continues from the callback's 10.0
    assert seen == [(10, 64.0, 8.0), (11, 64.0, 4.0), (12, 16.0, 2.0), (13, 100.0, 5.0), (14, 25.0, 2.5)]
    assert float(arrays[0][0]) == 2.5 and float(state["arrays"][0][0]) == 2.5 and info.epochs == 4
    # without a callback the engine is still warmed up by one evaluation before the optimizer runs
    Problem.evals = 0
    args.callback_update_state = 0
    util.optimize_grad(args, "fake", Problem(), {"arrays": [np.array([1.0])]}, None)
    assert Problem.evals == 1 + 4


def test_optimize_newton_glue(monkeypatch):
    """util.optimize_newton around a stand-in problem and solver: state += solve(J, -F) per epoch, the callback sees
    the initial state first and the solver's status afterwards."""
    from odil_b200 import linsolver, util

    class Domain:
        def pack_state(self, state):
            return state["x"].copy()

        def unpack_state(self, packed, state):
            state["x"] = packed

        def arrays_from_state(self, state):
            return [state["x"]]

    class Problem:
        domain = Domain()

        def linearize(self, state):  # F(x) = 3 x - 6
            return 3 * state["x"] - 6, 3.0

        def eval_loss_grad(self, state):
            f = 3 * state["x"] - 6
            return float(f @ f), None, [float(f @ f)], ["f"], [float(np.sqrt(f @ f))]

    def fake_solve(matrix, rhs, args, status, name):
        status.update(niter=1, solver=name)
        return rhs / matrix

    monkeypatch.setattr(linsolver, "solve", fake_solve)
    monkeypatch.setattr(util, "printlog", lambda *a: None)
    seen = []
    args = argparse.Namespace(epochs=2, epoch_start=0, linsolver="direct", linsolver_verbose=1)
    state = {"x": np.array([5.0, -1.0])}
    arrays, info = util.optimize(args, "newton", Problem(), state,
                                 lambda st, epoch, pinfo: seen.append((epoch, pinfo["loss"], pinfo.get("linsolver"))))
    assert seen == [(0, 162.0, None), (1, 0.0, {"niter": 1, "solver": "direct"}),
                    (2, 0.0, {"niter": 1, "solver": "direct"})]
    assert np.array_equal(arrays[0], [2.0, 2.0]) and info.epochs == 2


# --------------------------------------------------------------------------------------------------
# Round-2 advisor findings
# --------------------------------------------------------------------------------------------------
def test_stop_gradient_merges_frozen_and_live_terms():
    """mod.stop_gradient(u + ctx.field(k, frozen=True)) must ADD the two coefficients of the same (key, offset)."""
    from odil_b200.backend import Affine

    domain = odil.Domain(cshape=(8,), dtype=np.float64, multigrid=False, mod=odil.backend.ModB200(device="cpu"))

    u = Affine.symbol("u", (0,), (8,), np.float64, frozen=False, device="cpu")
    uf = Affine.symbol("u", (0,), (8,), np.float64, frozen=True, device="cpu")
    e = domain.mod.stop_gradient(u + uf)
    assert list(e.lin) == [("u", (0,), True)]
    assert float(e.lin[("u", (0,), True)].dense((1,), torch.float64, "cpu")[0]) == 2.0


def test_lower_block_folds_offsets_on_size_one_axes():
    """roll() along a size-1 axis is the identity: offsets there fold to 0 and merge (the kernels wrap an index
    with a single conditional add, so an offset of 2 on a size-1 axis must never reach them)."""
    from odil_b200.backend import Coef
    from odil_b200.engine import lower_block

    one = torch.ones((1, 1), dtype=torch.float64)
    offsets, rr, table = lower_block((1, 8), {(2, 0): Coef([one]), (0, 0): Coef([3 * one]), (0, 1): Coef([one]),
                                              (0, 9): Coef([one])})
    assert sorted(map(tuple, offsets.tolist())) == [(0, 0), (0, 1)]
    mine = {tuple(o): table[:, i] for i, o in enumerate(offsets.tolist())}
    assert np.all(mine[(0, 0)] == 4.0) and np.all(mine[(0, 1)] == 2.0)


def test_tracer_reads_are_recorded_and_invalidate_the_lowering():
    """An affine operator whose weight depends on tracers['epoch'] is re-lowered when the tracer changes
    (the reference passes tracers as run-time arguments of the jitted function, core.py:1076-1110)."""
    def op(ctx):
        k = 0.5 ** (ctx.tracers["epoch"] / 10)
        return [(ctx.field("u") - 1.0) * k]

    domain = odil.Domain(cshape=(8,), dtype=np.float64, multigrid=False, mod=odil.backend.ModB200(device="cpu"))
    state = odil.State(fields={"u": odil.Field(torch.zeros(8, dtype=torch.float64))}, initialized=True)
    problem = odil.Problem(op, domain, tracers={"epoch": 0, "unused": 7})
    eng = ResidualEngine(problem, state, trace_only=True)
    assert eng.tracer_view.reads == {"epoch": 0}
    assert not eng.tracer_view.stale({"epoch": 0, "unused": 8})
    assert eng.tracer_view.stale({"epoch": 10, "unused": 7})
    problem.tracers["epoch"] = 10
    eng2 = ResidualEngine(problem, state, trace_only=True)
    t0, t1 = eng.outputs[0].blocks[0].spec["table"], eng2.outputs[0].blocks[0].spec["table"]
    assert np.allclose(t0, 1.0) and np.allclose(t1, 0.5)


def test_native_handles_are_not_destroyed_during_a_graph_capture(monkeypatch):
    """A destructor that runs while a CUDA graph is being captured (Python's cyclic collector can run one at any
    allocation) must not free device memory: `native._release` parks the handle and `flush_deferred` destroys it
    once the capture has ended."""
    from odil_b200 import native

    destroyed = []
    state = {"capturing": True}
    monkeypatch.setattr(native, "_capturing", lambda: state["capturing"])
    monkeypatch.setattr(native, "_deferred", [])
    native._release(destroyed.append, 11)
    native.flush_deferred()            # still capturing: nothing happens
    assert destroyed == [] and len(native._deferred) == 1
    state["capturing"] = False
    native._release(destroyed.append, 12)  # outside a capture: destroyed at once
    assert destroyed == [12]
    native.flush_deferred()
    assert destroyed == [12, 11] and native._deferred == []


def test_capture_guard_parks_destructors_of_other_threads(monkeypatch):
    """capture_guard announces a capture to destructors on every thread (the collector may run on a sampler thread whose
    current stream is not the capturing one) and releases what was parked when it closes."""
    import threading

    from odil_b200 import native

    destroyed = []
    monkeypatch.setattr(native, "_deferred", [])
    with native.capture_guard():
        t = threading.Thread(target=lambda: native._release(destroyed.append, 7))
        t.start()
        t.join()
        assert destroyed == [] and len(native._deferred) == 1
    assert destroyed == [7] and native._deferred == [] and native._capture_depth == 0


def test_lbfgs_optimizers_host_glue(monkeypatch):
    """LbfgsbOptimizer (SciPy on the host) and LbfgsDeviceOptimizer (vector algebra replaced by torch-CPU stand-ins
    here) behind the optimizer seam: same minimiser, arrays in the problem dtype and shape, one callback per iteration
    with arrays the callback may keep (they are not the optimizer's conversion buffers)."""
    from odil_b200 import native
    from odil_b200.optimizer import LbfgsbOptimizer, LbfgsDeviceOptimizer

    monkeypatch.setattr(native, "multi_dot", lambda V, k, g, out: out.__setitem__(slice(0, k), V[:k] @ g))
    monkeypatch.setattr(native, "multi_axpy", lambda V, k, coef, a0, g, d: d.copy_(a0 * g + coef[:k] @ V[:k]))
    monkeypatch.setattr(native, "dot", lambda a, b, out: out.__setitem__(0, torch.dot(a, b)))
    monkeypatch.setattr(native, "axpby", lambda a, x, b, y: y.copy_(a * x + (b * y if b != 0 else 0)))
    rng = np.random.default_rng(2)
    A = rng.standard_normal((12, 12))
    A = torch.as_tensor(A @ A.T + 0.5 * np.eye(12))
    b = torch.as_tensor(rng.standard_normal(12))

    def loss_grad(arrays):
        x = torch.cat([a.reshape(-1) for a in arrays]).to(torch.float64)
        r = A @ x - b
        g = (A.T @ r).to(arrays[0].dtype)
        return 0.5 * float(r @ r), [g[:8].reshape(2, 4), g[8:].reshape(4)], {"loss": 0.5 * float(r @ r)}

    results = {}
    for cls in (LbfgsbOptimizer, LbfgsDeviceOptimizer):
        seen = []
        x0 = [torch.zeros(2, 4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64)]
        arrays, info = cls(m=5, dtype=np.float64).run(x0, loss_grad, epochs=6,
                                                      callback=lambda arr, epoch, pinfo: seen.append((epoch, arr)))
        assert [e for e, _ in seen] == [1, 2, 3, 4, 5, 6] and info.epochs == 6
        assert [tuple(a.shape) for a in arrays] == [(2, 4), (4,)] and arrays[0].dtype == torch.float64
        # the arrays of successive callbacks are distinct objects holding distinct iterates
        assert len({a[0].data_ptr() for _, a in seen}) == 6
        assert not torch.equal(seen[0][1][0], seen[5][1][0])
        assert torch.equal(seen[5][1][0], arrays[0]) and torch.equal(seen[5][1][1], arrays[1])
        results[cls.__name__] = torch.cat([a.reshape(-1) for a in arrays])
    assert torch.allclose(results["LbfgsbOptimizer"], results["LbfgsDeviceOptimizer"], rtol=1e-7, atol=1e-9)
