"""
Drop-in check at the script level: the UNMODIFIED example scripts of the reference
(/root/reference/examples/poisson/poisson.py, examples/wave/wave.py) are loaded against THIS repository's `odil`
package -- their own argument parsers, `make_problem`, operators and boundary helpers run as they are -- and the
operators are traced and lowered (`ResidualEngine(trace_only=True)`, no GPU).  The plans must be the ones the oracle
writes down.  Needs the reference checkout, so it is skipped where /root/reference does not exist (the GPU box);
matplotlib is absent from this image and is replaced by an inert stand-in while the scripts import.
"""
import os
import runpy
import sys
import types

import numpy as np
import pytest

import odil
from odil_b200.engine import ResidualEngine
from oracle import odil_oracle as orc
from tests.test_host_cpu import plan_apply

REF = os.environ.get("ODIL_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")), reason="reference checkout absent")


class _Inert(types.ModuleType):
    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return lambda *a, **k: None


@pytest.fixture
def load_example(monkeypatch):
    def load(relpath, argv):
        for name in ["matplotlib", "matplotlib.pyplot"]:
            if name not in sys.modules:
                m = _Inert(name)
                m.__file__ = os.devnull
                monkeypatch.setitem(sys.modules, name, m)
        monkeypatch.setattr(sys, "argv", [relpath] + argv)
        ns = runpy.run_path(os.path.join(REF, relpath), run_name="example")
        assert ns["odil"] is odil  # the script imported this repository's package, not the reference's
        return ns
    return load


@pytest.mark.parametrize("ndim,N", [(1, 32), (2, 16), (3, 8)])
def test_reference_poisson_script_traces_to_the_oracle_plan(load_example, ndim, N):
    ns = load_example("examples/poisson/poisson.py", ["--ndim", str(ndim), "--N", str(N)])
    args = ns["parse_args"]()
    assert args.optimizer == "adam" and args.multigrid == 1 and args.double == 1 and args.lr == 0.005
    problem, state = ns["make_problem"](args)
    assert problem.domain.multigrid and isinstance(state.fields["u"], odil.MultigridField)
    eng = ResidualEngine(problem, state, trace_only=True)
    out = eng.outputs[0]
    assert len(eng.outputs) == 1 and out.fused
    spec = out.blocks[0].spec
    steps = [1.0 / N] * ndim
    offsets, table, rr = orc.poisson_plan(ndim, steps)
    assert spec["rwidth"] == tuple(rr)
    order = [list(map(tuple, spec["offsets"])).index(tuple(o)) for o in offsets]
    got = np.asarray(spec["table"]).reshape(-1, len(offsets))[:, order]
    assert np.allclose(got, np.asarray(table).reshape(-1, len(offsets)), rtol=1e-12, atol=0)
    # and the lowered plan applied to a random field equals the directly written operator with the script's rhs
    rng = np.random.default_rng(0)
    U = rng.standard_normal((N,) * ndim)
    F = plan_apply(spec, U, out.const.cpu().numpy())
    F_ref = orc.poisson_residual(U, np.asarray(problem.extra.rhs), steps)
    assert np.max(np.abs(F - F_ref)) < 1e-9 * np.max(np.abs(F_ref))


def test_reference_wave_operator_traces_to_the_oracle_plan(load_example):
    """wave.py evaluates its exact solution with TensorFlow itself (`tf.Variable` / `tf.GradientTape` in
    `get_exact`, wave.py:13-26), so its `make_problem` cannot run without TF whatever backend odil uses; the
    script's own `operator_wave` and argument parser are taken as they are and the boundary / initial data come
    from the same formula evaluated with NumPy (tests/operators.py)."""
    from tests import operators as ops

    ns = load_example("examples/wave/wave.py", ["--Nt", "16", "--Nx", "12", "--multigrid", "0"])
    args = ns["parse_args"]()
    assert args.optimizer == "lbfgsb" and args.double == 1 and args.kimp == 1
    mine, state = ops.make_wave((args.Nt, args.Nx), 0, np.float64)
    extra = mine.extra
    extra.args = args
    problem = odil.Problem(ns["operator_wave"], mine.domain, extra)
    eng = ResidualEngine(problem, state, trace_only=True)
    out = eng.outputs[0]
    assert eng.names == ["fu"] and out.fused
    spec = out.blocks[0].spec
    assert spec["rwidth"] == (2, 1)
    assert sorted(map(tuple, spec["offsets"])) == sorted([(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)])
    rng = np.random.default_rng(1)
    U = rng.standard_normal((16, 12))
    F = plan_apply(spec, U, out.const.cpu().numpy())
    F_ref = orc.wave_residual(U, 1.0 / 16, 2.0 / 12, np.asarray(extra.left_u), np.asarray(extra.right_u),
                              np.asarray(extra.init_u), np.asarray(extra.init_ut), args.kimp)
    assert np.max(np.abs(F - F_ref)) < 1e-10 * np.max(np.abs(F_ref))


def test_reference_basic_fields_script_traces(load_example):
    """examples/basic/fields.py: cell, node and face fields as multigrid unknowns plus an unused NeuralNet in the
    state; every output is `field - func(points)`, i.e. an identity plan with the constant -func."""
    ns = load_example("examples/basic/fields.py", [])
    args = ns["parse_args"]()
    problem, state = ns["make_problem"](args)
    domain = problem.domain
    assert {k: type(v).__name__ for k, v in state.fields.items()} == {
        "uc": "MultigridField", "un": "MultigridField", "ufx": "MultigridField", "ufy": "MultigridField",
        "net": "NeuralNet"}
    eng = ResidualEngine(problem, state, trace_only=True)
    assert eng.names == ["uc", "un", "ufx", "ufy"] and all(o.fused for o in eng.outputs)
    for out, loc in zip(eng.outputs, ["cc", "nn", "nc", "cn"]):
        spec = out.blocks[0].spec
        assert np.asarray(spec["table"]).tolist() == [[1.0]] and tuple(map(tuple, spec["offsets"])) == ((0, 0),)
        x, y = (np.asarray(p) for p in domain.points(loc=loc))
        assert tuple(out.shape) == tuple(domain.size(loc=loc))
        assert np.allclose(out.const.cpu().numpy(), -(x * 0.25 + y * 0.5), rtol=0, atol=1e-15)


@pytest.mark.parametrize("cmd", [["tests/test_domain.py"], ["-m", "pytest", "-q", "-p", "no:cacheprovider",
                                                           "tests/test_io.py"]])
def test_reference_host_side_tests_pass_unmodified(cmd, tmp_path):
    """The reference's own host-side tests (state packing round trips, tests/test_domain.py; RAW + XMF round trip,
    tests/test_io.py) run as they are with this repository's `odil` on the path.  (Its other tests evaluate
    operators and therefore need the GPU.)"""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    argv = [sys.executable] + [os.path.join(REF, a) if a.endswith(".py") else a for a in cmd]
    env = dict(os.environ, PYTHONPATH=root)
    r = subprocess.run(argv, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert "FAIL" not in out, out[-2000:]
    if cmd[0].endswith("test_domain.py"):
        assert out.count("PASS") == 4
