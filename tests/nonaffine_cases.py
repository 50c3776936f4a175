"""
Parity cases for operators that are NOT affine stencils (SURVEY.md 8f-1/2/3, a-1 loc changes, a-5 Raw terms).

Every case pairs an UNMODIFIED operator function of the reference (examples/heat/heat.py:operator_odil,
examples/velocity_from_tracer/veltracer.py:operator_advection, examples/heat_tmax/heat_tmax.py:operator_heat,
examples/infer_constant/infer_constant.py:operator_adv, tests/test_optimize.py:operator, tests/test_newton.py:operator,
examples/poisson/poisson.py:operator with mgloss) with a small set-up written against the public ODIL API only.
`build(case, odil, mod, dtype, scripts)` is executed twice with the same code:

  * by tests/golden/make_nonaffine_goldens.py with `odil` = the REFERENCE package and `mod` = the torch shim
    (oracle/ref_shim.py): the reference's own core.py evaluates the operator, torch.autograd differentiates it;
  * by the tests with `odil` = this repository's package and `mod` = ModB200: the operator is traced into the
    expression graph and run by the generated kernels (on the GPU; through their host twin in the CPU suite).

State arrays are not initialised here: both sides load them from the golden file (`arrays_from_state` order).
"""
import argparse

import numpy as np

SCRIPTS = {
    "heat": "examples/heat/heat.py",
    "veltracer": "examples/velocity_from_tracer/veltracer.py",
    "heat_tmax": "examples/heat_tmax/heat_tmax.py",
    "infer_constant": "examples/infer_constant/infer_constant.py",
    "test_optimize": "tests/test_optimize.py",
    "test_newton": "tests/test_newton.py",
    "poisson": "examples/poisson/poisson.py",
}

CASES = ["heat_k", "heat_k_mg", "heat_knet", "veltracer", "heat_tmax", "infer_constant", "optimize", "newton",
         "poisson_mgloss", "raw_term"]
# cases whose Jacobian (Problem.linearize) is pinned as well; multigrid off
NEWTON_CASES = ["heat_k", "newton", "infer_constant"]


def _ns(**kw):
    return argparse.Namespace(**kw)


def _heat(odil, mod, dtype, heat, infer_k, multigrid):
    Nt, Nx = (8, 8) if multigrid else (8, 6)
    domain = odil.Domain(cshape=(Nt, Nx), dimnames=("t", "x"), lower=(0, 0), upper=(1, 1), dtype=dtype,
                         multigrid=multigrid, mg_nlvl=2 if multigrid else None, mod=mod)
    args = _ns(keep_frozen=1, keep_init=1, infer_k=infer_k, kmax=0.1, kimp=2.0, kxreg=0.3, kxregdecay=5.0, ktreg=0.2,
               ktregdecay=0.0, kwreg=0.1 if infer_k else 0.0, kwregdecay=3.0)
    rng = np.random.default_rng(21)
    x1 = (np.arange(Nx) + 0.5) / Nx
    extra = _ns(args=args)
    extra.init_u = (np.exp(-((x1 - 0.5) ** 2) * 50) - np.exp(-0.25 * 50)).astype(dtype)
    mask = (rng.random((Nt, Nx)) < 0.3)
    extra.imp_size = int(mask.sum())
    extra.imp_mask = mask.astype(dtype)
    extra.imp_u = rng.random((Nt, Nx)).astype(dtype)
    fields = {"u": odil.Field(np.zeros((Nt, Nx), dtype=dtype), loc="cc")}
    if infer_k:
        fields["k_net"] = domain.make_neural_net([1, 5, 5, 1])
    state = domain.init_state(odil.State(fields=fields))
    return heat.operator_odil, domain, state, extra, {"epoch": 7}


def _veltracer(odil, mod, dtype, vt):
    domain = odil.Domain(cshape=(6, 8, 8), dimnames=("t", "x", "y"), lower=(0, 0, 0), upper=(1, 1, 1), dtype=dtype,
                         multigrid=True, mg_nlvl=2, mod=mod)
    x1 = (np.arange(8) + 0.5) / 8
    X, Y = np.meshgrid(x1, x1, indexing="ij")
    extra = _ns(args=_ns(kxreg=0.01, ktreg=1.0, kimp=10.0))
    extra.u_init = np.asarray(vt.u_init_blob(X, Y, 0), dtype=dtype)
    extra.u_final = np.asarray(vt.u_init_blob(X, Y, 1), dtype=dtype)
    state = odil.State()
    for key in ["u", "vx", "vy"]:
        state.fields[key] = odil.Field(None, loc="ncc")
    state = domain.init_state(state)
    return vt.operator_advection, domain, state, extra, {"epoch": 0}


def _heat_tmax(odil, mod, dtype, ht):
    Nt, Nx = 8, 6
    domain = odil.Domain(cshape=(Nt, Nx), dimnames=("t", "x"), lower=(0, 0), upper=(1, np.pi), dtype=dtype,
                         multigrid=True, mg_nlvl=2, mod=mod)
    args = _ns(kimp=1.0, tmax_ref=4.5, tmax_init=1.0)
    xone = domain.points_1d("x", loc="c")
    extra = _ns(args=args)
    extra.u_init = np.asarray(ht.get_ref_u(np.full_like(xone, 0.0), xone, args), dtype=dtype)
    extra.u_final = np.asarray(ht.get_ref_u(np.full_like(xone, 1.0), xone, args), dtype=dtype)
    state = odil.State(fields={"u": odil.Field(np.tile(extra.u_init, [Nt + 1, 1]), loc="nc"),
                               "coeff": odil.Array([args.tmax_init])})
    state = domain.init_state(state)
    return ht.operator_heat, domain, state, extra, {"epoch": 0}


def _infer_constant(odil, mod, dtype, ic):
    Nt, Nx = 8, 6
    domain = odil.Domain(cshape=(Nt, Nx), dimnames=("t", "x"), lower=(0, -1), upper=(1, 1), dtype=dtype,
                         multigrid=False, mod=mod)
    args = _ns(c_diff=0.01, c_src=0.1, c_vel=0.2)
    xone = domain.points_1d("x", loc="c")
    extra = _ns(args=args)
    extra.u_init = np.asarray(ic.get_ref_u(xone * 0 + 0.0, xone, args), dtype=dtype)
    extra.u_final = np.asarray(ic.get_ref_u(xone * 0 + 1.0, xone, args), dtype=dtype)
    state = odil.State(fields={"coeff": odil.Array([0, 0, 0.001]), "u": odil.Field(None, loc="nc")})
    state = domain.init_state(state)
    return ic.operator_adv, domain, state, extra, {"epoch": 0}


def _optimize(odil, mod, dtype, to):
    """tests/test_optimize.py:make_problem with the backend and dtype made explicit."""
    Nx, Ny = 8, 4
    domain = odil.Domain(cshape=(Nx, Ny), dimnames=["x", "y"], lower=(0, 0), upper=(2, 1), dtype=dtype, multigrid=True,
                         mg_axes=[True, True], mg_nlvl=2, mod=mod)
    state = odil.State(fields={
        "uc": odil.Field(np.zeros(domain.size(loc="cc")), loc="cc"),
        "un": odil.Field(np.zeros(domain.size(loc="nn")), loc="nn"),
        "ufx": odil.Field(np.zeros(domain.size(loc="nc")), loc="nc"),
        "ufy": odil.Field(np.zeros(domain.size(loc="cn")), loc="cn"),
        "a": odil.Array(np.zeros(5)),
        "net": domain.make_neural_net([1, 7, 1]),
    })
    state = domain.init_state(state)

    def func(x, y):
        return x * 0.25 + y * 0.5

    extra = _ns()
    extra.ref = {loc_key: func(*[np.asarray(p) for p in domain.points(loc=loc)])
                 for loc_key, loc in [("uc", "cc"), ("un", "nn"), ("ufx", "nc"), ("ufy", "cn")]}
    extra.ref["a"] = np.arange(5, dtype=dtype)
    extra.ref["net_a"] = extra.ref["a"] * 0.5
    return to.operator, domain, state, extra, {"epoch": 0}


def _newton(odil, mod, dtype, tn):
    """tests/test_newton.py:make_problem with the backend and dtype made explicit."""
    Nx, Ny, Na, Nnet = 3, 2, 5, 5
    domain = odil.Domain(cshape=(Nx, Ny), dimnames=["x", "y"], lower=(0, 0), dtype=dtype, upper=(Nx, Ny),
                         multigrid=False, mod=mod)
    state = odil.State(fields={
        "uc": odil.Field(np.ones(domain.size(loc="cc")), loc="cc"),
        "ufx": odil.Field(np.ones(domain.size(loc="nc")), loc="nc"),
        "a": odil.Array(np.zeros(Na, dtype=dtype)),
        "net": domain.make_neural_net([Nnet, Nnet], activation="none"),
    })
    state = domain.init_state(state)
    xc, yc = [np.asarray(p) for p in domain.points(loc="cc")]
    xfx, yfx = [np.asarray(p) for p in domain.points(loc="nc")]
    rng = np.random.default_rng(33)
    extra = _ns(args=_ns(Nnet=Nnet))
    extra.ref = {"uc": 0.25 * xc * yc, "ufx": 0.25 * xfx * yfx, "dudx": 0.25 * yc,
                 "a": np.linspace(0, 1, Na, dtype=dtype),
                 "net_in": rng.random((Nnet, Nnet + 1)).astype(dtype),
                 "net_out": rng.random((Nnet, Nnet + 1)).astype(dtype)}
    return tn.operator, domain, state, extra, {"epoch": 0}


def _poisson_mgloss(odil, mod, dtype, poisson):
    domain = odil.Domain(cshape=[8, 8], dimnames=["x", "y"], lower=0.0, upper=1.0, dtype=dtype, multigrid=True,
                         mg_nlvl=2, mod=mod)
    args = _ns(mgloss=1, ref="hat", osc_k=2)
    rng = np.random.default_rng(5)
    extra = _ns(args=args, rhs=rng.standard_normal((8, 8)).astype(dtype))
    state = odil.State()
    state.fields["u"] = None
    state = domain.init_state(state)
    return poisson.operator, domain, state, extra, {"epoch": 0}


def raw_operator(ctx):
    """A squared residual plus a `Context.Raw` regulariser (mean(value) instead of mean(value^2), core.py:867-870,
    :1093), a node-to-cell location change and an epoch-dependent weight."""
    mod = ctx.mod
    u = ctx.field("u")
    up = ctx.field("u", 1, 0)
    fn = ctx.field("fn", 1, 0, loc="cc") - ctx.field("fn", 0, 0, loc="cc")
    k = 0.5 ** (ctx.tracers["epoch"] / 4)
    return [("res", mod.tanh(u) * up - fn), ("reg", ctx.Raw(mod.square(up - u) * k))]


def _raw_term(odil, mod, dtype):
    domain = odil.Domain(cshape=(6, 4), dimnames=["x", "y"], lower=0.0, upper=1.0, dtype=dtype, multigrid=False,
                         mod=mod)
    state = odil.State(fields={"u": odil.Field(np.zeros((6, 4)), loc="cc"),
                               "fn": odil.Field(np.zeros((7, 4)), loc="nc")})
    state = domain.init_state(state)
    return raw_operator, domain, state, _ns(), {"epoch": 3}


def build(case, odil, mod, dtype, scripts):
    """Returns (operator, domain, state, extra, tracers)."""
    if case == "heat_k":
        return _heat(odil, mod, dtype, scripts["heat"], 0, False)
    if case == "heat_k_mg":
        return _heat(odil, mod, dtype, scripts["heat"], 0, True)
    if case == "heat_knet":
        return _heat(odil, mod, dtype, scripts["heat"], 1, False)
    if case == "veltracer":
        return _veltracer(odil, mod, dtype, scripts["veltracer"])
    if case == "heat_tmax":
        return _heat_tmax(odil, mod, dtype, scripts["heat_tmax"])
    if case == "infer_constant":
        return _infer_constant(odil, mod, dtype, scripts["infer_constant"])
    if case == "optimize":
        return _optimize(odil, mod, dtype, scripts["test_optimize"])
    if case == "newton":
        return _newton(odil, mod, dtype, scripts["test_newton"])
    if case == "poisson_mgloss":
        return _poisson_mgloss(odil, mod, dtype, scripts["poisson"])
    if case == "raw_term":
        return _raw_term(odil, mod, dtype)
    raise KeyError(case)


def scripts_for(case):
    return {"heat_k": ["heat"], "heat_k_mg": ["heat"], "heat_knet": ["heat"], "veltracer": ["veltracer"],
            "heat_tmax": ["heat_tmax"], "infer_constant": ["infer_constant"], "optimize": ["test_optimize"],
            "newton": ["test_newton"], "poisson_mgloss": ["poisson"], "raw_term": []}[case]
