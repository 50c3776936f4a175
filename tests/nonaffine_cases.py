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
         "poisson_mgloss", "raw_term", "heat3"]
# cases whose Jacobian (Problem.linearize) is pinned as well; multigrid off
NEWTON_CASES = ["heat_k", "newton", "infer_constant", "heat3"]


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


def heat3_operator(ctx):
    """BASELINE configs[4]: heat equation u_t = div(k(u) grad u) on a (t, x, y) grid with the conductivity of
    examples/heat/heat.py:22-24, discretised like operator_odil (heat.py:36-137) with one more space dimension:
    time-centred fluxes through the four faces of a cell, k evaluated at FROZEN face averages (heat.py:86-90), zero
    Dirichlet walls by quadratic extrapolation, the initial condition imposed through the t-1 neighbours by linear
    extrapolation, plus the imposed-data term of the inverse problem."""
    import odil

    mod, extra = ctx.mod, ctx.extra
    dt, dx, dy = ctx.step()
    it, ix, iy = ctx.indices()
    nt, nx, ny = ctx.size()
    zero = ctx.cast(0)

    def stencil(frozen):
        st = {}
        for lvl in (0, -1):
            for name, sh in [("c", (0, 0)), ("xm", (-1, 0)), ("xp", (1, 0)), ("ym", (0, -1)), ("yp", (0, 1))]:
                st[(lvl, name)] = ctx.field("u", lvl, *sh, frozen=frozen)
        # initial condition: value at the lower time level of the first layer, by linear extrapolation (heat.py:62-70)
        u0 = extra.init_u
        q0 = {"c": u0, "xm": mod.roll(u0, 1, axis=0), "xp": mod.roll(u0, -1, axis=0), "ym": mod.roll(u0, 1, axis=1),
              "yp": mod.roll(u0, -1, axis=1)}
        for name in ["c", "xm", "xp", "ym", "yp"]:
            st[(-1, name)] = mod.where(it == 0, odil.core.extrap_linear(st[(0, name)], q0[name][None]), st[(-1, name)])
        ex = odil.core.extrap_quadh
        for lvl in (0, -1):
            c = st[(lvl, "c")]
            xm, xp, ym, yp = st[(lvl, "xm")], st[(lvl, "xp")], st[(lvl, "ym")], st[(lvl, "yp")]
            st[(lvl, "xm")] = mod.where(ix == 0, ex(xp, c, 0), xm)
            st[(lvl, "xp")] = mod.where(ix == nx - 1, ex(st[(lvl, "xm")], c, 0), xp)
            st[(lvl, "ym")] = mod.where(iy == 0, ex(yp, c, 0), ym)
            st[(lvl, "yp")] = mod.where(iy == ny - 1, ex(st[(lvl, "ym")], c, 0), yp)
        return st

    def conductivity(u):
        return 0.02 * mod.exp(-((u - 0.5) ** 2) * 20)

    q, qf = stencil(False), stencil(True)
    mid = lambda s, name: s[(0, name)] + s[(-1, name)]
    u_t = (q[(0, "c")] - q[(-1, "c")]) / dt
    flux = 0
    for name, h, sign in [("xm", dx, -1), ("xp", dx, 1), ("ym", dy, -1), ("yp", dy, 1)]:
        grad = sign * (mid(q, name) - mid(q, "c")) / (2 * h)     # time-centred normal derivative at the face
        k = conductivity((mid(qf, name) + mid(qf, "c")) * 0.25)  # frozen face value
        flux = flux + sign * grad * k / h
    fu = u_t - flux
    res = [("fu", fu)]
    kimp = extra.kimp * (np.prod(ctx.size()) / extra.imp_size) ** 0.5
    res += [("imp", extra.imp_mask * (q[(0, "c")] - extra.imp_u) * kimp)]
    return res


def make_heat3(odil, mod, dtype, cshape, seed=3, device_data=False):
    """Problem data of configs[4]: Gaussian initial profile, imposed values of a smooth field at ~2 % of the points."""
    nt, nx, ny = cshape
    domain = odil.Domain(cshape=tuple(cshape), dimnames=("t", "x", "y"), lower=(0, 0, 0), upper=(1, 1, 1), dtype=dtype,
                         multigrid=False, mod=mod)
    x1 = (np.arange(nx) + 0.5) / nx
    y1 = (np.arange(ny) + 0.5) / ny
    X, Y = np.meshgrid(x1, y1, indexing="ij")
    extra = _ns(kimp=2.0)
    extra.init_u = np.exp(-((X - 0.5) ** 2 + (Y - 0.5) ** 2) * 50).astype(dtype)
    rng = np.random.default_rng(seed)
    if device_data:  # large grids: build the data on the device (bench.py)
        import torch

        gen = torch.Generator(device="cuda").manual_seed(seed)
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        mask = (torch.rand(cshape, device="cuda", generator=gen) < 0.02).to(tdt)
        extra.imp_size = int(mask.sum().item())
        extra.imp_mask = odil.backend.Known(mask)
        t1 = torch.linspace(0.5 / nt, 1 - 0.5 / nt, nt, device="cuda", dtype=tdt)
        decay = torch.exp(-3 * t1)[:, None, None]
        extra.imp_u = odil.backend.Known(decay * torch.as_tensor(extra.init_u, device="cuda")[None])
    else:
        mask = rng.random(cshape) < 0.2
        extra.imp_size = int(mask.sum())
        extra.imp_mask = mask.astype(dtype)
        t1 = (np.arange(nt) + 0.5) / nt
        extra.imp_u = (np.exp(-3 * t1)[:, None, None] * extra.init_u[None]).astype(dtype)
    state = odil.State(fields={"u": odil.Field(np.zeros(cshape, dtype=dtype), loc="ccc")})
    state = domain.init_state(state)
    return heat3_operator, domain, state, extra, {"epoch": 0}


def build(case, odil, mod, dtype, scripts):
    """Returns (operator, domain, state, extra, tracers)."""
    if case == "heat3":
        return make_heat3(odil, mod, dtype, (6, 5, 4))
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
            "newton": ["test_newton"], "poisson_mgloss": ["poisson"], "raw_term": [], "heat3": []}[case]
