#!/usr/bin/env python3
"""
Generates the golden vectors under tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src/odil + examples/poisson/poisson.py + examples/wave/wave.py) in this
container.  The reference's JAX/TF backends are not installable here (no wheels, no network),
so the reference code is driven through

  * `odil.backend.ModNumpy()` (the reference's own NumPy namespace) for forward values, and
  * `TorchMod` below: a thin `mod` namespace over torch (CPU, fp64/fp32) that lets
    `torch.autograd` stand in for `jax.value_and_grad` (reference core.py:1100) while every
    arithmetic statement that runs is still the reference's own core.py / example operator.

Run:   python tests/golden/make_goldens.py        (needs /root/reference; NOT run on the GPU box)
The .npz files it writes are committed; tests never import /root/reference.
"""
import argparse
import os
import sys

import numpy as np
import torch

REF = os.environ.get("ODIL_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------------
# The reference is imported and driven through oracle/ref_shim.py (stubs for what the image lacks + the torch `mod`
# shim); the same module drives the reference arm of bench.py.
# ----------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.abspath(os.path.join(OUT, "..", "..")))
from oracle import ref_shim  # noqa: E402

TorchMod = ref_shim.TorchMod


def import_reference():
    odil, mods, origin = ref_shim.import_reference(("poisson", "wave"))
    assert origin == REF, origin   # goldens come from /root/reference itself, not from a staged copy
    return odil, mods["poisson"], mods["wave"]


def numpy_conv_valid(input, filters, strides, padding):
    """Stand-in for jax.lax.conv (cross-correlation, VALID) used by reference core.py:751."""
    from scipy.signal import correlate

    assert padding == "VALID"
    if isinstance(strides, int):
        strides = (strides,) * input.ndim
    res = correlate(input, filters, mode="valid", method="direct")
    return res[tuple(slice(None, None, s) for s in strides)]


# ----------------------------------------------------------------------------------------------
# Case builders
# ----------------------------------------------------------------------------------------------
def gen_interp(odil, out):
    mod = odil.backend.ModNumpy()
    rng = np.random.default_rng(11)
    for ndim in [1, 2, 3, 4]:
        for loc4 in ["cccc", "nnnn", "cnnn", "nccc", "c.cn"]:
            loc = loc4[:ndim]
            cshapeh = 3 + np.arange(ndim)
            shapeh = tuple(int(s + (1 if l == "n" else 0)) for s, l in zip(cshapeh, loc))
            uh = rng.standard_normal(shapeh)
            ui = odil.core.interp_to_finer(uh, loc=loc, mod=mod, method="stack")
            out[f"interp_{ndim}_{loc}_in"] = uh
            out[f"interp_{ndim}_{loc}_out"] = np.asarray(ui)


def gen_restrict(odil, out):
    mod = odil.backend.ModNumpy()
    mod.convolution = numpy_conv_valid
    rng = np.random.default_rng(12)
    for ndim in [1, 2, 3, 4]:
        for loc4 in ["cccc", "nnnn", "cnnn", "nccc"]:
            loc = loc4[:ndim]
            cshape = (3 + np.arange(ndim)) * 2
            shape = tuple(int(s + (1 if l == "n" else 0)) for s, l in zip(cshape, loc))
            u = rng.standard_normal(shape)
            ur = odil.core.restrict_to_coarser(u, loc=loc, mod=mod, method="conv")
            out[f"restrict_{ndim}_{loc}_in"] = u
            out[f"restrict_{ndim}_{loc}_out"] = np.asarray(ur)


def make_ns(**kw):
    return argparse.Namespace(**kw)


def poisson_setup(odil, poisson, cshape, nlvl, dtype, seed, lower=0.0, upper=1.0):
    """Domain + rhs (numpy path) + random multigrid terms for the reference Poisson operator."""
    ndim = len(cshape)
    modn = odil.backend.ModNumpy()
    dimnames = ["x", "y", "z", "sx"][:ndim]
    domain = odil.Domain(cshape=list(cshape), dimnames=dimnames, lower=lower, upper=upper, dtype=dtype,
                         multigrid=nlvl > 0, mg_nlvl=nlvl if nlvl > 0 else None, mod=modn)
    args = make_ns(mgloss=0, ref="hat", osc_k=2)
    ref_u = poisson.get_ref_u("hat", args, domain)
    rhs = np.asarray(poisson.get_discrete_rhs(ref_u, domain, modn))
    rng = np.random.default_rng(seed)
    if nlvl > 0:
        terms = [rng.standard_normal(cs).astype(dtype) for cs in domain.mg_cshapes]
    else:
        terms = [rng.standard_normal(cshape).astype(dtype)]
    return domain, args, ref_u.astype(dtype), rhs.astype(dtype), terms


def state_from_terms(odil, domain, key, terms, loc=None):
    loc = loc or "c" * domain.ndim
    state = odil.State()
    if domain.multigrid:
        fterms = [odil.Field(t, loc=loc, cshape=tuple(cs)) for t, cs in zip(terms, domain.mg_cshapes)]
        state.fields[key] = odil.MultigridField(terms=fterms, loc=loc, factors=[1] * len(terms))
    else:
        state.fields[key] = odil.Field(terms[0], loc=loc, cshape=tuple(domain.cshape))
    return domain.init_state(state)


def eval_loss_grad_torch(odil, operator, domain_np, extra_np, terms, key, tdtype, tracers=None):
    """
    Re-runs reference Context + operator under TorchMod and differentiates with autograd.
    Mirrors reference core.py:1082-1096 (loss assembly) verbatim in meaning.
    """
    tm = TorchMod(tdtype)
    npdtype = np.float64 if tdtype == torch.float64 else np.float32
    domain = odil.Domain(cshape=list(domain_np.cshape), dimnames=list(domain_np.dimnames),
                         lower=domain_np.lower, upper=domain_np.upper, dtype=npdtype,
                         multigrid=domain_np.multigrid,
                         mg_nlvl=domain_np.mg_nlvl if domain_np.multigrid else None, mod=tm)
    leaves = [torch.tensor(np.asarray(t), dtype=tdtype, requires_grad=True) for t in terms]
    state = state_from_terms(odil, domain, key, leaves)
    # init_state passes arrays through mod.variable -> same leaf objects (as_tensor keeps identity)
    domain.arrays_to_state(leaves, state)
    extra = argparse.Namespace(**vars(extra_np))
    for k, v in vars(extra).items():
        if isinstance(v, np.ndarray):
            setattr(extra, k, torch.tensor(v, dtype=tdtype))
    ctx = odil.core.Context(domain, state, extra=extra, tracers=tracers or {"epoch": 0})
    ff = operator(ctx)
    names = [f[0] if isinstance(f, tuple) else "" for f in ff]
    values = [f[1] if isinstance(f, tuple) else f for f in ff]
    terms_l = [tm.mean(tm.square(v)) for v in values]
    loss = sum(terms_l)
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(x) for g, x in zip(grads, leaves)]
    return (loss.detach().numpy(), [g.numpy() for g in grads], [t.detach().numpy() for t in terms_l], names,
            [v.detach().numpy() for v in values])


def gen_poisson(odil, poisson, out):
    cases = {
        "p1d_16_L0": ((16,), 0),
        "p1d_16_L3": ((16,), 3),
        "p2d_16_L3": ((16, 16), 3),
        "p2d_12x8_L0": ((12, 8), 0),
        "p3d_8_L3": ((8, 8, 8), 3),
        "p3d_16x8x12_L2": ((16, 8, 12), 2),
        "p3d_12_L0": ((12, 12, 12), 0),
    }
    for name, (cshape, nlvl) in cases.items():
        for dt, tdt in [(np.float64, torch.float64), (np.float32, torch.float32)]:
            tag = name + ("_f64" if dt == np.float64 else "_f32")
            domain, args, ref_u, rhs, terms = poisson_setup(odil, poisson, cshape, nlvl, dt, seed=100 + len(name))
            extra = make_ns(args=args, rhs=rhs, ref_u=ref_u)
            # Forward through the reference's own NumPy backend.
            state = state_from_terms(odil, domain, "u", terms)
            ctx = odil.core.Context(domain, state, extra=extra, tracers={"epoch": 0})
            F = np.asarray(poisson.operator(ctx)[0])
            U = np.asarray(domain.field(state, "u"))
            loss_np = np.mean(np.square(F))
            loss, grads, tl, names, values = eval_loss_grad_torch(odil, poisson.operator, domain, extra, terms,
                                                                  "u", tdt)
            if dt == np.float64:
                assert abs(loss - loss_np) <= 1e-12 * abs(loss_np), (loss, loss_np)
            out[tag + "_rhs"] = rhs
            out[tag + "_U"] = U
            out[tag + "_F"] = F
            out[tag + "_loss"] = np.asarray(loss)
            for i, (t, g) in enumerate(zip(terms, grads)):
                out[f"{tag}_term{i}"] = t
                out[f"{tag}_grad{i}"] = g


def wave_exact(t, x):
    ii = [1, 2, 3, 4, 5]
    u = np.zeros(np.broadcast(t, x).shape)
    ut = np.zeros_like(u)
    for i in ii:
        k = i * np.pi
        u = u + np.cos((x - t + 0.5) * k) + np.cos((x + t - 0.5) * k)
        ut = ut + k * np.sin((x - t + 0.5) * k) - k * np.sin((x + t - 0.5) * k)
    return u / (2 * len(ii)), ut / (2 * len(ii))


def wave_setup(odil, cshape, nlvl, dtype, seed):
    modn = odil.backend.ModNumpy()
    domain = odil.Domain(cshape=tuple(cshape), dimnames=("t", "x"), lower=(0, -1), upper=(1, 1), dtype=dtype,
                         multigrid=nlvl > 0, mg_nlvl=nlvl if nlvl > 0 else None, mod=modn)
    t1, x1 = domain.points_1d()
    left_u, _ = wave_exact(t1, t1 * 0 + domain.lower[1])
    right_u, _ = wave_exact(t1, t1 * 0 + domain.upper[1])
    init_u, init_ut = wave_exact(x1 * 0 + domain.lower[0], x1)
    extra = make_ns(args=make_ns(kimp=1.0), left_u=left_u.astype(dtype), right_u=right_u.astype(dtype),
                    init_u=init_u.astype(dtype), init_ut=init_ut.astype(dtype))
    rng = np.random.default_rng(seed)
    if nlvl > 0:
        terms = [rng.standard_normal(cs).astype(dtype) for cs in domain.mg_cshapes]
    else:
        terms = [rng.standard_normal(cshape).astype(dtype)]
    return domain, extra, terms


def gen_wave(odil, wave, out):
    for name, (cshape, nlvl) in {"w_16x12_L0": ((16, 12), 0), "w_16x8_L2": ((16, 8), 2)}.items():
        for dt, tdt in [(np.float64, torch.float64), (np.float32, torch.float32)]:
            tag = name + ("_f64" if dt == np.float64 else "_f32")
            domain, extra, terms = wave_setup(odil, cshape, nlvl, dt, seed=7)
            state = state_from_terms(odil, domain, "u", terms)
            ctx = odil.core.Context(domain, state, extra=extra, tracers={"epoch": 0})
            F = np.asarray(wave.operator_wave(ctx)[0][1])
            loss, grads, tl, names, values = eval_loss_grad_torch(odil, wave.operator_wave, domain, extra, terms,
                                                                  "u", tdt)
            if dt == np.float64:
                assert abs(loss - np.mean(F ** 2)) <= 1e-12 * abs(loss)
            for k in ["left_u", "right_u", "init_u", "init_ut"]:
                out[f"{tag}_{k}"] = getattr(extra, k)
            out[tag + "_F"] = F
            out[tag + "_loss"] = np.asarray(loss)
            for i, (t, g) in enumerate(zip(terms, grads)):
                out[f"{tag}_term{i}"] = t
                out[f"{tag}_grad{i}"] = g


def load_test_operators():
    """tests/operators.py of this repository, imported with `odil` bound to the REFERENCE package: its operators are
    written against the public ODIL API only, so the unmodified reference core.py can run them."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("repo_test_operators", os.path.join(OUT, "..", "operators.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.odil.__file__.startswith(REF)
    return m


def wave2_setup(odil, ops, cshape, dtype, seed):
    modn = odil.backend.ModNumpy()
    domain = odil.Domain(cshape=tuple(cshape), dimnames=("t", "x", "y"), lower=(0, -1, -1), upper=(1, 1, 1),
                         dtype=dtype, multigrid=False, mod=modn)
    t1, x1, y1 = domain.points_1d()
    T, X, Y = np.meshgrid(t1, x1, y1, indexing="ij")
    lo, hi = domain.lower, domain.upper
    extra = make_ns(kimp=1.0)
    extra.xlo = ops.wave2_exact(T[:, 0, :], lo[1], Y[:, 0, :])[0].astype(dtype)
    extra.xhi = ops.wave2_exact(T[:, 0, :], hi[1], Y[:, 0, :])[0].astype(dtype)
    extra.ylo = ops.wave2_exact(T[:, :, 0], X[:, :, 0], lo[2])[0].astype(dtype)
    extra.yhi = ops.wave2_exact(T[:, :, 0], X[:, :, 0], hi[2])[0].astype(dtype)
    u0, ut0 = ops.wave2_exact(lo[0], X[0], Y[0])
    extra.init_u, extra.init_ut = u0.astype(dtype), ut0.astype(dtype)
    terms = [np.random.default_rng(seed).standard_normal(cshape).astype(dtype)]
    return domain, extra, terms


def gen_wave2(odil, out):
    """BASELINE configs[2]: the (t, x, y) wave operator of tests/operators.py::wave2_operator evaluated by the
    unmodified reference core.py (Context / field / roll / where): forward under ModNumpy, loss + gradient under
    the torch shim.  The input field is default_rng(4) N(0,1), as in tests/test_zz_wave2_gpu.py."""
    ops = load_test_operators()
    for name, cshape in {"w2_10x8x6": (10, 8, 6), "w2_20x16x24": (20, 16, 24)}.items():
        for dt, tdt in [(np.float64, torch.float64), (np.float32, torch.float32)]:
            tag = name + ("_f64" if dt == np.float64 else "_f32")
            domain, extra, terms = wave2_setup(odil, ops, cshape, dt, seed=4)
            state = state_from_terms(odil, domain, "u", terms)
            ctx = odil.core.Context(domain, state, extra=extra, tracers={"epoch": 0})
            F = np.asarray(ops.wave2_operator(ctx)[0][1])
            loss, grads, tl, names, values = eval_loss_grad_torch(odil, ops.wave2_operator, domain, extra, terms,
                                                                  "u", tdt)
            if dt == np.float64:
                assert abs(loss - np.mean(F ** 2)) <= 1e-12 * abs(loss)
            out[tag + "_U"] = terms[0]
            out[tag + "_F"] = F
            out[tag + "_loss"] = np.asarray(loss)
            out[tag + "_grad0"] = grads[0]
    # L-BFGS-B (SciPy, the reference's optimizer class) from a zero start, m = 20, 60 iterations, fp64
    domain, extra, terms = wave2_setup(odil, ops, (12, 12, 12), np.float64, seed=4)
    losses, arrays = run_reference_optimizer(odil, "lbfgsb", ops.wave2_operator, domain, extra,
                                             [t * 0 for t in terms], torch.float64, 60, lr=None, m=20)
    out["w2_lbfgsb_12_f64_losses"] = losses
    if arrays is not None:
        out["w2_lbfgsb_12_f64_x0"] = arrays[0]


def run_reference_optimizer(odil, optname, operator, domain_np, extra_np, terms0, tdtype, epochs, lr, **kw):
    """Runs the reference's own optimizer class (optimizer.py) on the torch-shim loss_grad."""
    import odil.optimizer as ropt

    tm = TorchMod(tdtype)
    npdtype = np.float64 if tdtype == torch.float64 else np.float32
    losses = []

    def loss_grad(arrays):
        arrays = [np.asarray(a.detach().numpy() if torch.is_tensor(a) else a) for a in arrays]
        loss, grads, tl, names, _ = eval_loss_grad_torch(odil, operator, domain_np, extra_np, arrays, "u", tdtype)
        pinfo = {"loss": loss}
        if optname.startswith("adam") or optname == "gd":
            grads = [torch.tensor(g) for g in grads]
        return loss, grads, pinfo

    def callback(arrays, epoch, pinfo):
        losses.append(float(pinfo["loss"]))

    opt = ropt.make_optimizer(optname, dtype=npdtype, mod=tm, **kw)
    x0 = [torch.tensor(t, dtype=tdtype) for t in terms0] if optname != "lbfgsb" else [np.asarray(t) for t in terms0]
    try:
        arrays, optinfo = opt.run(x0, loss_grad, epochs=epochs, callback=callback, lr=lr)
    except ropt.EarlyStopError:
        arrays = None
    if arrays is not None:
        arrays = [np.asarray(a.detach().numpy() if torch.is_tensor(a) else a) for a in arrays]
    return np.array(losses), arrays


def gen_optimizers(odil, poisson, wave, out):
    # Adam: 2D 16^2 L3, 20 epochs, fp64 + fp32 (callback loss = loss at the state BEFORE the update).
    for dt, tdt, tag in [(np.float64, torch.float64, "f64"), (np.float32, torch.float32, "f32")]:
        domain, args, ref_u, rhs, terms = poisson_setup(odil, poisson, (16, 16), 3, dt, seed=5)
        terms = [t * 0 for t in terms]  # start from zero as the example does
        extra = make_ns(args=args, rhs=rhs, ref_u=ref_u)
        losses, arrays = run_reference_optimizer(odil, "adam", poisson.operator, domain, extra, terms, tdt, 20,
                                                 lr=0.005)
        out[f"adam_p2d_16_L3_{tag}_rhs"] = rhs
        out[f"adam_p2d_16_L3_{tag}_losses"] = losses
        for i, a in enumerate(arrays):
            out[f"adam_p2d_16_L3_{tag}_x{i}"] = a
    # Config 1 (BASELINE.json configs[0]): 1D N=256, all 8 levels, fp64, Adam lr=0.005, 300 epochs.
    domain, args, ref_u, rhs, terms = poisson_setup(odil, poisson, (256,), 100, np.float64, seed=5)
    terms = [t * 0 for t in terms]
    extra = make_ns(args=args, rhs=rhs, ref_u=ref_u)
    losses, arrays = run_reference_optimizer(odil, "adam", poisson.operator, domain, extra, terms, torch.float64,
                                             300, lr=0.005)
    out["adam_p1d_256_f64_rhs"] = rhs
    out["adam_p1d_256_f64_ref_u"] = ref_u
    out["adam_p1d_256_f64_losses"] = losses
    out["adam_p1d_256_f64_nlvl"] = np.asarray(domain.mg_nlvl)
    # Same configuration from a small random (non-symmetric) initial state: with the exactly symmetric zero start
    # above, many gradient entries are pure rounding noise that Adam's 1/sqrt(v) normalisation amplifies (a 1e-16
    # relative perturbation of the gradient moves the loss by 1e-7 after 4 epochs), so that trajectory is only
    # reproducible by an implementation with the identical rounding sequence.  This one is well conditioned.
    for dt, tdt, tag in [(np.float64, torch.float64, "f64"), (np.float32, torch.float32, "f32")]:
        domain, args, ref_u, rhs, terms = poisson_setup(odil, poisson, (256,), 100, dt, seed=5)
        r0 = np.random.default_rng(42)
        terms = [(0.01 * r0.standard_normal(t.shape)).astype(dt) for t in terms]
        extra = make_ns(args=args, rhs=rhs, ref_u=ref_u)
        losses, arrays = run_reference_optimizer(odil, "adam", poisson.operator, domain, extra, terms, tdt, 300,
                                                 lr=0.005)
        out[f"adam_p1d_256_rinit_{tag}_rhs"] = rhs
        out[f"adam_p1d_256_rinit_{tag}_losses"] = losses
        for i, (t0, a) in enumerate(zip(terms, arrays)):
            out[f"adam_p1d_256_rinit_{tag}_init{i}"] = t0
            out[f"adam_p1d_256_rinit_{tag}_x{i}"] = a
    # GD on 1D.
    domain, args, ref_u, rhs, terms = poisson_setup(odil, poisson, (16,), 3, np.float64, seed=5)
    extra = make_ns(args=args, rhs=rhs, ref_u=ref_u)
    losses, arrays = run_reference_optimizer(odil, "gd", poisson.operator, domain, extra, terms, torch.float64, 10,
                                             lr=1e-6)
    out["gd_p1d_16_L3_f64_rhs"] = rhs
    for i, t in enumerate(terms):
        out[f"gd_p1d_16_L3_f64_x0_{i}"] = t
    out["gd_p1d_16_L3_f64_losses"] = losses
    for i, a in enumerate(arrays):
        out[f"gd_p1d_16_L3_f64_x{i}"] = a
    # L-BFGS-B (SciPy) on the wave problem, fp64, 25 iterations, m=50.
    domain, extra, terms = wave_setup(odil, (16, 12), 0, np.float64, seed=7)
    terms = [t * 0 for t in terms]
    losses, arrays = run_reference_optimizer(odil, "lbfgsb", wave.operator_wave, domain, extra, terms, torch.float64,
                                             25, lr=None)
    out["lbfgsb_w_16x12_f64_losses"] = losses
    if arrays is not None:
        out["lbfgsb_w_16x12_f64_x0"] = arrays[0]


def main():
    odil, poisson, wave = import_reference()
    groups = {
        "transfers": lambda o: (gen_interp(odil, o), gen_restrict(odil, o)),
        "poisson": lambda o: gen_poisson(odil, poisson, o),
        "wave": lambda o: gen_wave(odil, wave, o),
        "optim": lambda o: gen_optimizers(odil, poisson, wave, o),
        "wave2": lambda o: gen_wave2(odil, o),
    }
    only = sys.argv[1:]
    for name, fn in groups.items():
        if only and name not in only:
            continue
        out = {}
        fn(out)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, len(out), "arrays ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
