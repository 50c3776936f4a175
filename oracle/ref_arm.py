#!/usr/bin/env python3
"""
Reference arm of bench.py: times the reference's OWN CPU implementation of the benchmark epoch on the host cores.
TEST/BENCH INFRASTRUCTURE, NOT PRODUCT.  Run as a separate process (bench.py spawns it) so that the reference's
`odil` package and the repository's `odil` alias never meet in one interpreter.

What runs per epoch (all of it unmodified reference code, see oracle/ref_shim.py for how it is driven without
JAX/TF):  Domain.multigrid_to_regular / interp_to_finer "stack" (core.py:245-263, :606-700)  ->  Context.field = roll
(core.py:910-975)  ->  examples/poisson/poisson.py:operator (:89-123)  ->  mean(square(F)) (core.py:1093)  ->
gradient w.r.t. every multigrid term (torch.autograd in place of jax.value_and_grad, core.py:1100)  ->
optimizer.py AdamNativeOptimizer.run (:286-341), whose callback timestamps delimit the epochs.

  python oracle/ref_arm.py --size 512 --levels 4 --dtype f32 --steps 3 --warmup 1 [--threads N|0=auto] [--ndim 3]
Prints ONE JSON line: Mcells/s (from the mean epoch time over the timed epochs), min / median epoch, threads, origin.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def build_problem(odil, poisson, shim, cshape, levels, dtype):
    import torch

    npdt = np.float32 if dtype == "f32" else np.float64
    tdt = torch.float32 if dtype == "f32" else torch.float64
    tm = shim.TorchMod(tdt)
    ndim = len(cshape)
    domain = odil.Domain(cshape=list(cshape), dimnames=["x", "y", "z"][:ndim], lower=0.0, upper=1.0, dtype=npdt,
                         multigrid=levels > 0, mg_nlvl=levels if levels > 0 else None, mod=tm)
    gen = torch.Generator().manual_seed(0)
    rhs = torch.randn(tuple(cshape), dtype=tdt, generator=gen)   # rhs ~ N(0,1), as in the B200 arm
    extra = argparse.Namespace(args=argparse.Namespace(mgloss=0), rhs=rhs)
    cshapes = domain.mg_cshapes if levels > 0 else [tuple(cshape)]
    x0 = [torch.zeros(tuple(cs), dtype=tdt) for cs in cshapes]   # unknowns start at 0 (poisson.py:265)

    def state_builder(leaves):
        state = odil.State()
        if levels > 0:
            terms = [odil.Field(t, loc="c" * ndim, cshape=tuple(cs)) for t, cs in zip(leaves, cshapes)]
            state.fields["u"] = odil.MultigridField(terms=terms, loc="c" * ndim, factors=[1] * len(terms))
        else:
            state.fields["u"] = odil.Field(leaves[0], loc="c" * ndim, cshape=tuple(cshape))
        state = domain.init_state(state)
        domain.arrays_to_state(leaves, state)
        return state

    def loss_grad(arrays):
        leaves = [a.detach().requires_grad_(True) for a in arrays]
        loss, grads, terms, names, _ = shim.reference_loss_grad(odil, poisson.operator, domain, state_builder, leaves,
                                                                 extra)
        return loss, grads, {"loss": loss}

    return domain, tm, npdt, x0, loss_grad


def time_epochs(odil, tm, npdt, x0, loss_grad, steps, warmup, lr):
    import odil.optimizer as ropt

    stamps = [time.perf_counter()]
    last = {}

    def callback(arrays, epoch, pinfo):
        last["loss"] = float(pinfo["loss"])
        stamps.append(time.perf_counter())

    opt = ropt.make_optimizer("adam", dtype=npdt, mod=tm)
    opt.run(x0, loss_grad, epochs=warmup + steps, callback=callback, lr=lr)
    per = np.diff(stamps)[warmup:]
    return per, last["loss"]


def pick_threads(odil, poisson, shim, ndim, levels, dtype, candidates):
    """Thread count that serves the reference best on this host (small probe grid)."""
    import torch

    n = 96 if ndim == 3 else 512
    best = None
    for nthr in candidates:
        torch.set_num_threads(nthr)
        domain, tm, npdt, x0, loss_grad = build_problem(odil, poisson, shim, (n,) * ndim, min(levels, 4), dtype)
        per, _ = time_epochs(odil, tm, npdt, x0, loss_grad, 2, 1, 0.005)
        if best is None or per.min() < best[0]:
            best = (per.min(), nthr)
    return best[1]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--size", type=int, default=512)
    p.add_argument("--ndim", type=int, default=3)
    p.add_argument("--levels", type=int, default=4)
    p.add_argument("--dtype", type=str, default="f32")
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=1)
    p.add_argument("--lr", type=float, default=0.005)
    p.add_argument("--threads", type=int, default=0, help="0: probe 8/16/32/64 (capped at the core count)")
    args = p.parse_args()
    import torch

    from oracle import ref_shim as shim

    odil, mods, origin = shim.import_reference(("poisson",))
    poisson = mods["poisson"]
    ncpu = os.cpu_count()
    if args.threads > 0:
        threads = args.threads
    else:
        threads = pick_threads(odil, poisson, shim, args.ndim, args.levels, args.dtype,
                               sorted({min(ncpu, c) for c in (8, 16, 32, 64)}))
    torch.set_num_threads(threads)
    cshape = (args.size,) * args.ndim
    domain, tm, npdt, x0, loss_grad = build_problem(odil, poisson, shim, cshape, args.levels, args.dtype)
    per, loss = time_epochs(odil, tm, npdt, x0, loss_grad, args.steps, args.warmup, args.lr)
    cells = float(np.prod(cshape))
    cpu_model = ""
    try:
        with open("/proc/cpuinfo") as f:
            cpu_model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "")
    except OSError:
        pass
    print(json.dumps({
        "mcells_per_s": cells / per.mean() / 1e6, "ms_per_step": per.mean() * 1e3, "ms_min": per.min() * 1e3,
        "ms_median": float(np.median(per)) * 1e3, "steps": args.steps, "warmup": args.warmup, "threads": threads,
        "host_cores": ncpu, "cpu_model": cpu_model, "cshape": list(cshape), "levels": args.levels,
        "dtype": args.dtype, "origin": origin, "final_loss": loss,
    }), flush=True)


if __name__ == "__main__":
    main()
