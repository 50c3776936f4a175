#!/usr/bin/env python3
"""
Poisson equation with zero Dirichlet conditions in a cube, solved by ODIL on the B200 backend.
Same structure and flags as the reference's examples/poisson/poisson.py (operator written against
ctx.field / mod.where / mod.roll, odil.util.add_arguments, make_callback, optimize); no plotting.

  python examples/poisson3d.py --ndim 3 --N 256 --nlvl 4 --epochs 200 --report_every 50
  torchrun --nproc-per-node 8 examples/poisson3d.py --ndim 3 --N 512 ...      (slabs along axis 0)
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odil  # noqa: E402
from odil import printlog  # noqa: E402


def neighbours_with_bc(mod, u, um, up, idx, n, wall):
    ex = odil.core.extrap_quadh
    return mod.where(idx == 0, ex(up, u, wall), um), mod.where(idx == n - 1, ex(um, u, wall), up)


def laplacian(mod, fields, h, idx, n, zero):
    u = fields[0]
    res = None
    for a in range(len(h)):
        um, up = neighbours_with_bc(mod, u, fields[1 + 2 * a], fields[2 + 2 * a], idx[a], n[a], zero)
        t = (up - 2 * u + um) / h[a] ** 2
        res = t if res is None else res + t
    return res


def operator(ctx):
    mod, ndim = ctx.mod, ctx.domain.ndim
    h, idx, n = ctx.step(), ctx.indices(), ctx.size()
    if ndim == 1:
        h, idx, n = (h,), (idx,), (n,)
    fields = [ctx.field("u")]
    for a in range(ndim):
        e = [1 if b == a else 0 for b in range(ndim)]
        fields += [ctx.field("u", *[-s for s in e]), ctx.field("u", *e)]
    zero = mod.cast(0, fields[0].dtype)
    return [laplacian(mod, fields, h, idx, n, zero) - ctx.extra.rhs]


def reference_solution(domain):
    xs = domain.points()
    xs = xs if isinstance(xs, tuple) else (xs,)
    u = np.prod([(1 - np.asarray(x)) * np.asarray(x) * 5 for x in xs], axis=0)
    return (u ** 5 / (1 + u ** 5)) ** 0.2


def discrete_rhs(u, domain):
    mod, ndim = domain.mod, domain.ndim
    h, idx, n = domain.step(), domain.indices(), domain.size()
    if ndim == 1:
        h, idx, n = (h,), (idx,), (n,)
    fields = [u]
    for a in range(ndim):
        fields += [mod.roll(u, 1, a), mod.roll(u, -1, a)]
    return laplacian(mod, fields, h, idx, n, mod.cast(0, u.dtype))


def parse_args():
    p = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--ndim", type=int, default=3, choices=[1, 2, 3, 4])
    p.add_argument("--N", type=int, default=64, help="Grid size")
    odil.util.add_arguments(p)
    odil.linsolver.add_arguments(p)
    p.set_defaults(frames=0, report_every=100, history_every=10, plot_every=10 ** 9, optimizer="adam", multigrid=1,
                   lr=0.005, double=0, outdir="out_poisson3d", nlvl=4, epochs=500)
    return p.parse_args()


def main():
    args = parse_args()
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", 1)) > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
        args.outdir += "_rank{}".format(dist.get_rank())
    odil.setup_outdir(args)
    dtype = np.float64 if args.double else np.float32
    domain = odil.Domain(cshape=[args.N] * args.ndim, dimnames=["x", "y", "z", "w"][:args.ndim],
                         multigrid=args.multigrid, mg_nlvl=args.nlvl, dtype=dtype)
    if domain.multigrid:
        printlog("multigrid levels:", domain.mg_cshapes)
    ref_u = reference_solution(domain).astype(dtype)
    extra = argparse.Namespace(ref_u=ref_u, rhs=discrete_rhs(ref_u, domain), args=args)
    state = odil.State()
    state.fields["u"] = None
    state = domain.init_state(state)
    problem = odil.Problem(operator, domain, extra)

    def report(problem, state, epoch, cbinfo):
        u = np.asarray(problem.domain.field(state, "u"))
        printlog("error: u:{:.5g}".format(np.sqrt(np.mean((u - extra.ref_u) ** 2))))

    callback = odil.make_callback(problem, args, report_func=report)
    odil.util.optimize(args, args.optimizer, problem, state, callback)
    printlog("final throughput: {:.1f} Mcells/s".format(getattr(callback.cbinfo, "throughput", 0.0)))


if __name__ == "__main__":
    main()
