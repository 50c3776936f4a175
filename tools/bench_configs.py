"""
Secondary configurations of BASELINE.json measured through the public API (not bench lines; the headline is
bench.py): configs[1] 2-D Poisson 1024^2 / 3 levels / Adam / fp32, configs[2]-like wave inverse with the device
L-BFGS, configs[2] itself (wave in two space dimensions, `3` on the command line; not in the default set),
configs[4]-like Newton + CG on a Poisson system, plus the raw L-BFGS building blocks.
Usage: python tools/bench_configs.py
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import odil
from odil_b200 import linsolver, native
from tests import operators as ops
from tests.test_api_gpu import run_args

native.load()


def sync_time(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def config1(shape=(1024, 1024), nlvl=3, n=500):
    import os
    cells = int(np.prod(shape))
    for flag in ["0", "1"]:
        os.environ["ODIL_B200_GRAPH"] = flag
        problem, state = ops.make_poisson(shape, nlvl, np.float32)
        args = run_args(epochs=50, lr=0.005)
        odil.util.optimize_grad(args, "adam", problem, state, None)  # warm-up (trace, plans)
        args = run_args(epochs=n, lr=0.005)
        dt = sync_time(lambda: odil.util.optimize_grad(args, "adam", problem, state, None), n)
        how = "one CUDA graph per epoch" if flag == "1" else "eager launches"
        print(f"configs[1] Poisson {'x'.join(map(str, shape))}, {nlvl}-level multigrid, Adam, f32, {how}: "
              f"{dt*1e6:.1f} us/epoch, {cells/dt/1e6:.0f} Mcells/s", flush=True)
    os.environ["ODIL_B200_GRAPH"] = "0"


def config2(nt=2048, nx=4096, iters=20):
    problem, state = ops.make_wave((nt, nx), 0, np.float32)
    args = run_args(epochs=3, bfgs_m=50)
    try:
        odil.util.optimize_grad(args, "lbfgsb", problem, state, None)
    except odil.EarlyStopError:
        pass
    args = run_args(epochs=iters, bfgs_m=50)

    def run():
        try:
            odil.util.optimize_grad(args, "lbfgsb", problem, state, None)
        except odil.EarlyStopError:
            pass

    dt = sync_time(run, iters)
    print(f"configs[2]-like wave inverse (t,x) = {nt}x{nx} ({nt*nx/1e6:.1f} M unknowns), device L-BFGS m=50, f32: "
          f"{dt*1e3:.2f} ms/iteration, {nt*nx/dt/1e6:.0f} Mcells/s", flush=True)


def config3(shape=(256, 512, 512), iters=10):
    """BASELINE configs[2] at full size: (t, x, y) = 256 x 512 x 512 fp32 (67 M unknowns), device L-BFGS m=50.
    History: 2 x 50 x 67 M x 8 B = 54 GB of the 180 GB; the sweep runs in the per-cell kernel (not a star)."""
    problem, state = ops.make_wave2(shape, np.float32)
    args = run_args(epochs=2, bfgs_m=50)

    def run(a):
        try:
            odil.util.optimize_grad(a, "lbfgsb", problem, state, None)
        except odil.EarlyStopError:
            pass

    run(args)
    dt = sync_time(lambda: run(run_args(epochs=iters, bfgs_m=50)), iters)
    cells = int(np.prod(shape))
    print(f"configs[2] wave (t,x,y) = {'x'.join(map(str, shape))} ({cells/1e6:.1f} M unknowns), device L-BFGS m=50, "
          f"f32: {dt*1e3:.2f} ms/iteration, {cells/dt/1e6:.0f} Mcells/s", flush=True)


def lbfgs_blocks(n=64 * 1024 * 1024, k=100):
    V = torch.randn(k, n, dtype=torch.float64, device="cuda")
    g = torch.randn(n, dtype=torch.float64, device="cuda")
    d = torch.empty_like(g)
    out = torch.zeros(k, dtype=torch.float64, device="cuda")
    coef = torch.randn(k, dtype=torch.float64, device="cuda")
    for name, fn, nbytes in [("multi_dot", lambda: native.multi_dot(V, k, g, out), (k + 1) * n * 8),
                             ("multi_axpy", lambda: native.multi_axpy(V, k, coef, 1.0, g, d), (k + 2) * n * 8)]:
        fn()
        dt = sync_time(lambda: [fn() for _ in range(3)], 3)
        print(f"{name}: k={k} rows x {n/1e6:.0f} M fp64: {dt*1e3:.2f} ms, {nbytes/dt/1e9:.0f} GB/s", flush=True)


def config4(N=512):
    problem, state = ops.make_poisson((N, N), 0, np.float64)
    vector, matrix = problem.linearize(state)
    status = {}
    args = argparse.Namespace(linsolver_tol=0.0, linsolver_maxiter=200, linsolver_damp=0)
    linsolver.solve(matrix, -vector, args, status, "cg_b200")
    dt = sync_time(lambda: linsolver.solve(matrix, -vector, args, status, "cg_b200"), 200)
    print(f"configs[4]-like Newton system, 2-D Poisson {N}^2 f64, matrix-free CG on the normal equations: "
          f"{dt*1e6:.1f} us/iteration ({status['niter']} iterations)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "blocks", "4"]
    if "1" in which:
        config1()
        config1((128, 128, 128), 3, 300)
    if "2" in which:
        config2()
    if "3" in which:
        config3()
    if "blocks" in which:
        lbfgs_blocks()
    if "4" in which:
        config4()
