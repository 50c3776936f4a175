"""Kernel-level timing on one GPU (CUDA events, L2-exceeding inputs). Scratch tool for tuning."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from oracle import odil_oracle as orc

native.load()


def timeit(fn, warm=3, rep=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    res = {}
    for prec, td, nd in [("f32", torch.float32, np.float32), ("f64", torch.float64, np.float64)]:
        shape = (N, N, N) if prec == "f32" else (N // 2, N // 2, N // 2)
        n = int(np.prod(shape))
        es = 4 if prec == "f32" else 8
        steps = [nd(1) / nd(s) for s in shape]
        offsets, table, rr = orc.poisson_plan(3, steps)
        plan = native.StencilPlan(shape, td, offsets, rr, table)
        U = torch.randn(shape, dtype=td, device="cuda")
        c = torch.randn(shape, dtype=td, device="cuda")
        G = torch.empty_like(U)
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        for variant in [0, 1, 2, 3, 20, 12]:
            for zchunk in [0, 32, 64, 128]:
                plan.tune(zchunk=zchunk, variant=variant)
                med, mn = timeit(lambda: plan.fused(U, c, 2.0 / n, G, ss))
                gbs = 3 * es * n / (mn * 1e-3) / 1e9
                res[f"fused_{prec}_v{variant}_z{zchunk}"] = dict(ms=mn, med=med, GBs=gbs, Gcells=n / mn / 1e6)
                print(f"fused {prec} N={shape[0]} variant={variant} zchunk={zchunk}: {mn:.3f} ms  {gbs:.0f} GB/s (alg 3s B/cell)", flush=True)
        # generic kernels for context
        F = torch.empty_like(U)
        med, mn = timeit(lambda: plan.forward(U, c, F), rep=3)
        print(f"generic forward {prec}: {mn:.3f} ms {3*es*n/(mn*1e-3)/1e9:.0f} GB/s")
        med, mn = timeit(lambda: plan.adjoint(F, 1.0, None, G), rep=3)
        print(f"generic adjoint {prec}: {mn:.3f} ms {2*es*n/(mn*1e-3)/1e9:.0f} GB/s")
        # multigrid transfers
        half = tuple(s // 2 for s in shape)
        coarse = torch.randn(half, dtype=td, device="cuda")
        med, mn = timeit(lambda: native.mg_interp_add(half, "ccc", coarse, 1.0, U, 1.0, G))
        print(f"interp_add {prec}: {mn:.3f} ms {(2+1/8)*es*n/(mn*1e-3)/1e9:.0f} GB/s")
        gc = torch.empty_like(coarse)
        med, mn = timeit(lambda: native.mg_interp_adjoint(half, "ccc", U, 1.0, gc))
        print(f"interp_adjoint {prec}: {mn:.3f} ms {(1+1/8)*es*n/(mn*1e-3)/1e9:.0f} GB/s")
        m = torch.zeros_like(U)
        v = torch.zeros_like(U)
        med, mn = timeit(lambda: native.adam_step([U], [m], [v], [G], 1e-3, 0.1, 0.001, 1e-7))
        print(f"adam {prec}: {mn:.3f} ms {7*es*n/(mn*1e-3)/1e9:.0f} GB/s")
        med, mn = timeit(lambda: G.copy_(U))
        print(f"torch copy {prec}: {mn:.3f} ms {2*es*n/(mn*1e-3)/1e9:.0f} GB/s")
        del U, c, G, F, m, v, coarse, gc
    json.dump(res, open("gpurun_out/bench_kernels.json", "w"), indent=1)


main()
