"""k_star7 vs the previous kernels: agreement of g and sum F^2 on random data, then a tile / z-chunk timing sweep."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from oracle import odil_oracle as orc

native.load()


def timeit(fn, warm=2, rep=7):
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


def check(shape, td, nd, ref_variant=20):
    n = int(np.prod(shape))
    steps = [nd(1) / nd(s) for s in shape]
    offsets, table, rr = orc.poisson_plan(len(shape), steps)
    plan = native.StencilPlan(shape, td, offsets, rr, table)
    g = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(shape, dtype=td, device="cuda", generator=g)
    c = torch.randn(shape, dtype=td, device="cuda", generator=g) * 100
    G0, G1 = torch.empty_like(U), torch.empty_like(U)
    s0 = torch.zeros(1, dtype=torch.float64, device="cuda")
    s1 = torch.zeros(1, dtype=torch.float64, device="cuda")
    plan.tune(zchunk=0, variant=ref_variant)
    plan.fused(U, c, 2.0 / n, G0, s0)
    out = []
    for variant in [50, 51, 52, 60, 62]:
        for zc in [0, 7, 13]:
            plan.tune(zchunk=zc, variant=variant)
            G1.fill_(float("nan"))
            try:
                plan.fused(U, c, 2.0 / n, G1, s1)
            except Exception as e:
                print(f"check {tuple(shape)} {td} variant={variant} zchunk={zc}: LAUNCH FAILED {e}", flush=True)
                out.append((variant, zc, 1.0, 1.0))
                continue
            torch.cuda.synchronize()
            err = float((G1 - G0).abs().max() / G0.abs().max())
            serr = abs(float(s1) - float(s0)) / abs(float(s0))
            out.append((variant, zc, err, serr))
            print(f"check {tuple(shape)} {td} variant={variant} zchunk={zc}: g rel err {err:.3e}, sumsq rel err {serr:.3e}", flush=True)
    return out


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    ok = True
    for shape, td, nd in [((64, 40, 136), torch.float32, np.float32), ((33, 50, 260), torch.float64, np.float64),
                          ((200, 264), torch.float32, np.float32), ((256, 256, 256), torch.float32, np.float32)]:
        for variant, zc, err, serr in check(shape, td, nd):
            tol = 2e-5 if td == torch.float32 else 1e-12
            if not (err < tol and serr < tol):
                ok = False
    print("AGREEMENT", "OK" if ok else "FAILED", flush=True)
    res = {}
    for prec, td, nd, shape in [("f32", torch.float32, np.float32, (N, N, N)), ("f64", torch.float64, np.float64, (N // 2, N // 2, N))]:
        n = int(np.prod(shape))
        es = 4 if prec == "f32" else 8
        steps = [nd(1) / nd(s) for s in shape]
        offsets, table, rr = orc.poisson_plan(3, steps)
        plan = native.StencilPlan(shape, td, offsets, rr, table)
        U = torch.randn(shape, dtype=td, device="cuda")
        c = torch.randn(shape, dtype=td, device="cuda")
        G = torch.empty_like(U)
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        for variant, zcs in [(50, [0, 64, 128, 256]), (51, [0]), (52, [0]), (60, [0]), (30, [32])]:
            for zchunk in zcs:
                plan.tune(zchunk=zchunk, variant=variant)
                try:
                    med, mn = timeit(lambda: plan.fused(U, c, 2.0 / n, G, ss))
                except Exception as e:
                    print(f"fused {prec} variant={variant} zchunk={zchunk}: FAILED {e}", flush=True)
                    continue
                gbs = 3 * es * n / (med * 1e-3) / 1e9
                res[f"fused_{prec}_v{variant}_z{zchunk}"] = dict(ms_min=mn, ms_med=med, GBs=gbs)
                print(f"fused {prec} {shape} variant={variant} zchunk={zchunk}: med {med:.3f} min {mn:.3f} ms  {gbs:.0f} GB/s", flush=True)
        del U, c, G
    json.dump(res, open("gpurun_out/tune_star7.json", "w"), indent=1)


main()
