"""Times the fused 2-D sweep in k_tile2d (ODIL_B200_TILE2W=0) and in the variants of k_tile2w (ODIL_B200_TILE2W=2;
rows in flight, warps per CTA, shared-memory carveout, rows per chunk) on the two 2-D workloads of the reference's
examples: the 5-point Poisson star at 1024 x 1024 (BASELINE configs[1]) and the wave footprint at 2048 x 4096
(examples/wave/wave.py at scale).  20 launches captured in a CUDA graph and replayed (no host time between launches),
CUDA events around the replay; second figure: single launches behind a 512 MB L2 flush.
Usage: python tools/time_tile2w.py [quick]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from tests.test_tile_emulation_cpu import wrap_free_table

STAR = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]
WAVE = [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)]
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
burn = torch.randn(4096, 4096, device="cuda")
VARIANTS = [("k_tile2d", dict(ODIL_B200_TILE2W="0"))]
for pf, warps, carve, rows in [(2, 4, 1, 0), (2, 4, 0, 0), (2, 4, 1, 4), (2, 4, 1, 8), (2, 4, 1, 12), (2, 4, 1, 16), (2, 4, 1, 24),
                               (4, 4, 1, 0), (2, 8, 1, 0), (4, 4, 1, 8)]:
    env = dict(ODIL_B200_TILE2W="2", ODIL_B200_T2W_PF=str(pf), ODIL_B200_T2W_WARPS=str(warps),
               ODIL_B200_T2W_CARVE=str(carve))
    if rows:
        env["ODIL_B200_T2W_ROWS"] = str(rows)
    VARIANTS.append((f"k_tile2w pf{pf} w{warps} carve{carve} rows{rows or 'auto'}", env))
CASES = [((1024, 1024), STAR, (1, 1), "poisson star", torch.float32), ((2048, 4096), WAVE, (2, 1), "wave footprint", torch.float32)]
if "quick" not in sys.argv:
    CASES += [((2048, 4096), WAVE, (2, 1), "wave footprint", torch.float64), ((4096, 4096), STAR, (1, 1), "poisson star", torch.float32)]
for _ in range(200):  # bring the clocks up
    burn @ burn
for shape, offs, rr, label, td in CASES:
    s = 4 if td == torch.float32 else 8
    tshape = tuple(2 * r + 1 for r in rr) + (len(offs),)
    table = wrap_free_table(np.random.default_rng(0).standard_normal(tshape), offs, rr).reshape(-1, len(offs))
    U = torch.randn(shape, device="cuda", dtype=td)
    c = torch.randn(shape, device="cuda", dtype=td)
    G = torch.empty_like(U)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    cells = int(np.prod(shape))
    ref = None
    for name, env in VARIANTS:
        for k in [k for k in os.environ if k.startswith("ODIL_B200_T2W_")]:
            del os.environ[k]
        os.environ.update(env)
        plan = native.StencilPlan(shape, td, offs, rr, table)
        plan.fused(U, c, 0.5, G, ss)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                plan.fused(U, c, 0.5, G, ss)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        torch.cuda.synchronize()
        t0.record()
        for _ in range(5):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        warm = t0.elapsed_time(t1) / 100
        cold = []
        for _ in range(5):
            flush.zero_()
            t0.record()
            plan.fused(U, c, 0.5, G, ss)
            t1.record()
            torch.cuda.synchronize()
            cold.append(t0.elapsed_time(t1))
        cold = min(cold)
        same = ""
        if ref is None:
            ref = G.clone()
        else:
            same = f", G identical to k_tile2d: {bool(torch.equal(ref, G))}"
        gbs = 3 * s * cells / 1e6
        print(f"{shape[0]}x{shape[1]} {label} f{8 * s} {name}: replayed {warm * 1e3:.1f} us = {gbs / warm:.0f} GB/s algorithmic "
              f"({gbs / warm / 6450.3:.3f} of measured HBM peak); after an L2 flush {cold * 1e3:.1f} us = "
              f"{gbs / cold / 6450.3:.3f}{same}", flush=True)
