"""Times the fused 2-D sweep in k_tile2d (ODIL_B200_TILE2W=0) and k_tile2w (ODIL_B200_TILE2W=2) on the two 2-D workloads
of the reference's examples: the 5-point Poisson star at 1024 x 1024 (BASELINE configs[1]) and the wave footprint at
2048 x 4096 (examples/wave/wave.py at scale), fp32 and fp64.  CUDA events around 50 back-to-back launches (L2-resident, as
in the epoch) and around single launches behind a 512 MB L2 flush.  Usage: python tools/time_tile2w.py"""
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
for shape, offs, rr, label in [((1024, 1024), STAR, (1, 1), "poisson star"), ((2048, 4096), WAVE, (2, 1), "wave footprint"),
                               ((4096, 4096), STAR, (1, 1), "poisson star")]:
    for td, s in ((torch.float32, 4), (torch.float64, 8)):
        tshape = tuple(2 * r + 1 for r in rr) + (len(offs),)
        table = wrap_free_table(np.random.default_rng(0).standard_normal(tshape), offs, rr).reshape(-1, len(offs))
        U = torch.randn(shape, device="cuda", dtype=td)
        c = torch.randn(shape, device="cuda", dtype=td)
        G = torch.empty_like(U)
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        cells = int(np.prod(shape))
        ref = None
        for name, flag in (("k_tile2d", "0"), ("k_tile2w", "2")):
            os.environ["ODIL_B200_TILE2W"] = flag
            plan = native.StencilPlan(shape, td, offs, rr, table)
            plan.fused(U, c, 0.5, G, ss)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record()
            for _ in range(50):
                plan.fused(U, c, 0.5, G, ss)
            t1.record()
            torch.cuda.synchronize()
            warm = t0.elapsed_time(t1) / 50
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
            print(f"{name}: {shape[0]}x{shape[1]} {label} f{8 * s}: back to back {warm * 1e3:.1f} us = {gbs / warm:.0f} GB/s "
                  f"algorithmic ({gbs / warm / 6450.3:.3f} of measured HBM peak); after an L2 flush {cold * 1e3:.1f} us = "
                  f"{gbs / cold / 6450.3:.3f}; sum F^2 {float(ss):.6e}{same}", flush=True)
