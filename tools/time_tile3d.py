"""Times the fused sweep of a non-star 3-D plan (the configs[2] wave footprint, wrap-free table) in k_tile3t (TMA-fed,
default), k_tile3d (ODIL_B200_TILE3T=0), k_tile3d8 (ODIL_B200_TILE3D8=1) and, on small grids, the per-cell kernel.
CUDA events around 10 launches each.  Usage: python tools/time_tile3d.py [N0 N1 N2]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native

shape = tuple(int(v) for v in sys.argv[1:4]) or (128, 256, 256)
offs = [(0, 0, 0), (-1, 0, 0), (-2, 0, 0), (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1)]
rr = (2, 1, 1)
table = np.random.default_rng(0).standard_normal((5, 3, 3, 7))
for cls in np.ndindex(5, 3, 3):
    for o, off in enumerate(offs):
        for a in range(3):
            ci, r, d = cls[a], rr[a], off[a]
            if (ci < r and ci + d < 0) or (ci > r and d > 2 * r - ci):
                table[cls + (o,)] = 0.0
table = table.reshape(45, 7)
U = torch.randn(shape, device="cuda")
c = torch.randn(shape, device="cuda")
G = torch.empty_like(U)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
cells = int(np.prod(shape))
runs = [("k_tile3t", {"ODIL_B200_TILE3T": "1"}, 80), ("k_tile3d", {"ODIL_B200_TILE3T": "0"}, 80)]
if cells <= 2 ** 24:
    runs.append(("k_generic", {"ODIL_B200_TILE3T": "0"}, 81))
ref = None
for name, env, variant in runs:
    os.environ.update(env)
    plan = native.StencilPlan(shape, torch.float32, offs, rr, table)
    plan.tune(variant=variant)
    plan.fused(U, c, 0.5, G, ss)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(10):
        plan.fused(U, c, 0.5, G, ss)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    same = ""
    if ref is None:
        ref = G.clone()
    else:
        same = f", G identical to k_tile3t: {bool(torch.equal(ref, G))}"
    print(f"{name}: {'x'.join(map(str, shape))} f32 wave footprint: {ms:.3f} ms, {cells / ms / 1e6:.1f} Gcells/s, "
          f"{12 * cells / ms / 1e6:.0f} GB/s algorithmic = {12 * cells / ms / 1e6 / 6450.3:.3f} of measured peak, "
          f"sum F^2 {float(ss):.6e}{same}", flush=True)
