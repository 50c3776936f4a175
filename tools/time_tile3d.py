"""Times the fused sweep of a non-star 3-D plan (the configs[2] wave footprint) in the per-cell kernel and in
k_tile3d, CUDA events around 5 launches each.  Usage: python tools/time_tile3d.py [N0 N1 N2]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native

shape = tuple(int(v) for v in sys.argv[1:4]) or (128, 256, 256)
offs = [(0, 0, 0), (-1, 0, 0), (-2, 0, 0), (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1)]
table = np.random.default_rng(0).standard_normal((45, 7))
U = torch.randn(shape, device="cuda")
c = torch.randn(shape, device="cuda")
G = torch.empty_like(U)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
for variant, name in ((81, "k_generic"), (80, "k_tile3d")):
    plan = native.StencilPlan(shape, torch.float32, offs, (2, 1, 1), table)
    plan.tune(variant=variant)
    plan.fused(U, c, 0.5, G, ss)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(5):
        plan.fused(U, c, 0.5, G, ss)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    cells = int(np.prod(shape))
    print(f"{name}: {'x'.join(map(str, shape))} f32 wave footprint: {ms:.3f} ms, {cells / ms / 1e6:.1f} Gcells/s, "
          f"{12 * cells / ms / 1e6:.0f} GB/s algorithmic", flush=True)
