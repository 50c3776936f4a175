"""Launches the fused 2-D sweep twice per kernel (k_tile2d, k_tile2w) on two workloads; meant to run under
`ncu --set full -k regex:k_tile2 ...`.  Usage: python tools/prof_tile2w.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from tests.test_tile_emulation_cpu import wrap_free_table

STAR = [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)]
WAVE = [(0, 0), (-1, 0), (-2, 0), (-1, -1), (-1, 1)]
for shape, offs, rr in [((2048, 4096), WAVE, (2, 1)), ((1024, 1024), STAR, (1, 1))]:
    tshape = tuple(2 * r + 1 for r in rr) + (len(offs),)
    table = wrap_free_table(np.random.default_rng(0).standard_normal(tshape), offs, rr).reshape(-1, len(offs))
    U = torch.randn(shape, device="cuda")
    c = torch.randn(shape, device="cuda")
    G = torch.empty_like(U)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    for flag in ("0", "2"):
        os.environ["ODIL_B200_TILE2W"] = flag
        plan = native.StencilPlan(shape, torch.float32, offs, rr, table)
        for _ in range(2):
            plan.fused(U, c, 0.5, G, ss)
        torch.cuda.synchronize()
