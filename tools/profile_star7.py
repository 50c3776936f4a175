"""Runs the fused star kernel a few times (for ncu). Usage: python tools/profile_star7.py N variant zchunk [variant zchunk ...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from oracle import odil_oracle as orc

native.load()
N = int(sys.argv[1])
cfgs = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(2, len(sys.argv) - 1, 2)]
td, nd = torch.float32, np.float32
shape = (N, N, N)
n = N ** 3
offsets, table, rr = orc.poisson_plan(3, [nd(1) / nd(N)] * 3)
plan = native.StencilPlan(shape, td, offsets, rr, table)
U = torch.randn(shape, dtype=td, device="cuda")
c = torch.randn(shape, dtype=td, device="cuda")
G = torch.empty_like(U)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
for variant, zchunk in cfgs:
    plan.tune(zchunk=zchunk, variant=variant)
    for _ in range(2):
        plan.fused(U, c, 2.0 / n, G, ss)
torch.cuda.synchronize()
print("done", ss.item())
