"""Runs the epoch kernels of the 512^3 fp32 benchmark a few times (for ncu). Usage: python tools/profile_epoch.py [N]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from oracle import odil_oracle as orc

native.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
td, nd = torch.float32, np.float32
shape = (N, N, N)
n = N ** 3
offsets, table, rr = orc.poisson_plan(3, [nd(1) / nd(N)] * 3)
plan = native.StencilPlan(shape, td, offsets, rr, table)
U = torch.randn(shape, dtype=td, device="cuda")
c = torch.randn(shape, dtype=td, device="cuda")
G = torch.empty_like(U)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
half = (N // 2,) * 3
coarse = torch.randn(half, dtype=td, device="cuda")
gc = torch.empty_like(coarse)
m = torch.zeros_like(U)
v = torch.zeros_like(U)
T0 = torch.randn(shape, dtype=td, device="cuda")
for _ in range(3):
    native.mg_interp_add(half, "ccc", coarse, 1.0, T0, 1.0, U)
    plan.fused(U, c, 2.0 / n, G, ss)
    native.mg_interp_adjoint(half, "ccc", G, 1.0, gc)
    native.adam_step([T0], [m], [v], [G], 1e-3, 0.1, 0.001, 1e-7)
torch.cuda.synchronize()
print("done", ss.item())
