"""Runs the hot kernels a few times (for ncu). Usage: python tools/profile_fused.py [N] [variant] [zchunk]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from oracle import odil_oracle as orc

native.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 2
zchunk = int(sys.argv[3]) if len(sys.argv) > 3 else 64
td, nd = torch.float32, np.float32
shape = (N, N, N)
n = N ** 3
offsets, table, rr = orc.poisson_plan(3, [nd(1) / nd(N)] * 3)
plan = native.StencilPlan(shape, td, offsets, rr, table)
plan.tune(zchunk=zchunk, variant=variant)
U = torch.randn(shape, dtype=td, device="cuda")
c = torch.randn(shape, dtype=td, device="cuda")
G = torch.empty_like(U)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
half = (N // 2,) * 3
coarse = torch.randn(half, dtype=td, device="cuda")
gc = torch.empty_like(coarse)
m = torch.zeros_like(U)
v = torch.zeros_like(U)
for _ in range(3):
    plan.fused(U, c, 2.0 / n, G, ss)
    native.mg_interp_add(half, "ccc", coarse, 1.0, U, 1.0, G)
    native.mg_interp_adjoint(half, "ccc", U, 1.0, gc)
    native.adam_step([U], [m], [v], [G], 1e-3, 0.1, 0.001, 1e-7)
torch.cuda.synchronize()
print("done", ss.item())
