"""Debug driver: one fused launch of the wave footprint on a small grid through k_tile3t (fp32 / fp64)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from odil_b200 import native
from tests.test_zz_wave2_gpu import WAVE2, wrap_free_table
prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
shape = tuple(int(v) for v in sys.argv[2:5]) or (7, 10, 12)
zchunk = int(sys.argv[5]) if len(sys.argv) > 5 else 3
td = torch.float32 if prec == "f32" else torch.float64
rr = (2, 1, 1)
table = wrap_free_table(np.random.default_rng(0), WAVE2, rr)
plan = native.StencilPlan(shape, td, WAVE2, rr, table.reshape(-1, 7))
plan.tune(zchunk=zchunk, variant=80)
U = torch.randn(shape, dtype=td, device="cuda")
c = torch.randn(shape, dtype=td, device="cuda")
G = torch.empty_like(U)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
plan.fused(U, c, 0.5, G, ss)
torch.cuda.synchronize()
print(prec, shape, "ok", float(ss))
