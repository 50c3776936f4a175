"""Times the fused sweep (k_star8) of the 3-D Poisson plan at N^3 fp32, CUDA events around 20 launches.
Usage: [ODIL_B200_S8_ASYNC=1] [ODIL_B200_SPIN_NS=..] python tools/time_star8.py [N]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native
from oracle import odil_oracle as orc

native.load()
torch.manual_seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
offsets, table, rr = orc.poisson_plan(3, [np.float32(1) / np.float32(N)] * 3)
plan = native.StencilPlan((N,) * 3, torch.float32, offsets, rr, table)
U = torch.randn((N,) * 3, device="cuda")
c = torch.randn((N,) * 3, device="cuda")
G = torch.empty_like(U)
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
for _ in range(5):
    plan.fused(U, c, 2.0 / N ** 3, G, ss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    plan.fused(U, c, 2.0 / N ** 3, G, ss)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
import hashlib

sha = hashlib.sha1(G.cpu().numpy().tobytes()).hexdigest()[:16]
print(f"ODIL_B200_S8_ASYNC={os.environ.get('ODIL_B200_S8_ASYNC', '0')} ODIL_B200_SPIN_NS={os.environ.get('ODIL_B200_SPIN_NS', '0')}: {ms:.4f} ms, {12 * N ** 3 / ms / 1e6:.0f} GB/s = "
      f"{12 * N ** 3 / ms / 1e6 / 6450.3:.3f} of measured peak, sum F^2 {float(ss):.8e}, G sha1 {sha}")
