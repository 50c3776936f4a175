"""Times the transposed interpolation of level 0 (256^3 coarse -> 512^3 fine, fp32) for the occupancy variant chosen
with ODIL_B200_ADJ_OCC (4 | 5 | 6).  Usage: for o in 4 5 6; do ODIL_B200_ADJ_OCC=$o python tools/time_adjoint.py; done"""
import os
import sys

import torch

sys.path.insert(0, ".")
from odil_b200 import native

native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.randn((2 * n,) * 3, dtype=torch.float32, device="cuda")
gc = torch.empty((n,) * 3, dtype=torch.float32, device="cuda")
for _ in range(3):
    native.mg_interp_adjoint((n,) * 3, "ccc", g, 1.0, gc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    native.mg_interp_adjoint((n,) * 3, "ccc", g, 1.0, gc)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = (g.numel() + gc.numel()) * 4
print(f"ODIL_B200_ADJ_OCC={os.environ.get('ODIL_B200_ADJ_OCC', '4')}: {ms:.4f} ms, {nbytes / ms / 1e6:.0f} GB/s, "
      f"checksum {float(gc.double().sum()):.6e}")
