"""Times the transposed interpolation of level 0 (256^3 coarse -> 512^3 fine, fp32): k_interp_adjoint3t (TMA-fed, the
default) and k_interp_adjoint3m (ODIL_B200_ADJ_TMA=0; occupancy variant ODIL_B200_ADJ_OCC = 4 | 5 | 6), each followed
by k_adjoint_joint_fix.  Usage: python tools/time_adjoint.py [n] [tma|ldg]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from odil_b200 import native

native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2:] or ["tma", "ldg"]
g = torch.randn((2 * n,) * 3, dtype=torch.float32, device="cuda")
outs = []
for name in which:
    os.environ["ODIL_B200_ADJ_TMA"] = "1" if name == "tma" else "0"
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
    outs.append(gc)
    print(f"{name} (ODIL_B200_ADJ_OCC={os.environ.get('ODIL_B200_ADJ_OCC', '4')}): {ms:.4f} ms incl. the joint fix, "
          f"{nbytes / ms / 1e6:.0f} GB/s = {nbytes / ms / 1e6 / 6450.3:.3f} of measured peak, "
          f"checksum {float(gc.double().sum()):.6e}", flush=True)
if len(outs) == 2:
    print("bit-identical:", bool(torch.equal(outs[0], outs[1])))
