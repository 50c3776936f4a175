"""Times the multigrid synthesis of level 0 (256^3 coarse -> 512^3 fine, fp32): out = term + I(coarse), default kernel
and the variant that loads the fine term one step ahead (ODIL_B200_ADD_PF=1).  Usage: python tools/time_interp_add.py [n]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from odil_b200 import native

native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
coarse = torch.randn((n,) * 3, dtype=torch.float32, device="cuda")
term = torch.randn((2 * n,) * 3, dtype=torch.float32, device="cuda")
outs = []
for pf in ("0", "1"):
    os.environ["ODIL_B200_ADD_PF"] = pf
    out = torch.empty_like(term)
    for _ in range(3):
        native.mg_interp_add((n,) * 3, "ccc", coarse, 1.0, term, 1.0, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        native.mg_interp_add((n,) * 3, "ccc", coarse, 1.0, term, 1.0, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = (2 * term.numel() + coarse.numel()) * 4
    outs.append(out)
    print(f"ODIL_B200_ADD_PF={pf}: {ms:.4f} ms, {nbytes / ms / 1e6:.0f} GB/s = {nbytes / ms / 1e6 / 6450.3:.3f} of measured peak",
          flush=True)
print("bit-identical:", bool(torch.equal(outs[0], outs[1])))
