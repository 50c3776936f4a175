"""Determinism / race stress of the TMA-fed 3-D tile kernel at the configs[2] size: the same input swept 12 times with
other traffic in between (L2 flush, a large matmul); every output must equal the first one bit for bit, and the output
of k_tile3d (ODIL_B200_TILE3T=0).  Usage: python tools/stress_tile3t.py [N0 N1 N2]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from odil_b200 import native

shape = tuple(int(v) for v in sys.argv[1:4]) or (256, 512, 512)
offs = [(0, 0, 0), (-1, 0, 0), (-2, 0, 0), (-1, -1, 0), (-1, 1, 0), (-1, 0, -1), (-1, 0, 1)]
rr = (2, 1, 1)
table = np.random.default_rng(0).standard_normal((5, 3, 3, 7))
for cls in np.ndindex(5, 3, 3):
    for o, off in enumerate(offs):
        for a in range(3):
            ci, r, d = cls[a], rr[a], off[a]
            if (ci < r and ci + d < 0) or (ci > r and d > 2 * r - ci):
                table[cls + (o,)] = 0.0
table = table.reshape(45, 7)
torch.manual_seed(0)
U = torch.randn(shape, device="cuda")
c = torch.randn(shape, device="cuda")
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
A = torch.randn(4096, 4096, device="cuda")
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
outs = {}
for name, flag in (("k_tile3t", "1"), ("k_tile3d", "0")):
    os.environ["ODIL_B200_TILE3T"] = flag
    plan = native.StencilPlan(shape, torch.float32, offs, rr, table)
    ref, sref, bad = None, None, 0
    for it in range(12 if flag == "1" else 2):
        G = torch.full(shape, float("nan"), device="cuda")
        if it % 3 == 1:
            flush.zero_()
        if it % 3 == 2:
            A @ A
        plan.fused(U, c, 0.5, G, ss)
        torch.cuda.synchronize()
        if ref is None:
            ref, sref = G, float(ss)
        else:
            same = bool(torch.equal(ref, G)) and float(ss) == sref
            bad += 0 if same else 1
            if not same:
                print(f"{name}: sweep {it} differs in {(ref != G).sum().item()} cells, sum F^2 {float(ss)!r} vs {sref!r}")
    print(f"{name}: {bad} differing sweeps, sum F^2 {sref!r}, NaNs in G: {torch.isnan(ref).sum().item()}")
    outs[name] = ref
print("k_tile3t == k_tile3d:", bool(torch.equal(outs["k_tile3t"], outs["k_tile3d"])))
