"""Traces the device L-BFGS run of configs[2] (bench.py's workload): after every function evaluation one line with the
iteration, bitwise checksums (int64 wrap-around sums of the bit patterns) of x and of the returned gradient, and the
loss.  Two runs (different builds, ODIL_B200_TILE3T=0 / 1, ...) are compared line by line to find the first operation
whose result differs.  Usage: python tools/trace_lbfgs.py out.txt [iterations] [size]"""
import argparse
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import odil
from odil_b200 import lbfgs
from tests import operators as ops

out, iters = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 33
N = int(sys.argv[3]) if len(sys.argv) > 3 else 512
problem, state = ops.make_wave2((N // 2, N, N), np.float32)
lines = []
orig = lbfgs.minimize


def bits(t):
    return int(t.view(torch.int64).sum().item())


def traced(func, x0, **kw):
    n = [0]

    def f2(x):
        f, g = func(x)
        n[0] += 1
        lines.append(f"eval {n[0]:3d}  x {bits(x):22d}  f {f!r:24s}  g {bits(g):22d}")
        return f, g

    cb = kw.pop("callback", None)

    def cb2(x):
        lines.append(f"iter        x {bits(x):22d}")
        if cb is not None:
            cb(x)

    return orig(f2, x0, callback=cb2, **kw)


lbfgs.minimize = traced
args = argparse.Namespace(epochs=iters, epoch_start=0, lr=0.005, callback_update_state=0, bfgs_m=50, bfgs_pgtol=None,
                          bfgs_maxls=None, adam_epsilon=None, adam_beta_1=None, adam_beta_2=None)
try:
    odil.util.optimize_grad(args, "lbfgsb", problem, state, lambda *a: None)
except odil.EarlyStopError:
    pass
open(out, "w").write("\n".join(lines) + "\n")
print(len(lines), "lines ->", out)
