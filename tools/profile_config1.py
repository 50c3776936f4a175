"""
Launch list of a few Adam epochs of the secondary configurations through the public API (run under
`ncu --metrics gpu__time_duration.sum --clock-control none`): which kernels the small-grid epochs consist of.
Usage: python tools/profile_config1.py [2d|3d|wave]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import odil
from tests import operators as ops
from tests.test_api_gpu import run_args

which = sys.argv[1] if len(sys.argv) > 1 else "2d"
if which == "wave":
    problem, state = ops.make_wave((2048, 4096), 0, np.float32)
    args = run_args(epochs=3, bfgs_m=50)
    try:
        odil.util.optimize_grad(args, "lbfgsb", problem, state, None)
    except odil.EarlyStopError:
        pass
else:
    shape = (1024, 1024) if which == "2d" else (128, 128, 128)
    problem, state = ops.make_poisson(shape, 3, np.float32)
    odil.util.optimize_grad(run_args(epochs=4, lr=0.005), "adam", problem, state, None)
torch.cuda.synchronize()
