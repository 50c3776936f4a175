"""
Parity bookkeeping for the GPU tests: every comparison against a reference-generated golden (or the oracle) goes
through `check(name, err, tol)`, which prints the MEASURED error next to its bar and records it; at the end of the
session the table is written to gpurun_out/parity_errors.json (copied to profiles/ per round) so the bars in the
tests can be justified by what the kernels actually achieve.

fp32 bars: the reference's own fp32 run differs from its fp64 run by 0.5e-7 .. 3e-7 (loss, relative) and
0.4e-7 .. 3e-7 (gradient, relative to max|g|) on the golden cases (tests/golden/*.npz hold both precisions of the
same inputs), i.e. a few fp32 ulps.  Measured on B200 (profiles/r02_parity_errors_call1.json): the kernels are within
4.3e-7 of the reference's fp32 gradients and 1.2e-7 of its fp32 losses on every golden case, i.e. inside the
reference's own spread.  The bars below are ~5x the measured worst case.
"""
import json
import os

import numpy as np

F32_LOSS = 1e-6      # relative error of the loss
F32_GRAD = 2e-6      # max |g - g_ref| / max |g_ref|
F32_FIELD = 2e-6     # same measure for residual fields / synthesised U
F64_LOSS = 1e-11
F64_GRAD = 1e-11

RECORD = {}


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check(name, err, tol):
    err = float(err)
    prev = RECORD.get(name)
    RECORD[name] = {"err": max(err, prev["err"]) if prev else err, "tol": float(tol)}
    print(f"parity {name}: measured {err:.3e} (bar {tol:.1e})")
    assert err < tol, f"{name}: measured {err:.3e} exceeds the bar {tol:.1e}"


def dump():
    if not RECORD:
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump(RECORD, f, indent=1, sort_keys=True)
    except OSError:
        pass
