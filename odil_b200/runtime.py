"""
Process-wide configuration read from the environment at import (reference src/odil/runtime.py).

  ODIL_BACKEND  b200 (default and only compute backend of this package)
  ODIL_DTYPE    float32 (default) | float64
  ODIL_JIT      kept for compatibility (the operator is always traced once; nothing to toggle)
  ODIL_WARN     1 shows Python warnings
Unlike the reference this module never calls exit(): a bad value raises.
"""
import os
import warnings

import numpy

from .backend import ModB200

if not int(os.environ.get("ODIL_WARN", 0)):
    warnings.simplefilter(action="ignore", category=FutureWarning)

enable_jit = bool(int(os.environ.get("ODIL_JIT", 0)))
enable_gpu = os.environ.get("CUDA_VISIBLE_DEVICES", "") not in ["-1"]
backend_name = os.environ.get("ODIL_BACKEND", "") or "b200"
if backend_name != "b200":
    raise ImportError(f"Unknown ODIL_BACKEND='{backend_name}', options are: b200")

tf = None
jax = None
mod = ModB200()

dtype_name = os.environ.get("ODIL_DTYPE", "float32")
if dtype_name not in ["float32", "float64"]:
    raise ImportError(f"Expected ODIL_DTYPE=float32 or float64, got '{dtype_name}'")
dtype = numpy.dtype(dtype_name)
