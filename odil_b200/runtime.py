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



class _ForeignJit:
    """`from odil.runtime import tf` / `jax`: the reference's example scripts import the backend module and decorate
    helper functions with `@tf.function()` (examples/heat/heat.py:282).  There is no TensorFlow / JAX here; the
    decorator is the identity (functions run eagerly on ModB200 values) and any other attribute fails loudly."""

    def __init__(self, name):
        self._name = name

    def function(self, func=None, **kwargs):
        return func if callable(func) else (lambda f: f)

    jit = function

    def __bool__(self):
        return False

    def __getattr__(self, attr):
        raise AttributeError(f"odil.runtime.{self._name}.{attr}: the B200 backend has no {self._name} module")


tf = _ForeignJit("tf")
jax = _ForeignJit("jax")
mod = ModB200()

dtype_name = os.environ.get("ODIL_DTYPE", "float32")
if dtype_name not in ["float32", "float64"]:
    raise ImportError(f"Expected ODIL_DTYPE=float32 or float64, got '{dtype_name}'")
dtype = numpy.dtype(dtype_name)
