"""Alias of odil_b200.runtime so that `from odil.runtime import mod, dtype, tf` works."""
from odil_b200.runtime import *  # noqa: F401,F403
from odil_b200.runtime import backend_name, dtype, dtype_name, enable_gpu, enable_jit, jax, mod, tf  # noqa: F401
