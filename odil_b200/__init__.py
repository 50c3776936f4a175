"""
odil_b200 -- B200-native (sm_100a) residual-and-gradient engine behind the ODIL Python API.

Host-side mirror of cselab/odil's Domain / operator(ctx) / Problem / optimizer surface
(reference src/odil/__init__.py:1-61) on top of hand-written CUDA kernels reached through a ctypes
C ABI (include/odil_b200.h -> odil_b200/lib/libodil_b200.so).  `import odil` (the thin alias package
at the repository root) exposes the same names, so problem scripts written for the reference run
unchanged with ODIL_BACKEND=b200 (the default and only backend).
"""
# ruff: noqa: F401
__version__ = "0.1.0"

from . import backend, core, history, io, linsolver, native, optimizer, plotutil, util
from .backend import ModB200, ModBase, ModNumpy, ModTensorflow, NonAffineError
from .core import (
    Array,
    Context,
    Domain,
    Field,
    MultigridField,
    NeuralNet,
    Problem,
    State,
    interp_to_finer,
    restrict_to_coarser,
)
from .history import History
from .io import parse_raw_xmf, read_raw, read_raw_with_xmf, write_raw_with_xmf, write_raw_xmf, write_vtk_poly
from .optimizer import EarlyStopError
from .util import make_callback, optimize, printlog, set_log_file, setup_outdir


def __getattr__(name):
    # `runtime` reads the environment at import time, like the reference (lazy: __init__.py:46-61).
    if name == "runtime":
        import importlib

        return importlib.import_module(".runtime", __name__)
    raise AttributeError(name)
