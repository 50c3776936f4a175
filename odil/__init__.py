"""
`import odil` compatibility alias: existing ODIL problem scripts (`import odil`, `odil.Domain`,
`odil.core.extrap_quadh`, `from odil.runtime import mod`, `odil.util.optimize` ...) resolve to the
B200-native package `odil_b200`.  Nothing lives here but the aliasing.
"""
import importlib
import sys

import odil_b200 as _impl
from odil_b200 import *  # noqa: F401,F403
from odil_b200 import (  # noqa: F401
    Array, Context, Domain, EarlyStopError, Field, History, ModB200, MultigridField, NeuralNet, NonAffineError,
    Problem, State, interp_to_finer, make_callback, optimize, parse_raw_xmf, printlog, read_raw, read_raw_with_xmf,
    restrict_to_coarser, set_log_file, setup_outdir, write_raw_with_xmf, write_raw_xmf, write_vtk_poly,
)

for _name in ["backend", "core", "history", "io", "plotutil", "linsolver", "native", "optimizer", "util", "engine"]:
    sys.modules[__name__ + "." + _name] = importlib.import_module("odil_b200." + _name)
    globals()[_name] = sys.modules[__name__ + "." + _name]


class _LazyModule:
    """`odil.runtime` / `from odil.runtime import mod` import odil_b200.runtime on first touch."""

    def __init__(self, target):
        self._target = target

    def __getattr__(self, item):
        return getattr(importlib.import_module(self._target), item)


def __getattr__(name):
    if name == "runtime":
        mod = importlib.import_module("odil_b200.runtime")
        sys.modules[__name__ + ".runtime"] = mod
        return mod
    if name == "plot":
        raise AttributeError("odil.plot (plot_1d / plot_2d figure layouts) is outside the B200 hot-path build")
    raise AttributeError(name)


__version__ = _impl.__version__
