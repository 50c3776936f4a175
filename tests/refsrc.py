"""
Where the tests find the UNMODIFIED reference scripts (examples/*.py, tests/*.py): /root/reference in the build
container, else the staged copy oracle/_ref (oracle/stage_reference.py; travels to the GPU box, digests pinned).
`load(relpath)` executes one of them as a module against THIS repository's `odil` package -- the scripts import
`odil`, matplotlib and (some) `from odil.runtime import tf`; matplotlib is absent from the image and is replaced by an
inert stand-in.
"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def root():
    ref = os.environ.get("ODIL_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref, "examples")):
        return ref
    staged = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(os.path.join(staged, "examples")):
        return staged
    return None


def available():
    return root() is not None


class _Inert(types.ModuleType):
    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return lambda *a, **k: None


def load(relpath, name=None):
    import odil

    for modname in ["matplotlib", "matplotlib.pyplot"]:
        if modname not in sys.modules:
            m = _Inert(modname)
            m.__file__ = os.devnull
            sys.modules[modname] = m
    path = os.path.join(root(), relpath)
    name = name or "refscript_" + os.path.splitext(os.path.basename(relpath))[0]
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    assert module.odil is odil, "the script must have imported this repository's odil package"
    return module
