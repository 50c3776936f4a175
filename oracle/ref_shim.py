"""
Drives the UNMODIFIED reference (cselab/odil: core.py Domain / Context / multigrid transfers, the example operators,
optimizer.py) on the host CPU.  TEST/BENCH INFRASTRUCTURE, NOT PRODUCT: imported only by tests/golden/make_goldens.py,
oracle/ref_arm.py (bench.py's reference arm and cpu_baseline leg, run in a separate process) and tests.

The reference's own array backends are JAX and TensorFlow; neither is installable in this image (no wheels, no
network; runtime.py:29-44 exits without them).  Its code is backend-agnostic through the `mod` namespace
(backend.py:12-47), so the same reference statements run under

  * `odil.backend.ModNumpy()`  -- the reference's own NumPy namespace: forward values only, and
  * `TorchMod` (below)         -- the ~40 callables core.py and the operators touch, mapped onto torch-CPU, with
                                  `torch.autograd` standing in for `jax.value_and_grad` (core.py:1100).

Where the reference comes from: /root/reference when it exists (the build container), else the staged copy under
oracle/_ref/ (oracle/stage_reference.py), whose files are verified against oracle/reference_manifest.json.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def locate_reference():
    """(package parent dir, examples dir, origin) of the reference sources."""
    ref = os.environ.get("ODIL_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(ref, "src", "odil")):
        return os.path.join(ref, "src"), os.path.join(ref, "examples"), ref
    staged = os.path.join(HERE, "_ref")
    if os.path.isdir(os.path.join(staged, "odil")):
        with open(os.path.join(HERE, "reference_manifest.json")) as f:
            pinned = json.load(f)["sha256"]
        for rel, digest in pinned.items():
            with open(os.path.join(staged, rel), "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != digest:
                    raise RuntimeError(f"oracle/_ref/{rel} differs from the pinned reference file")
        return staged, os.path.join(staged, "examples"), "oracle/_ref (staged copy, digests verified)"
    raise RuntimeError("reference sources not found: neither /root/reference nor oracle/_ref (run "
                       "oracle/stage_reference.py in the build container)")


def import_reference(examples=("poisson",)):
    """
    Imports the reference package as `odil` (this process must not have imported the repository's own `odil` alias
    package) plus the named example modules.  Stubs what the image lacks: matplotlib, odil.plotutil, odil.runtime
    (which would exit for want of TF/JAX).  Returns (odil, {example name: module}, origin).
    """
    import importlib.util

    pkg_parent, ex_dir, origin = locate_reference()
    if "odil" in sys.modules and not os.path.abspath(sys.modules["odil"].__file__).startswith(
            os.path.abspath(pkg_parent)):
        raise RuntimeError("the repository's `odil` package is already imported in this process")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.path.insert(0, pkg_parent)
    mpl = _stub_module("matplotlib", use=lambda *a, **k: None)
    mpl.style = types.SimpleNamespace(use=lambda *a, **k: None)
    _stub_module("matplotlib.pyplot")
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    import odil

    assert os.path.abspath(odil.__file__).startswith(os.path.abspath(pkg_parent)), odil.__file__
    _stub_module("odil.plotutil")
    fake_tf = types.SimpleNamespace(function=lambda f=None, **k: (f if f is not None else (lambda g: g)))
    _stub_module("odil.runtime", tf=fake_tf, jax=None, mod=None, dtype=np.dtype("float64"), enable_jit=False,
                 backend_name="numpy", dtype_name="float64", enable_gpu=False)
    odil.runtime = sys.modules["odil.runtime"]
    mods = {}
    for name in examples:
        path = {"poisson": "poisson/poisson.py", "wave": "wave/wave.py", "heat": "heat/heat.py"}[name]
        spec = importlib.util.spec_from_file_location("odil_example_" + name, os.path.join(ex_dir, path))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return odil, mods, origin


class TorchMod:
    """`mod` namespace over torch-CPU: only the callables the reference core.py / operators / optimizer.py touch."""
    jax = None
    tf = None
    modsp = None

    def __init__(self, dtype=torch.float64):
        self.tdtype = dtype
        self.float32 = np.float32
        self.float64 = np.float64
        self.random = types.SimpleNamespace(
            set_seed=lambda s: torch.manual_seed(s),
            uniform=lambda shape, minval, maxval, dtype: (minval + (maxval - minval) * torch.rand(
                tuple(shape), dtype=torch.float64)).to(self._tt(dtype)))

    @staticmethod
    def _tt(dtype):
        if isinstance(dtype, torch.dtype):
            return dtype
        return {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
                np.dtype("int64"): torch.int64, np.dtype("int32"): torch.int32}[np.dtype(dtype)]

    def cast(self, x, dtype):
        tt = self._tt(dtype)
        if torch.is_tensor(x):
            return x.to(tt)
        # Python/NumPy scalars are rounded ONCE to the target dtype (as jnp.array(x, dtype) does).
        return torch.as_tensor(np.asarray(x), dtype=tt) if not isinstance(x, (int, float)) else torch.tensor(x, dtype=tt)

    array = staticmethod(lambda x, dtype=None: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x))
    constant = staticmethod(lambda x: torch.as_tensor(x))

    def variable(self, x, dtype=None):
        t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
        return t.to(self._tt(dtype)) if dtype is not None else t

    def zeros(self, shape, dtype=None):
        return torch.zeros(tuple(int(s) for s in np.atleast_1d(shape)), dtype=self._tt(dtype or np.float64))

    zeros_like = staticmethod(torch.zeros_like)
    ones_like = staticmethod(torch.ones_like)
    copy = staticmethod(lambda x: x.clone())
    is_tensor = staticmethod(torch.is_tensor)
    stop_gradient = staticmethod(lambda x: x.detach())
    mean = staticmethod(torch.mean)
    sum = staticmethod(torch.sum)
    square = staticmethod(torch.square)
    sqrt = staticmethod(lambda x: torch.sqrt(torch.as_tensor(x)))
    exp = staticmethod(torch.exp)
    tanh = staticmethod(torch.tanh)
    sigmoid = staticmethod(torch.sigmoid)
    relu = staticmethod(torch.relu)
    abs = staticmethod(torch.abs)
    max = staticmethod(torch.max)
    stack = staticmethod(lambda xs, axis=0: torch.stack([torch.as_tensor(x) for x in xs], dim=axis))
    reshape = staticmethod(lambda x, shape: torch.reshape(torch.as_tensor(x), tuple(int(s) for s in shape)))
    flatten = staticmethod(lambda x: torch.reshape(x, (-1,)))
    concatenate = staticmethod(lambda xs, axis=0: torch.cat([torch.as_tensor(x) for x in xs], dim=axis))
    transpose = staticmethod(lambda x, perm: x.permute(*[int(p) for p in perm]))
    matmul = staticmethod(torch.matmul)

    sin = staticmethod(torch.sin)
    cos = staticmethod(torch.cos)
    log = staticmethod(torch.log)

    @staticmethod
    def convolution(input, filters, strides, padding):
        """Stand-in for jax.lax.conv as core.py:751 calls it (cross-correlation, VALID, one stride for all axes)."""
        assert padding == "VALID"
        nd = input.dim()
        conv = {1: torch.nn.functional.conv1d, 2: torch.nn.functional.conv2d, 3: torch.nn.functional.conv3d}[nd]
        w = torch.as_tensor(filters, dtype=input.dtype)
        return conv(input[None, None], w[None, None], stride=int(strides))[0, 0]

    @staticmethod
    def where(c, a, b):
        c = torch.as_tensor(c)
        ref = a if torch.is_tensor(a) else b
        if not torch.is_tensor(a):
            a = torch.as_tensor(a, dtype=ref.dtype)
        if not torch.is_tensor(b):
            b = torch.as_tensor(b, dtype=ref.dtype)
        return torch.where(c, a, b)

    @staticmethod
    def roll(x, shift, axis=None):
        x = torch.as_tensor(x)
        if np.ndim(shift) == 0:
            return torch.roll(x, int(shift), int(axis))
        return torch.roll(x, [int(s) for s in shift], [int(a) for a in axis])

    @staticmethod
    def meshgrid(*xx, indexing="ij"):
        return torch.meshgrid(*[torch.as_tensor(x) for x in xx], indexing=indexing)

    @staticmethod
    def pad(x, pad_width, mode):
        # numpy.pad semantics for the modes core.py uses: reflect, symmetric, constant (zero).
        for ax, (lo, hi) in enumerate(pad_width):
            if lo == 0 and hi == 0:
                continue
            n = x.shape[ax]
            if mode == "constant":
                shp = list(x.shape)
                parts = []
                if lo:
                    shp[ax] = lo
                    parts.append(torch.zeros(shp, dtype=x.dtype))
                parts.append(x)
                if hi:
                    shp[ax] = hi
                    parts.append(torch.zeros(shp, dtype=x.dtype))
                x = torch.cat(parts, dim=ax)
                continue
            assert lo <= 1 and hi <= 1
            if mode == "reflect":
                left, right = [1], [n - 2]
            elif mode == "symmetric":
                left, right = [0], [n - 1]
            else:
                raise ValueError(mode)
            idx = (left if lo else []) + list(range(n)) + (right if hi else [])
            x = torch.index_select(x, ax, torch.as_tensor(idx))
        return x


def reference_loss_grad(odil, operator, domain, state_builder, leaves, extra, tracers=None):
    """
    One evaluation the way the reference's jitted function does it (core.py:1082-1104): Context -> operator ->
    terms = mean(square(F_k)) -> loss = sum -> gradient w.r.t. every state array (autograd for jax.value_and_grad).
    `leaves`: torch tensors with requires_grad; `state_builder(leaves)` returns the initialised reference State.
    """
    state = state_builder(leaves)
    ctx = odil.core.Context(domain, state, extra=extra, tracers=tracers or {"epoch": 0})
    ff = operator(ctx)
    names = [f[0] if isinstance(f, tuple) else "" for f in ff]
    values = [f[1] if isinstance(f, tuple) else f for f in ff]
    mod = domain.mod
    terms = [mod.mean(v.value) if isinstance(v, odil.core.Context.Raw) else mod.mean(mod.square(v)) for v in values]
    loss = sum(terms)
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(x) for g, x in zip(grads, leaves)]
    return loss.detach(), grads, [t.detach() for t in terms], names, values
