"""
Host twin of the generated kernels -- TEST INFRASTRUCTURE ONLY.

`odil_b200.codegen` emits one source text per traced operator and mode; the same text compiles for the host when
ODIL_HOST is defined (sequential loop over the cells, `+=` in place of atomicAdd).  This module compiles that text
with g++ and runs it on CPU tensors, so the tracer, the index algebra, the reverse- and forward-mode programs and the
parameter block can be checked against reference-generated goldens in the CPU suite.  The product (GraphEngine) only
ever runs the NVRTC-compiled sm_100a build of the same text on the device.
"""
import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import torch

from odil_b200 import codegen
from odil_b200.engine_graph import GraphEngine
from oracle import odil_oracle as orc

_CACHE = os.path.join(tempfile.gettempdir(), "odil_b200_host_twins")


def compile_host(source):
    os.makedirs(_CACHE, exist_ok=True)
    tag = hashlib.sha1(source.encode()).hexdigest()[:16]
    so = os.path.join(_CACHE, f"twin_{tag}.so")
    if not os.path.exists(so):
        cpp = os.path.join(_CACHE, f"twin_{tag}.cpp")
        with open(cpp, "w") as f:
            f.write(source)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DODIL_HOST", "-w", cpp, "-o", so])
    return ctypes.CDLL(so)


class HostTwin:
    """Runs a GraphEngine (built with trace_only=True on a CPU `mod`) through the host build of its kernels."""

    def __init__(self, problem, state):
        self.engine = eng = GraphEngine(problem, state, trace_only=True)
        self.gen = eng.gen
        self.libs = {}
        self.tdtype = eng.tdtype

    def lib(self, mode):
        if mode not in self.libs:
            self.libs[mode] = compile_host(self.engine.source(mode))
        return self.libs[mode]

    def consts(self):
        out = []
        for t, strides, is_bool in self.gen.consts:
            out.append((t.cpu().to(torch.uint8) if is_bool else t.cpu().to(self.tdtype)).contiguous())
        return out

    def inputs(self, arrays):
        """State arrays + regular fields of multigrid unknowns (synthesised with the oracle)."""
        eng = self.engine
        ins = [a.contiguous() for a in arrays]
        for key, slot in eng.trace.regular_slot.items():
            unk = eng.unknowns[key]
            terms = [arrays[unk.first + i].numpy().astype(np.float64) * unk.factors[i] for i in range(unk.narrays)]
            U = orc.mg_synthesize(terms, unk.mgloc)
            ins.append(torch.as_tensor(U, dtype=self.tdtype).contiguous())
        return ins

    def run(self, mode, ins, gin=None, tin=None, out=None, seed=None, sums=None, jcol=None, jval=None, colbase=None,
            only=None):
        eng, gen = self.engine, self.gen
        lib = self.lib(mode)
        keep = self.consts()
        ptr = lambda xs: [x.data_ptr() if x is not None else 0 for x in xs] if xs is not None else []
        for g, name, which in gen.kernels(mode):
            if only is not None and (g.gid, which) != only:
                continue
            params = gen.pack(g.ncell, 1, ptr(ins), ptr(gin), ptr(tin), colbase or [], ptr(keep), ptr(out), ptr(seed),
                              0, sums.data_ptr() if sums is not None else 0,
                              jcol.data_ptr() if jcol is not None else 0, jval.data_ptr() if jval is not None else 0,
                              eng._prm())
            buf = ctypes.create_string_buffer(params, len(params))
            getattr(lib, "h_" + name)(buf)

    def loss_grad(self, arrays):
        eng = self.engine
        ins = self.inputs(arrays)
        gin = [torch.zeros(s, dtype=self.tdtype) for s in eng.trace.shapes]
        sums = torch.zeros(len(eng.outputs), dtype=torch.float64)
        self.run("lossgrad", ins, gin=gin, sums=sums)
        grads = [g.numpy().astype(np.float64) for g in gin[: eng.narrays]]
        for key, slot in eng.trace.regular_slot.items():
            unk = eng.unknowns[key]
            shapes = [tuple(s) for s in unk.shapes]
            gl = orc.mg_adjoint(gin[slot].numpy().astype(np.float64), shapes, unk.mgloc)
            for i in range(unk.narrays):
                grads[unk.first + i] = grads[unk.first + i] + gl[i] * unk.factors[i]
        terms = [float(s) / o.n for s, o in zip(sums.numpy(), eng.outputs)]
        return sum(terms), grads, terms

    def values(self, arrays):
        eng = self.engine
        out = [torch.zeros(o.shape, dtype=self.tdtype) for o in eng.outputs]
        self.run("values", self.inputs(arrays), out=out)
        return [o.numpy() for o in out]

    def jacobian_dense(self, arrays):
        """Dense Jacobian from the 'jac' rows; also checks jvp / vjp against it."""
        eng = self.engine
        sizes = [a.numel() for a in arrays]
        col0 = np.concatenate([[0], np.cumsum(sizes)])
        row0 = np.concatenate([[0], np.cumsum([o.n for o in eng.outputs])])
        J = np.zeros((row0[-1], col0[-1]))
        ins = [a.contiguous() for a in arrays]
        for g, name, which in self.gen.kernels("jac"):
            k = g.results[which][0]
            nl = self.gen.nloads(g)
            if nl == 0:
                continue
            jcol = torch.zeros(g.ncell * nl, dtype=torch.int64)
            jval = torch.zeros(g.ncell * nl, dtype=self.tdtype)
            self.run("jac", ins, jcol=jcol, jval=jval, colbase=[int(c) for c in col0[:-1]], only=(g.gid, which))
            r = np.repeat(np.arange(g.ncell), nl) + row0[k]
            np.add.at(J, (r, jcol.numpy()), jval.numpy().astype(np.float64))
        return J, row0, col0

    def jvp(self, arrays, x):
        eng = self.engine
        sizes = [a.numel() for a in arrays]
        col0 = np.concatenate([[0], np.cumsum(sizes)])
        tin = [torch.as_tensor(x[col0[i]:col0[i + 1]], dtype=self.tdtype).reshape(arrays[i].shape).contiguous()
               for i in range(len(arrays))]
        out = [torch.zeros(o.shape, dtype=self.tdtype) for o in eng.outputs]
        self.run("jvp", [a.contiguous() for a in arrays], tin=tin, out=out)
        return np.concatenate([o.numpy().reshape(-1) for o in out]).astype(np.float64)

    def vjp(self, arrays, y):
        eng = self.engine
        row0 = np.concatenate([[0], np.cumsum([o.n for o in eng.outputs])])
        seed = [torch.as_tensor(y[row0[k]:row0[k + 1]], dtype=self.tdtype).reshape(o.shape).contiguous()
                for k, o in enumerate(eng.outputs)]
        gin = [torch.zeros(a.shape, dtype=self.tdtype) for a in arrays]
        self.run("vjp", [a.contiguous() for a in arrays], gin=gin, seed=seed)
        return np.concatenate([g.numpy().reshape(-1) for g in gin]).astype(np.float64)

    def diagonals(self, arrays):
        """Per group the stored diagonals of mode 'jacd' (D[pair][cell])."""
        ins = [a.contiguous() for a in arrays]
        dia = {}
        for g in self.gen.groups:
            dia[g.gid] = torch.zeros(max(1, len(g.pairs()) * g.ncell), dtype=self.tdtype)
            self.run("jacd", ins, jval=dia[g.gid], only=(g.gid, None))
        return dia

    def jvpd(self, arrays, x, dia):
        eng = self.engine
        sizes = [a.numel() for a in arrays]
        col0 = np.concatenate([[0], np.cumsum(sizes)])
        tin = [torch.as_tensor(x[col0[i]:col0[i + 1]], dtype=self.tdtype).reshape(arrays[i].shape).contiguous()
               for i in range(len(arrays))]
        out = [torch.zeros(o.shape, dtype=self.tdtype) for o in eng.outputs]
        for g in self.gen.groups:
            self.run("jvpd", [a.contiguous() for a in arrays], tin=tin, out=out, jval=dia[g.gid], only=(g.gid, None))
        return np.concatenate([o.numpy().reshape(-1) for o in out]).astype(np.float64)

    def vjpg(self, arrays, y, dia):
        return self.vjpd(arrays, y, dia, mode="vjpg")

    def vjpd(self, arrays, y, dia, mode="vjpd"):
        eng = self.engine
        row0 = np.concatenate([[0], np.cumsum([o.n for o in eng.outputs])])
        seed = [torch.as_tensor(y[row0[k]:row0[k + 1]], dtype=self.tdtype).reshape(o.shape).contiguous()
                for k, o in enumerate(eng.outputs)]
        gin = [torch.zeros(a.shape, dtype=self.tdtype) for a in arrays]
        for g in self.gen.groups:
            self.run(mode, [a.contiguous() for a in arrays], gin=gin, seed=seed, jval=dia[g.gid], only=(g.gid, None))
        return np.concatenate([g.numpy().reshape(-1) for g in gin]).astype(np.float64)
