#!/usr/bin/env python3
"""
Writes tests/golden/io/*: the files the UNMODIFIED reference io module (/root/reference/src/odil/io.py, pure
NumPy) produces for a fixed set of inputs.  tests/test_io_cpu.py rebuilds the same inputs (`cases()` below is
imported by the test) and requires odil_b200/io.py to write byte-identical files.

Run:   python tests/golden/make_io_goldens.py        (needs /root/reference; NOT run on the GPU box)
"""
import importlib.util
import os
import sys

import numpy as np

REF = os.environ.get("ODIL_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "io")


def cases():
    """name -> (kind, kwargs) with deterministic small inputs."""
    rng = np.random.default_rng(42)
    c = {}
    c["cell3d_f64"] = ("xmf", dict(u=rng.standard_normal((3, 4, 5)), spacing=(0.1, 0.2, 0.3), cell=True, name="u"))
    c["node3d_f32"] = ("xmf", dict(u=rng.standard_normal((2, 3, 4)).astype(np.float32), spacing=(1, 0.5, 0.25),
                                   cell=False, name="rhs"))
    c["cell2d_f64"] = ("xmf", dict(u=rng.standard_normal((6, 7)), spacing=(0.125, 0.0625), cell=True, name=None))
    c["cell1d_f32"] = ("xmf", dict(u=rng.standard_normal((9,)).astype(np.float32), spacing=(0.5,), cell=True,
                                   name="line"))
    pts = rng.standard_normal((6, 3))
    c["poly_ascii"] = ("vtk", dict(points=pts, polygons=[[0, 1, 2], [2, 3, 4, 5]], point_fields={"p": np.arange(6.0)},
                                   cell_fields={"c": np.array([1.5, -2.0])}, comment="two polygons"))
    c["lines_ascii"] = ("vtk", dict(points=pts, lines=[[0, 1], [1, 2, 3], [3, 4, 5, 0]], tcoords=pts[:, :2],
                                    fmt="%.6g"))
    c["lines_binary"] = ("vtk", dict(points=pts, lines=[[0, 1], [1, 2, 3]], point_fields={"q": np.arange(6.0) ** 2},
                                     binary=True, comment="binary"))
    c["points_only"] = ("vtk", dict(points=pts[:2]))
    return c


def write_case(io, name, kind, kw, outdir):
    """Writes one case with the module `io` into `outdir`; returns the list of files."""
    if kind == "xmf":
        kw = dict(kw)
        u = kw.pop("u")
        path = os.path.join(outdir, name + ".xmf")
        io.write_raw_with_xmf(u, path, **kw)
        return [path, os.path.join(outdir, name + ".raw")]
    path = os.path.join(outdir, name + ".vtk")
    kw = dict(kw)
    io.write_vtk_poly(path, kw.pop("points"), **kw)
    return [path]


def main():
    spec = importlib.util.spec_from_file_location("ref_io", os.path.join(REF, "src", "odil", "io.py"))
    ref_io = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_io)
    os.makedirs(OUT, exist_ok=True)
    for name, (kind, kw) in cases().items():
        for p in write_case(ref_io, name, kind, kw, OUT):
            print(p, os.path.getsize(p))


if __name__ == "__main__":
    sys.exit(main())
