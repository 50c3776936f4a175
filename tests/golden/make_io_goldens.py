#!/usr/bin/env python3
"""
Writes tests/golden/io/*: the files the UNMODIFIED reference io module (/root/reference/src/odil/io.py, pure
NumPy) produces for a fixed set of inputs, and the train.csv / pickle its `History` class
(/root/reference/src/odil/history.py) leaves after a scripted session.  tests/test_io_cpu.py rebuilds the same inputs (`cases()` below is
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


def history_session(History, outdir):
    """Drives a `History` class through a scripted session: series that first appear in the second entry
    (back-filled with zeros of their kind), skipped values, NumPy and Python scalars, 0-d and 1-element arrays,
    an entry that cannot be closed, a series that arrives after the CSV header, save / load into a second object.
    Returns the files written."""
    csv1, pkl, csv2 = (os.path.join(outdir, n) for n in ("history_train.csv", "history.pickle", "history_reload.csv"))
    h = History(csvpath=csv1, warmup=1)
    for e in range(6):
        h.append("epoch", e)
        h.append("loss", np.float64(1.0) / (e + 1))
        h.append("norm_fu", np.array(np.float32(0.5) ** e))
        if e >= 1:
            h.append("late", np.array([2.5 * e]) if e != 3 else None)  # entry 3 skips it: zero of the series' kind
            h.append_dict({"walltime": np.round(0.1234567 * e, 3), "memory": 100 + e})
        h.write()
    assert h.get("late")[0] == 0.0 and h.get("late")[3] == 0.0 and h.get("nope", 7) == 7 and h.count == 6
    h.save(pkl)
    h.append("epoch", 6)
    for k in ("loss", "norm_fu", "late", "walltime", "memory"):
        h.append(k)
    h.append("too_late", 1.0)  # a series that appears after the header went out
    try:
        h.write()
        raise AssertionError("expected RuntimeError")
    except RuntimeError as err:
        assert str(err) == "Unexpected keys in history: ['too_late']", err
    h.close()
    h2 = History(csvpath=csv2)
    h2.load(pkl)
    h2.close()
    # an entry in which one series got two values cannot be closed
    h3 = History()
    h3.append("a", 1)
    h3.append("b", 2.0)
    h3.write()
    h3.append("a", 3)
    h3.append("a", 4)
    h3.append("b", 5.0)
    try:
        h3.commit()
        raise AssertionError("expected RuntimeError")
    except RuntimeError as err:
        assert str(err) == "Missing values for columns: b,", err
    try:
        h3.append("d", None)  # a new series needs a first value
        raise RuntimeError("expected AssertionError")
    except AssertionError:
        pass
    try:
        h3.append("c", "text")  # a late series is back-filled with zeros, and a zero of kind str does not exist
        raise AssertionError("expected ValueError")
    except ValueError as err:
        assert "Unknown type" in str(err)
    return [csv1, pkl, csv2]


def reference_flags():
    """{module: {flag: [type name, default, choices]}} of the reference's `add_arguments` functions (util.py:70-149,
    linsolver.py:90-131), obtained by executing just those functions from the reference sources."""
    import argparse
    import ast

    out = {}
    for module in ("util", "linsolver"):
        src = open(os.path.join(REF, "src", "odil", module + ".py")).read()
        fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "add_arguments"][0]
        ns = {}
        exec(compile(ast.Module(body=[fn], type_ignores=[]), module, "exec"), ns)
        parser = argparse.ArgumentParser()
        ns["add_arguments"](parser)
        out[module] = describe_flags(parser)
    return out


def describe_flags(parser):
    return {a.dest: [getattr(a.type, "__name__", None), a.default, a.choices and list(a.choices)]
            for a in parser._actions if a.dest != "help"}


def main():
    import json

    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "cli_flags.json"), "w") as f:
        json.dump(reference_flags(), f, indent=1, sort_keys=True)
    spec = importlib.util.spec_from_file_location("ref_io", os.path.join(REF, "src", "odil", "io.py"))
    ref_io = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_io)
    os.makedirs(OUT, exist_ok=True)
    spec = importlib.util.spec_from_file_location("ref_history", os.path.join(REF, "src", "odil", "history.py"))
    ref_history = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_history)
    for p in history_session(ref_history.History, OUT):
        print(p, os.path.getsize(p))
    for name, (kind, kw) in cases().items():
        for p in write_case(ref_io, name, kind, kw, OUT):
            print(p, os.path.getsize(p))


if __name__ == "__main__":
    sys.exit(main())
