#!/usr/bin/env python3
"""
Golden vectors for operators that are not affine stencils -> tests/golden/nonaffine.npz.

Every case of tests/nonaffine_cases.py is evaluated by the UNMODIFIED reference: its core.py (Domain / State /
Context.field incl. location changes / eval_neural_net / multigrid_to_regular / restrict_to_coarser) runs the
reference's own operator function under the torch `mod` shim of oracle/ref_shim.py, the loss is assembled as in
core.py:1082-1096 (mean(square(F)), mean(value) for Context.Raw) and torch.autograd stands in for
jax.value_and_grad (core.py:1100).  For the Newton cases the dense Jacobian d(concatenated outputs)/d(packed state)
is taken with torch.autograd.functional.jacobian of the same evaluation (the reference's own `linearize`,
core.py:1113-1217, needs TensorFlow).

  python tests/golden/make_nonaffine_goldens.py       (needs /root/reference; the .npz is committed)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(OUT, "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

REF = os.environ.get("ODIL_REFERENCE", "/root/reference")


def load_by_path(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    odil, _, origin = ref_shim.import_reference(())
    assert origin == REF, origin
    cases = load_by_path(os.path.join(ROOT, "tests", "nonaffine_cases.py"), "nonaffine_cases")
    scripts = {name: load_by_path(os.path.join(REF, rel), "refscript_" + name) for name, rel in cases.SCRIPTS.items()}
    for m in scripts.values():
        assert m.odil is odil
    sys.modules["odil"] = odil  # `import odil` inside the API-only operators of nonaffine_cases resolves to the reference
    out = {}
    for case in cases.CASES:
        for npdt, tdt, tag in [(np.float64, torch.float64, "f64"), (np.float32, torch.float32, "f32")]:
            tm = ref_shim.TorchMod(tdt)
            operator, domain, state, extra, tracers = cases.build(case, odil, tm, npdt, scripts)
            for k, v in list(vars(extra).items()):
                if isinstance(v, np.ndarray):
                    setattr(extra, k, torch.tensor(v, dtype=tdt if v.dtype.kind == "f" else None))
            if hasattr(extra, "ref"):
                extra.ref = {k: torch.tensor(np.asarray(v), dtype=tdt) for k, v in extra.ref.items()}
            shapes = [tuple(a.shape) for a in domain.arrays_from_state(state)]
            rng = np.random.default_rng(1000 + len(case))
            init = [(0.5 + 0.4 * rng.standard_normal(s)).astype(npdt) for s in shapes]

            def evaluate(leaves, want_values=False):
                domain.arrays_to_state(list(leaves), state)
                ctx = odil.core.Context(domain, state, extra=extra, tracers=dict(tracers))
                ff = operator(ctx)
                names = [f[0] if isinstance(f, tuple) else "" for f in ff]
                values = [f[1] if isinstance(f, tuple) else f for f in ff]
                raws = [isinstance(v, odil.core.Context.Raw) for v in values]
                values = [v.value if r else v for v, r in zip(values, raws)]
                values = [torch.as_tensor(v, dtype=tdt) if not torch.is_tensor(v) else v for v in values]
                if want_values:
                    return values, names, raws
                terms = [tm.mean(v) if r else tm.mean(tm.square(v)) for v, r in zip(values, raws)]
                return sum(terms), terms

            leaves = [torch.tensor(a, requires_grad=True) for a in init]
            loss, terms = evaluate(leaves)
            grads = torch.autograd.grad(loss, leaves, allow_unused=True)
            grads = [g if g is not None else torch.zeros_like(x) for g, x in zip(grads, leaves)]
            values, names, raws = evaluate(leaves, want_values=True)
            key = f"{case}_{tag}"
            out[key + "_loss"] = np.asarray(loss.detach())
            out[key + "_terms"] = np.asarray([float(t) for t in terms])
            out[key + "_names"] = np.asarray(names)
            out[key + "_raws"] = np.asarray(raws)
            for i, (a, g) in enumerate(zip(init, grads)):
                out[f"{key}_x{i}"] = a
                out[f"{key}_g{i}"] = g.detach().numpy()
            for i, v in enumerate(values):
                out[f"{key}_F{i}"] = v.detach().numpy()
            if case in cases.NEWTON_CASES and tag == "f64":
                def packed(*xs):
                    vals, _, _ = evaluate(xs, want_values=True)
                    return torch.cat([v.reshape(-1) for v in vals])

                jac = torch.autograd.functional.jacobian(packed, tuple(torch.tensor(a) for a in init))
                nrow = jac[0].shape[0]
                out[key + "_jac"] = np.concatenate([j.reshape(nrow, -1).numpy() for j in jac], axis=1)
            print(key, "loss", float(loss), "outputs", [tuple(v.shape) for v in values], "arrays", shapes)
    path = os.path.join(OUT, "nonaffine.npz")
    np.savez_compressed(path, **out)
    print(len(out), "arrays ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
