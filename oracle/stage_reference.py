#!/usr/bin/env python3
"""
Stages the UNMODIFIED reference into oracle/_ref/ so that the reference arm of bench.py can run on the GPU box
(where /root/reference does not exist).  TEST/BENCH INFRASTRUCTURE, NOT PRODUCT.

The sanctioned route (`pip install --target ... /root/reference`) fails in this image: the reference builds with
hatchling, which is neither installed nor in /opt/wheelhouse, and there is no network.  A pure-Python package
installs as a verbatim copy of its files, so this recipe does exactly what the wheel would: it copies
src/odil/*.py (the package), the example scripts (the operator of the benchmark configuration, which the package
does not ship, and the operators of the parity cases) and the reference's own test scripts into oracle/_ref/, and writes their SHA-256 digests next to them.  oracle/_ref/ is
git-ignored (nothing of the reference enters the history) but not gpurun-ignored, like a built .so.
oracle/reference_manifest.json (committed) pins the digests: oracle/ref_shim.py refuses files that differ.

  python oracle/stage_reference.py            (needs /root/reference; run by __graft_entry__.build())
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ODIL_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(HERE, "reference_manifest.json")
FILES = {
    "odil": "src/odil",                                   # directory: every *.py and the style file
    "examples/poisson/poisson.py": "examples/poisson/poisson.py",
    "examples/wave/wave.py": "examples/wave/wave.py",
    "examples/heat/heat.py": "examples/heat/heat.py",
    "examples/heat_tmax/heat_tmax.py": "examples/heat_tmax/heat_tmax.py",
    "examples/infer_constant/infer_constant.py": "examples/infer_constant/infer_constant.py",
    "examples/velocity_from_tracer/veltracer.py": "examples/velocity_from_tracer/veltracer.py",
    "examples/basic/fields.py": "examples/basic/fields.py",
    # the reference's own test scripts: run unmodified against this package by tests/test_reference_scripts_gpu.py
    "tests/test_optimize.py": "tests/test_optimize.py",
    "tests/test_newton.py": "tests/test_newton.py",
    "tests/test_domain.py": "tests/test_domain.py",
    "tests/test_mg_interp.py": "tests/test_mg_interp.py",
    "tests/test_mg_restrict.py": "tests/test_mg_restrict.py",
}


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def digests(root):
    out = {}
    for base, _, names in os.walk(root):
        for n in sorted(names):
            if n.endswith(".pyc") or n == "DIGESTS.json":
                continue
            p = os.path.join(base, n)
            out[os.path.relpath(p, root)] = sha256(p)
    return dict(sorted(out.items()))


def stage(write_manifest=False):
    if not os.path.isdir(REF):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for dst, src in FILES.items():
        s, d = os.path.join(REF, src), os.path.join(DST, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copyfile(s, d)
    dig = digests(DST)
    with open(os.path.join(DST, "DIGESTS.json"), "w") as f:
        json.dump(dig, f, indent=1)
    if write_manifest or not os.path.exists(MANIFEST):
        with open(MANIFEST, "w") as f:
            json.dump({"reference": "cselab/odil 0.1.8 (pyproject.toml)", "sha256": dig}, f, indent=1)
    else:
        with open(MANIFEST) as f:
            pinned = json.load(f)["sha256"]
        if pinned != dig:
            raise RuntimeError("staged reference files differ from oracle/reference_manifest.json")
    return True


if __name__ == "__main__":
    ok = stage(write_manifest="--write-manifest" in sys.argv)
    print("staged" if ok else f"{REF} not present; nothing staged", DST)
