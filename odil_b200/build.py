"""
Builds libodil_b200.so (sm_100a) in-tree with nvcc.  `python -m odil_b200.build` or
`__graft_entry__.build()`.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libodil_b200.so")
SOURCES = ["stencil.cu", "multigrid.cu", "optim.cu", "jit.cu", "comm.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I" + os.path.join(ROOT, "include"),
]


def find_nvcc():
    for cand in [os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"]:
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "odil_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = find_nvcc()
    # ODIL_B200_LEGACY=1 also compiles the superseded generations of the fused sweep (k_star7, k_star_tma, k_star3d:
    # 36 kernel instantiations kept as measured history, selectable with plan_tune); off by default.
    legacy = ["-DODIL_B200_LEGACY"] if os.environ.get("ODIL_B200_LEGACY", "0") not in ("", "0") else []
    # extra defines for A/B builds, e.g. ODIL_B200_DEFINES="-DODIL_B200_NO_FFMA2" with ODIL_B200_LIB_OUT=<path>
    legacy += os.environ.get("ODIL_B200_DEFINES", "").split()
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + legacy + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    out = os.environ.get("ODIL_B200_LIB_OUT") or LIB
    cmd = [nvcc, "-shared", "-o", out] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
