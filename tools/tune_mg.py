"""Marching multigrid transfer kernels vs the previous per-coarse-cell kernels: agreement + timing."""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, ".")


def run(old):
    if old:
        os.environ["ODIL_B200_MG_OLD"] = "1"
    from odil_b200 import native
    native.load()

    def timeit(fn, warm=2, rep=7):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(rep):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    out = {}
    g = torch.Generator(device="cuda").manual_seed(3)
    for td in (torch.float32, torch.float64):
        for cshape in [(8, 6, 10), (16, 20, 34), (33, 17, 64), (64, 64, 64)]:
            fshape = tuple(2 * s for s in cshape)
            coarse = torch.randn(cshape, dtype=td, device="cuda", generator=g)
            term = torch.randn(fshape, dtype=td, device="cuda", generator=g)
            o = torch.empty(fshape, dtype=td, device="cuda")
            native.mg_interp_add(cshape, "ccc", coarse, 0.7, term, 1.3, o)
            gc = torch.empty(cshape, dtype=td, device="cuda")
            native.mg_interp_adjoint(cshape, "ccc", term, 0.9, gc)
            out[(str(td), cshape)] = (o.double().cpu().numpy(), gc.double().cpu().numpy())
    for N, td, es in [(512, torch.float32, 4), (256, torch.float32, 4), (128, torch.float32, 4), (256, torch.float64, 8)]:
        half = (N // 2,) * 3
        n = N ** 3
        coarse = torch.randn(half, dtype=td, device="cuda")
        U = torch.randn((N,) * 3, dtype=td, device="cuda")
        G = torch.empty_like(U)
        gc = torch.empty_like(coarse)
        t1 = timeit(lambda: native.mg_interp_add(half, "ccc", coarse, 1.0, U, 1.0, G))
        t2 = timeit(lambda: native.mg_interp_adjoint(half, "ccc", U, 1.0, gc))
        print(f"{'old' if old else 'new'} N={N} {td}: interp_add {t1:.3f} ms {(2+1/8)*es*n/(t1*1e-3)/1e9:.0f} GB/s | "
              f"interp_adjoint {t2:.3f} ms {(1+1/8)*es*n/(t2*1e-3)/1e9:.0f} GB/s", flush=True)
        del coarse, U, G, gc
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1:
        res = run(sys.argv[1] == "old")
        np.save(f"/tmp/mg_{sys.argv[1]}.npy", np.array([res], dtype=object), allow_pickle=True)
    else:
        for w in ("old", "new"):
            subprocess.check_call([sys.executable, __file__, w])
        a = np.load("/tmp/mg_old.npy", allow_pickle=True)[0]
        b = np.load("/tmp/mg_new.npy", allow_pickle=True)[0]
        ok = True
        for key in a:
            for i, nm in enumerate(["interp_add", "interp_adjoint"]):
                err = np.max(np.abs(a[key][i] - b[key][i])) / np.max(np.abs(a[key][i]))
                tol = 1e-5 if "32" in key[0] else 1e-13
                print(f"agree {key} {nm}: {err:.2e}")
                ok = ok and err < tol
        print("MG AGREEMENT", "OK" if ok else "FAILED")
