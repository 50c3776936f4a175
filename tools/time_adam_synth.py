"""Times odil_b200_adam_synth (Adam of the finest term + synthesis of level 0) against the pair it replaces
(odil_b200_adam_step + odil_b200_mg_interp_add) at 256^3 -> 512^3, fp32."""
import sys

import torch

sys.path.insert(0, ".")
from odil_b200 import native

native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
f = (2 * n,) * 3
coarse = torch.randn((n,) * 3, device="cuda")
x, m, g, out = (torch.randn(f, device="cuda") for _ in range(4))
v = torch.rand(f, device="cuda")


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cells = x.numel()
t_f = timeit(lambda: native.adam_synth((n,) * 3, "ccc", coarse, 1.0, 1.0, x, m, v, g, out, 1e-3, 0.1, 0.001, 1e-7))
t_a = timeit(lambda: native.adam_step([x], [m], [v], [g], 1e-3, 0.1, 0.001, 1e-7))
t_i = timeit(lambda: native.mg_interp_add((n,) * 3, "ccc", coarse, 1.0, x, 1.0, out))
print(f"adam_synth: {t_f:.4f} ms = {8.125 * 4 * cells / t_f / 1e6:.0f} GB/s ({8.125 * 4 * cells / t_f / 1e6 / 6450.3:.3f} of measured "
      f"peak); adam_step {t_a:.4f} ms + mg_interp_add {t_i:.4f} ms = {t_a + t_i:.4f} ms")
