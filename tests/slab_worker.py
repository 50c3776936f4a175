"""
Multi-rank worker (launched by tests/test_slab_gpu.py under torchrun): evaluates loss + gradients of the
Poisson and wave problems on slab-decomposed grids and checks them against the undecomposed evaluation
on the same device, then runs a few Adam epochs both ways.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    backend = "nccl" if torch.cuda.is_available() else "gloo"
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    dist.init_process_group(backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    import odil
    from odil_b200.optimizer import adam_scalars
    from odil_b200 import native
    from tests import operators as ops

    def relerr(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))

    worst = 0.0
    for maker, cshape, nlvl in [(ops.make_poisson, (32 * world, 16, 24), 3), (ops.make_poisson, (16 * world, 12, 8), 0),
                                (ops.make_poisson, (16 * world, 16), 2), (ops.make_wave, (16 * world, 12), 0),
                                (ops.make_wave, (16 * world, 8), 2)]:
        for dt in (np.float64, np.float32):
            tol = 1e-11 if dt == np.float64 else 5e-4
            # decomposed (the wave stencil reaches 2 planes back in time: halo 4)
            os.environ["ODIL_HALO"] = "4" if maker is ops.make_wave else "2"
            problem, state = maker(cshape, nlvl, dt)
            domain = problem.domain
            assert domain.slab is not None and domain.slab.world == world
            rng = np.random.default_rng(7)
            if nlvl > 0:
                shapes = [tuple(cs) for cs in domain.mg_cshapes]
            else:
                shapes = [tuple(cshape)]
            terms = [rng.standard_normal(s).astype(dt) for s in shapes]
            arrays = [domain.slab.scatter(domain.mod.variable(t, dtype=dt)) for t in terms]
            domain.arrays_to_state(arrays, state)
            loss, grads, _, names, _ = problem.eval_loss_grad(state)
            loss = float(loss)
            g_glob = [domain.slab.gather(g) for g in grads]
            U_glob = domain.field(state, list(state.fields)[0]).full()
            # undecomposed on every rank
            problem1, state1 = maker(cshape, nlvl, dt)
            d1 = problem1.domain
            d1.slab = None
            st = odil.State()
            key = list(state.fields)[0]
            st.fields[key] = np.zeros(cshape, dtype=dt)
            state1 = d1.init_state(st)
            d1.arrays_to_state([d1.mod.variable(t, dtype=dt) for t in terms], state1)
            loss1, grads1, _, _, _ = problem1.eval_loss_grad(state1)
            assert abs(loss - float(loss1)) < tol * abs(float(loss1)), (loss, float(loss1))
            assert relerr(U_glob, d1.field(state1, key).full()) < tol
            for a, b in zip(g_glob, grads1):
                e = relerr(a, b)
                worst = max(worst, e if dt == np.float64 else 0.0)
                assert e < tol, (maker.__name__, cshape, nlvl, dt, e)
            # three Adam epochs both ways
            x = domain.arrays_from_state(state)
            x1 = d1.arrays_from_state(state1)
            m, v = [torch.zeros_like(a) for a in x], [torch.zeros_like(a) for a in x]
            m1, v1 = [torch.zeros_like(a) for a in x1], [torch.zeros_like(a) for a in x1]
            for t in range(1, 4):
                alpha, o1, o2 = adam_scalars(0.01, 0.9, 0.999, t, dt)
                _, g, _, _, _ = problem.eval_loss_grad(state)
                native.adam_step(x, m, v, g, alpha, o1, o2, 1e-7)
                _, g1, _, _, _ = problem1.eval_loss_grad(state1)
                native.adam_step(x1, m1, v1, g1, alpha, o1, o2, 1e-7)
            la, lb = float(problem.eval_loss_grad(state)[0]), float(problem1.eval_loss_grad(state1)[0])
            assert abs(la - lb) < 10 * tol * abs(lb), (la, lb)
    dist.barrier()
    if rank == 0:
        print(f"SLAB_WORKER_OK world={world} worst_f64_grad_relerr={worst:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
