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

    # -- the communicator itself (csrc/comm.cu) against torch.distributed: ragged sizes, repeated calls ------------
    if torch.cuda.is_available() and os.environ.get("ODIL_B200_COMM", "peer") != "nccl":
        from odil_b200.slab import SlabInfo

        slab = SlabInfo(rank, world, halo=2)
        ref = SlabInfo(rank, world, halo=2)
        ref.use_peer = False
        gen = torch.Generator(device="cuda").manual_seed(100 + rank)
        for it in range(6):
            shapes = [(8 + 4, 5, 7), (4 + 4, 3), (16 + 4, 64, 32), (6 + 4, 1, 9)][: 2 + it % 3]
            dts = [torch.float32, torch.float64, torch.float32, torch.float64]
            a = [torch.randn(s, dtype=d, device="cuda", generator=gen) for s, d in zip(shapes, dts)]
            b = [t.clone() for t in a]
            width = 1 + it % 2
            slab.exchange(a, width=width)
            ref.exchange(b, width=width)
            for x, y in zip(a, b):
                assert torch.equal(x, y), ("halo exchange differs from the NCCL exchange", it, x.shape)
            # accumulate (transpose of the exchange): peer-memory kernels == isend / irecv + add
            one = dts[it % 2]  # one element type per call (odil_b200_halo_accumulate takes a single dtype)
            pa = [torch.randn(s, dtype=one, device="cuda", generator=gen) for s in shapes]
            pb = [t.clone() for t in pa]
            items = lambda ts: [(t[0:1], t[-1:], t[1:2], t[-2:-1]) for t in ts]
            slab.accumulate(items(pa))
            ref.accumulate(items(pb))
            for x, y in zip(pa, pb):
                assert torch.equal(x, y), ("halo accumulation differs from the isend/irecv arrangement", it, x.shape)
            s1 = torch.randn(3 + it, dtype=torch.float64, device="cuda", generator=gen)
            s2 = s1.clone()
            slab.all_reduce_sum(s1)
            dist.all_reduce(s2)
            assert torch.allclose(s1, s2, rtol=1e-14, atol=0), (s1, s2)
            gathered = [torch.empty_like(s1) for _ in range(world)]
            dist.all_gather(gathered, s1)
            assert all(torch.equal(g, s1) for g in gathered), "all-reduce result differs between ranks"
        if rank == 0:
            print("COMM_OK peer-memory exchange == NCCL exchange; all-reduce identical on all ranks")

    worst = 0.0
    for maker, cshape, nlvl in [(ops.make_poisson, (32 * world, 16, 24), 3), (ops.make_poisson, (16 * world, 12, 8), 0),
                                (ops.make_poisson, (16 * world, 16), 2), (ops.make_wave, (16 * world, 12), 0),
                                (ops.make_wave, (16 * world, 8), 2)]:
        for dt in (np.float64, np.float32):
            tol = 1e-11 if dt == np.float64 else 5e-6
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
    # -- the Adam epoch through the public API: CUDA-graph replay of the slab epoch == eager epochs ---------------------
    if torch.cuda.is_available() and os.environ.get("ODIL_B200_COMM", "peer") != "nccl":
        import argparse

        # ... and the fusion of the finest term's Adam update with the synthesis of the next regular field
        # (odil_b200_adam_synth on the owned planes + one halo exchange of U) == the unfused epoch, on the owned planes
        finals = {}
        for synth in ("0", "1"):
            for flag in ("0", "1"):
                os.environ["ODIL_B200_FUSE_SYNTH"] = synth
                os.environ["ODIL_B200_GRAPH"] = flag
                os.environ["ODIL_HALO"] = "2"
                problem, state = ops.make_poisson((32 * world, 16, 24), 3, np.float32)
                args = argparse.Namespace(epochs=12, epoch_start=0, lr=0.005, callback_update_state=0, bfgs_m=None,
                                          bfgs_pgtol=None, bfgs_maxls=None, adam_epsilon=None, adam_beta_1=None,
                                          adam_beta_2=None)
                losses = []
                n0 = native.ADAM_SYNTH_APPLIED
                odil.util.optimize_grad(args, "adam", problem, state,
                                        lambda st, ep, pinfo: losses.append(float(pinfo["loss"])))
                sl = problem.domain.slab
                finals[synth, flag] = (losses, [sl.owned(a).clone() for a in problem.domain.arrays_from_state(state)],
                                       native.ADAM_SYNTH_APPLIED - n0)
        os.environ.pop("ODIL_B200_GRAPH")
        os.environ.pop("ODIL_B200_FUSE_SYNTH")
        ref = finals["0", "0"]
        assert ref[2] == 0 and finals["1", "0"][2] > 0 and finals["1", "1"][2] > 0
        for key, (losses, arrs, _) in finals.items():
            assert losses == ref[0], (key, losses, ref[0])
            for a, b in zip(arrs, ref[1]):
                assert torch.equal(a, b), key
        if rank == 0:
            print("GRAPH_OK slab epoch replayed as a CUDA graph is bit-identical to eager epochs")
            print("SYNTH_OK slab epoch with odil_b200_adam_synth is bit-identical to the unfused epoch")
    dist.barrier()
    if rank == 0:
        print(f"SLAB_WORKER_OK world={world} worst_f64_grad_relerr={worst:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
