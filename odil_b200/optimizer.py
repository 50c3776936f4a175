"""
Optimizers behind the reference's optimizer seam (src/odil/optimizer.py): `make_optimizer(name)` returns
an object with `.run(x0, loss_grad, epochs, callback, epoch_start, lr, **kw) -> (arrays, optinfo)`.

  adam / adamn   device-resident Adam: one multi-tensor CUDA launch per epoch
                 (odil_b200_adam_step; reference AdamNativeOptimizer, optimizer.py:280-341)
  gd             device-resident gradient descent (odil_b200_gd_step; optimizer.py:256-277)
  lbfgsb / lbfgs device-resident L-BFGS (odil_b200/lbfgs.py: compact-form two-pass update + More-Thuente line
                 search following L-BFGS-B 3.0's control flow); `lbfgsb_scipy` (or ODIL_LBFGS=scipy) is the
                 reference's own arrangement, SciPy fmin_l_bfgs_b on a host fp64 vector (optimizer.py:29-117).
  adam_tf, tfp lbfgs: TensorFlow-only wrappers of the reference, not provided.
"""
import os
from argparse import Namespace

import numpy as np
import torch

from . import native


LAST_RUN_INFO = {"graph": False}  # how the most recent optimizer run executed its epochs (read by bench.py)


class Optimizer:

    def __init__(self, name=None, displayname=None, dtype=None):
        self.name = name
        self.displayname = displayname if displayname is not None else name
        self.dtype = dtype
        self.pinfo = None
        self.evals = 0

    def run(self, x0, loss_grad, epochs, callback=None, epoch_start=0, **kwargs):
        return x0, Namespace(evals=0, epochs=0)


class EarlyStopError(Exception):

    def __init__(self, msg, optinfo):
        super().__init__(msg)
        self.optinfo = optinfo


def adam_scalars(lr, beta_1, beta_2, local_epoch, dtype):
    """Bias-corrected step size and (1 - beta) factors, evaluated IN `dtype` like optimizer.py:307-314."""
    d = np.dtype(dtype).type
    lr, b1, b2, t = d(lr), d(beta_1), d(beta_2), d(local_epoch)
    alpha = lr * np.sqrt(d(1) - b2 ** t) / (d(1) - b1 ** t)
    return float(d(alpha)), float(d(d(1) - b1)), float(d(d(1) - b2))


def _device_copy(arrays):
    out = []
    for a in arrays:
        if not (torch.is_tensor(a) and a.is_cuda):
            raise native.NativeError("optimizer state must be CUDA tensors (use domain.init_state)")
        out.append(a.clone().contiguous())
    return out


class AdamNativeOptimizer(Optimizer):

    def __init__(self, dtype=None, mod=None, **kwargs):
        super().__init__(name="adamn", displayname="AdamNative", dtype=dtype)
        self.mod = mod

    def run(self, x0, loss_grad, epochs=None, callback=None, lr=1e-3, epoch_start=0, beta_1=0.9, beta_2=0.999,
            epsilon=1e-7, jit=True, graph=None, **kwargs):
        """
        graph: replay each epoch (loss_grad + Adam update) as ONE CUDA graph after two eager epochs.  The
        reference's `jit` compiles the epoch into one XLA program (optimizer.py:321-326); this is the B200
        counterpart for launch-bound (small-grid) problems: ~10 kernel launches and their ctypes calls become
        one cudaGraphLaunch.  Default: env ODIL_B200_GRAPH (0).  Same kernels, same arithmetic, same results.
        """
        dtype = np.dtype(self.dtype if self.dtype is not None else np.float32)
        x = _device_copy(x0)
        m = [torch.zeros_like(e) for e in x]
        v = [torch.zeros_like(e) for e in x]
        eps = float(dtype.type(epsilon))
        if graph is None:
            env = os.environ.get("ODIL_B200_GRAPH", "")
            if env != "":
                graph = env != "0"
            else:
                # default: replay where the epoch is launch-bound (small grids) and the caller vouches that nothing
                # outside the captured kernels changes between epochs (`loss_grad.graph_safe`, set by optimize_grad:
                # affine operator, no tracer baked into its tables, state arrays not swapped by the callback)
                graph = bool(getattr(loss_grad, "graph_safe", False)) and sum(e.numel() for e in x) <= (1 << 25)
        if graph and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1 and os.environ.get("ODIL_B200_COMM", "peer") == "nccl":
            graph = False  # NCCL groups issued through torch.distributed are not captured; the peer-memory
            #                communicator (csrc/comm.cu) is plain kernels and replays
        first, last = epoch_start + 1, epoch_start + epochs
        eager_until = last if not graph else min(last, first + 1)
        LAST_RUN_INFO["graph"] = bool(graph and eager_until < last)
        # Fusion across the optimizer seam: an engine that can apply the update of the finest multigrid term while it
        # streams that term's gradient (odil_b200_mg_interp_adjoint_adam) is told the step in advance and hands back
        # None in place of the gradients it has consumed.
        fuse = getattr(loss_grad, "fuse_adam", None)
        # The other fusion across the seam (the default where the engine offers it): the engine applies the whole
        # update and, while it rewrites the finest multigrid term, synthesises the regular field of the next
        # evaluation (odil_b200_adam_synth) -- same arithmetic, the finest term is read once per epoch instead of twice.
        synth = getattr(loss_grad, "adam_synth", None) if fuse is None else None
        for epoch in range(first, eager_until + 1):
            self.evals += 1
            alpha, omb1, omb2 = adam_scalars(lr, beta_1, beta_2, epoch - epoch_start, dtype)
            if fuse is not None:
                fuse(x, m, v, alpha, omb1, omb2, eps)
            loss, grads, pinfo = loss_grad(x)
            if synth is not None:
                synth(x, m, v, grads, alpha, omb1, omb2, eps)
            else:
                rest = [i for i, g in enumerate(grads) if g is not None]
                native.adam_step([x[i] for i in rest], [m[i] for i in rest], [v[i] for i in rest],
                                 [grads[i] for i in rest], alpha, omb1, omb2, eps)
            if epoch > 0 and callback is not None:
                callback(x, epoch, pinfo)
        if eager_until < last:
            self._run_graph(x, m, v, loss_grad, callback, lr, beta_1, beta_2, eps, dtype, epoch_start,
                            eager_until + 1, last)
        return x, Namespace(epochs=epochs, evals=self.evals)

    def _run_graph(self, x, m, v, loss_grad, callback, lr, beta_1, beta_2, eps, dtype, epoch_start, first, last):
        """Epochs first..last as replays of one captured graph.  Only the step size changes between epochs: the
        host tabulates it for every epoch (in `dtype`, like the eager path) and the head of the graph picks
        entry `step` of the device copy and advances `step`, so a replay reads nothing from the host and the
        host may run ahead of the device freely."""
        dev = x[0].device
        table = torch.tensor([adam_scalars(lr, beta_1, beta_2, e - epoch_start, dtype)[0]
                              for e in range(first, last + 1)], dtype=torch.float64).to(dev)
        step = torch.zeros(1, dtype=torch.int64, device=dev)
        alpha_dev = torch.zeros(1, dtype=torch.float64, device=dev)
        _, omb1, omb2 = adam_scalars(lr, beta_1, beta_2, 1, dtype)
        held = list(x)  # the tensors whose addresses the graph holds
        g = torch.cuda.CUDAGraph()
        # A destructor that frees device memory (a dropped stencil plan, a torch tensor of an earlier problem held in a
        # reference cycle) must not run inside the capture: collect what is collectable now; native handles whose
        # destructor still runs during the capture are parked (native._release) and freed after it.
        import gc

        gc.collect()
        n_before = native.launch_count()
        with native.capture_guard(), torch.cuda.graph(g):
            native.table_pick(table, step, alpha_dev)  # alpha_dev = table[step]; step += 1 (one launch)
            fuse = getattr(loss_grad, "fuse_adam", None)
            if fuse is not None:
                fuse(held, m, v, 0.0, omb1, omb2, eps, alpha_dev=alpha_dev)
            loss, grads, pinfo = loss_grad(held)
            synth = getattr(loss_grad, "adam_synth", None) if fuse is None else None
            if synth is not None:
                synth(held, m, v, grads, 0.0, omb1, omb2, eps, alpha_dev=alpha_dev)
            else:
                rest = [i for i, gr in enumerate(grads) if gr is not None]
                native.adam_step_dev([held[i] for i in rest], [m[i] for i in rest], [v[i] for i in rest],
                                     [grads[i] for i in rest], alpha_dev, omb1, omb2, eps)
        native.flush_deferred()
        nodes = native.launch_count() - n_before  # library kernels captured into one replay
        fetch = getattr(loss, "_fetch", None)
        self._graph = g  # owns the memory pool of grads / sums that pinfo still points into after run()
        for epoch in range(first, last + 1):
            self.evals += 1
            for i in range(len(held)):
                if x[i] is not held[i]:  # a callback replaced the array (callback_update_state)
                    held[i].copy_(x[i])
                    x[i] = held[i]
            g.replay()
            native.note_replayed_launches(nodes)
            if fetch is not None:
                fetch.rearm()
            if epoch > 0 and callback is not None:
                callback(x, epoch, pinfo)


class GdOptimizer(Optimizer):

    def __init__(self, dtype=None, mod=None, **kwargs):
        super().__init__(name="gd", displayname="GD", dtype=dtype)
        self.mod = mod

    def run(self, x0, loss_grad, epochs=None, callback=None, lr=1e-3, epoch_start=0, **kwargs):
        x = _device_copy(x0)
        for epoch in range(epoch_start + 1, epoch_start + epochs + 1):
            self.evals += 1
            loss, grads, pinfo = loss_grad(x)
            native.gd_step(x, grads, lr)
            if epoch > 0 and callback is not None:
                callback(x, epoch, pinfo)
        return x, Namespace(epochs=epochs, evals=self.evals)


def _reject_slabs(name):
    """L-BFGS makes its line-search and stopping decisions from rank-local dot products; on slab-decomposed
    grids the ranks would disagree on the number of evaluations and dead-lock the halo exchange."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 \
            and int(os.environ.get("ODIL_SLABS", "1")) != 0:
        raise NotImplementedError(
            f"optimizer '{name}' is not available on slab-decomposed (multi-rank) grids; use 'adam' or 'gd'")


class LbfgsbOptimizer(Optimizer):

    def __init__(self, pgtol=1e-16, m=50, maxls=50, factr=0, dtype=None, mod=None, **kwargs):
        """
        pgtol: projected-gradient tolerance; m: number of correction pairs; maxls: line-search steps per
        iteration; factr: relative-reduction stop factor (0 disables) -- as scipy.optimize.fmin_l_bfgs_b.
        """
        super().__init__(name="lbfgsb", displayname="L-BFGS-B", dtype=dtype)
        self.mod = mod
        self.pgtol, self.m, self.maxls, self.factr = pgtol, m, maxls, factr

    def run(self, x0, loss_grad, epochs=None, callback=None, epoch_start=0, **kwargs):
        from scipy import optimize

        _reject_slabs(self.name)
        self.epoch = epoch_start
        tdtype = x0[0].dtype
        device = x0[0].device
        shapes = [tuple(a.shape) for a in x0]
        sizes = [int(np.prod(s)) for s in shapes]
        bounds = np.cumsum([0] + sizes)

        def to_arrays(flat):
            t = torch.from_numpy(np.ascontiguousarray(flat)).to(device=device, dtype=tdtype)
            return [t[bounds[i]:bounds[i + 1]].reshape(shapes[i]).contiguous() for i in range(len(sizes))]

        def to_flat(arrays):
            return torch.cat([a.reshape(-1) for a in arrays]).to(torch.float64).cpu().numpy()

        def func(flat):
            self.evals += 1
            loss, grads, pinfo = loss_grad(to_arrays(flat))
            self.pinfo = pinfo
            return float(loss), to_flat(grads)

        def on_iteration(flat):
            self.epoch += 1
            if callback:
                callback(to_arrays(flat), self.epoch, self.pinfo)

        x, f, info = optimize.fmin_l_bfgs_b(func=func, x0=to_flat(x0), maxiter=epochs, pgtol=self.pgtol, m=self.m,
                                            maxls=self.maxls, factr=self.factr, maxfun=np.inf,
                                            callback=on_iteration)
        optinfo = Namespace(warnflag=info["warnflag"], task=info["task"], evals=info["funcalls"], epochs=info["nit"])
        if optinfo.warnflag not in [0, 1] or optinfo.epochs < epochs:
            raise EarlyStopError(", ".join("{:}={:}".format(k, info.get(k, ""))
                                           for k in ["warnflag", "task", "funcalls", "nit"]), optinfo)
        return to_arrays(x), optinfo


class LbfgsDeviceOptimizer(Optimizer):
    """
    `lbfgsb` with the vector algebra in HBM (odil_b200/lbfgs.py): same arguments and stopping rules as
    LbfgsbOptimizer / scipy.optimize.fmin_l_bfgs_b for unconstrained problems, but the fp64 unknown
    vector and the 2m history vectors never leave the device.
    """

    def __init__(self, pgtol=1e-16, m=50, maxls=50, factr=0, dtype=None, mod=None, **kwargs):
        super().__init__(name="lbfgsb", displayname="L-BFGS-B (device)", dtype=dtype)
        self.mod = mod
        self.pgtol, self.m, self.maxls, self.factr = pgtol, m, maxls, factr

    def run(self, x0, loss_grad, epochs=None, callback=None, epoch_start=0, **kwargs):
        from . import lbfgs

        _reject_slabs(self.name)
        self.epoch = epoch_start
        tdtype, device = x0[0].dtype, x0[0].device
        shapes = [tuple(a.shape) for a in x0]
        sizes = [int(np.prod(s)) for s in shapes]
        bounds = np.cumsum([0] + sizes)

        # The optimizer's vector is fp64 (SciPy's L-BFGS-B is fp64 end to end, optimizer.py:63-73,81-88); the operator
        # works in the problem dtype.  One conversion pass each way per evaluation, into buffers allocated once: the
        # unknowns into `work` (what loss_grad reads), the gradients into the flat fp64 vector `g64` -- no torch.cat,
        # no intermediate copies.
        work = [torch.empty(shapes[i], dtype=tdtype, device=device) for i in range(len(sizes))]
        g64 = torch.empty(int(bounds[-1]), dtype=torch.float64, device=device)

        def to_arrays(flat, fresh=False):
            out = [torch.empty_like(w) for w in work] if fresh else work
            for i, w in enumerate(out):
                w.view(-1).copy_(flat[bounds[i]:bounds[i + 1]])
            return out

        def func(flat):
            self.evals += 1
            loss, grads, pinfo = loss_grad(to_arrays(flat))
            self.pinfo = pinfo
            for i, a in enumerate(grads):
                g64[bounds[i]:bounds[i + 1]].copy_(a.reshape(-1))
            return float(loss), g64

        def on_iteration(flat):
            self.epoch += 1
            if callback:
                callback(to_arrays(flat, fresh=True), self.epoch, self.pinfo)  # the callback may keep them

        flat0 = torch.cat([a.reshape(-1).to(torch.float64) for a in x0]).to(device)
        x, f, info = lbfgs.minimize(func, flat0, m=self.m, maxiter=epochs, maxls=self.maxls, pgtol=self.pgtol,
                                    factr=self.factr, callback=on_iteration)
        optinfo = Namespace(warnflag=info["warnflag"], task=info["task"], evals=info["funcalls"], epochs=info["nit"])
        if optinfo.warnflag not in [0, 1] or optinfo.epochs < epochs:
            raise EarlyStopError(", ".join("{:}={:}".format(k, info.get(k, ""))
                                           for k in ["warnflag", "task", "funcalls", "nit"]), optinfo)
        return to_arrays(x, fresh=True), optinfo


def make_optimizer(name, dtype=None, mod=None, **kwargs):
    if name in ("lbfgsb", "lbfgs"):
        import os

        if os.environ.get("ODIL_LBFGS", "device") == "scipy":
            return LbfgsbOptimizer(dtype=dtype, mod=mod, **kwargs)
        return LbfgsDeviceOptimizer(dtype=dtype, mod=mod, **kwargs)
    if name == "lbfgsb_scipy":
        return LbfgsbOptimizer(dtype=dtype, mod=mod, **kwargs)
    if name in ("adam", "adamn"):
        return AdamNativeOptimizer(dtype=dtype, mod=mod, **kwargs)
    if name == "gd":
        return GdOptimizer(dtype=dtype, mod=mod, **kwargs)
    if name == "adam_tf":
        raise ValueError("Optimizer 'adam_tf' wraps Keras and is not provided by the B200 backend; use 'adam'")
    raise ValueError("Unknown optimizer '{}'".format(name))
