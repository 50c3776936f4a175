"""
Driver utilities: shared command-line flags, the optimize() entry points and the per-epoch callback
that writes train.log / train.csv and reports throughput in Mcells/s (reference src/odil/util.py).
The throughput definition is the headline metric of this repository:
    Mcells/s = prod(domain.cshape) / (wall time per epoch, callback time excluded) / 1e6
(util.py:383-386, :408-419).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import psutil

from .history import History
from .optimizer import Optimizer, make_optimizer

g_log_file = sys.stderr
g_log_echo = False


def assert_equal(first, second, msg=""):
    if not (first == second):
        raise ValueError("Expected equal '{:}' and '{:}'{}".format(first, second, msg))


def set_log_file(f=None, echo=None):
    global g_log_file, g_log_echo
    if f is not None:
        g_log_file = f
    if echo is not None:
        g_log_echo = echo


def printlog(*msg):
    line = " ".join(map(str, msg)) + "\n"
    if g_log_echo and g_log_file != sys.stderr:
        sys.stderr.write(line)
        sys.stderr.flush()
    g_log_file.write(line)
    g_log_file.flush()


class Timer:
    """Stack of named wall-clock timers."""

    def __init__(self):
        self._starts = []
        self.counters = dict()

    def push(self, key=None):
        self._starts.append((key, time.time()))

    def pop(self, key=None):
        k0, t0 = self._starts.pop()
        assert k0 is None or key is None or k0 == key, \
            "Inconsistent keys passed to push() and pop(): {:} and {:}".format(k0, key)
        key = k0 if key is None else key
        self.counters[key] = self.counters.get(key, 0.0) + time.time() - t0

    def append(self, timer):
        for k, v in timer.counters.items():
            self.counters[k] = self.counters.get(k, 0.0) + v


def get_error(u, v):
    d = np.asarray(u) - np.asarray(v)
    return np.mean(abs(d)), np.mean(d ** 2) ** 0.5, np.max(abs(d))


def add_arguments(parser):
    """Flags shared by all problem scripts (util.py:70-149); names, defaults and help kept."""
    a = parser.add_argument
    a("--epochs", type=int, default=None, help="Maximum epochs, defaults to product of plot_every and frames")
    a("--every_factor", type=float, default=1, help="Multiplier for all *_every options")
    a("--plot_every", type=int, default=5, help="Epochs between plots")
    a("--report_every", type=int, default=10, help="Epochs between reports to stdout")
    a("--history_every", type=int, default=1, help="Epochs between entries of training history")
    a("--checkpoint_every", type=int, default=0, help="Epochs between checkpoints")
    a("--frames", type=int, default=10, help="Frames to plot. Zero disables first frame.")
    a("--outdir", type=str, default=".", help="Output directory")
    a("--optimizer", type=str, default="adamn", help="Optimizer")
    a("--seed", default=1000, type=int, help="Seed for numpy.random and the backend's random")
    a("--plot_title", type=int, default=0, help="Enable title in plots")
    a("--plotext", type=str, default="pdf", help="Extension of plots")
    a("--history_full", type=int, default=0, help="Number of epochs to write history at every point")
    a("--montage", type=int, default=1, help="Run montage after plotting")
    a("--double", type=int, default=None, help="Double precision. Defaults to runtime.dtype")
    a("--echo", type=int, default=0, help="Echo log to stderr")
    a("--epoch_start", type=int, default=0, help="Initial value of epoch")
    a("--frame_start", type=int, default=0, help="Initial value of frame")
    a("--checkpoint", type=str, help="Continue from checkpoint in state_*.pickle")
    a("--checkpoint_train", type=str,
      help="Continue from history in state_*_train.pickle. By default, infers the name from --checkpoint. "
           "Set to '' to disable default behavior")
    a("--callback_update_state", type=int, default=0, help="Update state after callback")
    a("--bfgs_m", type=int, default=50, help="History size for L-BFGS")
    a("--bfgs_maxls", type=int, default=50, help="Max evaluations in line search")
    a("--bfgs_pgtol", type=float, default=None, help="Convergence tolerance for L-BFGS-B")
    a("--adam_epsilon", type=float, help="Parameter epsilon in Adam")
    a("--adam_beta_1", type=float, help="Parameter beta_1 in Adam")
    a("--adam_beta_2", type=float, help="Parameter beta_2 in Adam")
    a("--multigrid", type=int, default=0, help="Use multigrid decomposition")
    a("--mg_interp", type=str, default="stack", choices=["conv", "stack"],
      help="Multigrid interpolation method (both map to the same CUDA kernel)")
    a("--dump_data", type=int, default=1, help="Dump data_*.pickle with every plot")
    a("--jac_nsmp0", type=int, default=50, help="Number of samples for initialization of Jacobi optimizer")
    a("--jac_nsmp1", type=int, default=1, help="Number of samples for each step of Jacobi optimizer")
    a("--jac_factor", type=float, default=1, help="Factor for the diagonal update of Jacobi optimizer")
    a("--jac_epsilon", type=float, default=1e-8, help="Parameter epsilon in Jacobi optimizer")
    a("--nn_initializer", type=str, default="legacy", choices=["legacy", "glorot", "lecun", "he"],
      help="Initializer for weights of neural networks")


def optimize_newton(args, problem, state, callback=None, **kwargs):
    domain = problem.domain

    def eval_pinfo(state):
        loss, _, terms, names, norms = problem.eval_loss_grad(state)
        return {"terms": terms, "names": names, "norms": norms, "loss": loss}

    from .linsolver import solve

    opt = Optimizer(name="newton", displayname="Newton")
    printlog("Running {} optimizer".format(opt.displayname))
    pinfo = eval_pinfo(state)
    if callback:
        callback(state, args.epoch_start, pinfo)
    for epoch in range(args.epoch_start, args.epochs):
        vector, matrix = problem.linearize(state)
        opt.evals += 1
        linstatus = dict()
        delta = solve(matrix, -vector, args, linstatus, args.linsolver)
        if args.linsolver_verbose:
            printlog(linstatus)
        packed = domain.pack_state(state)
        domain.unpack_state(packed + delta, state)
        if callback:
            pinfo = eval_pinfo(state)
            pinfo["linsolver"] = linstatus
            callback(state, epoch + 1, pinfo)
    return domain.arrays_from_state(state), argparse.Namespace(epochs=args.epochs, evals=args.epochs)


def optimize_grad(args, optname, problem, state, callback=None, **kwargs):
    """Gradient-based optimization of `state` (util.py:190-240)."""
    domain = problem.domain
    mod = domain.mod

    def loss_grad(arrays):
        domain.arrays_to_state(arrays, state)
        loss, grads, terms, names, norms = problem.eval_loss_grad(state)
        return loss, grads, {"terms": terms, "names": names, "norms": norms, "loss": loss}

    def callback_wrap(arrays, epoch, pinfo):
        domain.arrays_to_state(arrays, state)
        callback(state, epoch, pinfo)
        if args.callback_update_state:
            new = domain.arrays_from_state(state)
            for i in range(len(new)):
                arrays[i] = new[i]

    for flag, name in [("bfgs_m", "m"), ("bfgs_pgtol", "pgtol"), ("bfgs_maxls", "maxls"), ("adam_epsilon", "epsilon"),
                       ("adam_beta_1", "beta_1"), ("adam_beta_2", "beta_2")]:
        if getattr(args, flag, None) is not None:
            kwargs[name] = getattr(args, flag)

    opt = make_optimizer(optname, dtype=domain.dtype, mod=mod, **kwargs)
    printlog("Running {} optimizer".format(opt.displayname))
    # The row for epoch_start carries the loss of the initial state.
    arrays = domain.arrays_from_state(state)
    _, _, pinfo = loss_grad(arrays)
    if callback:
        callback(state, args.epoch_start, pinfo)
    arrays, optinfo = opt.run(arrays, loss_grad=loss_grad, epochs=args.epochs - args.epoch_start,
                              callback=callback_wrap if callback else None, epoch_start=args.epoch_start,
                              lr=args.lr, **kwargs)
    domain.arrays_to_state(arrays, state)
    return arrays, optinfo


def optimize(args, optname, problem, state, callback, **kwargs):
    if optname == "newton":
        return optimize_newton(args, problem, state, callback, **kwargs)
    return optimize_grad(args, optname, problem, state, callback, **kwargs)


def get_memory_usage_kb():
    return psutil.Process().memory_info().rss // 1024


def get_gpu_memory_usage_kb():
    """(bytes in use, bytes reserved by the allocator) on the current device, in KiB."""
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.memory_allocated() // 1024, torch.cuda.memory_reserved() // 1024
    except Exception:
        pass
    return 0, 0


def get_env_config():
    keys = ["OMP_NUM_THREADS", "CUDA_VISIBLE_DEVICES", "ODIL_WARN", "ODIL_BACKEND", "ODIL_JIT", "ODIL_MT", "ODIL_DTYPE"]
    return {k: os.environ.get(k, "") for k in keys}


def setup_outdir(args, relpath_args=None):
    """Creates the output directory with args.json and train.log, enters it, fixes *_every and seeds."""
    from . import runtime

    outdir = args.outdir
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "args.json"), "w") as f:
        d = dict(vars(args), **get_env_config(), runtime_backend=runtime.backend_name,
                 runtime_dtype=runtime.dtype_name, runtime_jit=runtime.enable_jit, runtime_gpu=runtime.enable_gpu)
        json.dump(d, f, sort_keys=True, indent=4)
    os.chdir(outdir)
    set_log_file(open("train.log", "w"), echo=args.echo)
    for k in relpath_args or []:
        if getattr(args, k):
            setattr(args, k, os.path.relpath(getattr(args, k), start=outdir))

    def scaled(v):
        return None if v is None else max(1, round(v * args.every_factor))

    args.plot_every = scaled(args.plot_every)
    args.history_every = scaled(args.history_every)
    args.report_every = scaled(args.report_every)
    if args.epochs is None:
        args.epochs = args.frames * args.plot_every
    if args.seed is not None:
        np.random.seed(args.seed)
        runtime.mod.random.set_seed(args.seed)
    printlog(" ".join(sys.argv))


def make_callback(problem, args=None, epoch_func=None, report_func=None, history_func=None, checkpoint_func=None,
                  plot_func=None):
    """
    Returns callback(state, epoch, pinfo) doing report / history / plot / checkpoint by modulo
    (util.py:337-466).  Time spent inside the callback is excluded from the walltime used for the
    throughput line.
    """
    cb = argparse.Namespace(walltime=0, epoch=0, time_callback=0, time_start=time.time(), problem=problem, args=args,
                            frame=0)
    cb.history = History(csvpath="train.csv", warmup=1) if args.history_every else None

    def device_sync():
        try:
            import torch

            if torch.cuda.is_available():
                torch.cuda.synchronize()
        except Exception:
            pass

    def callback(state, epoch, pinfo):
        problem, args, history = cb.problem, cb.args, cb.history
        domain = problem.domain
        cb.task_report = args.report_every and epoch % args.report_every == 0
        cb.task_history = history is not None and (epoch % args.history_every == 0 or epoch < args.history_full)
        cb.task_plot = epoch % args.plot_every == 0 and (epoch or args.frames)
        cb.task_checkpoint = args.checkpoint_every and epoch % args.checkpoint_every == 0
        if cb.task_report or cb.task_history or cb.task_plot or cb.task_checkpoint:
            device_sync()  # the device runs ahead of the host; account its time before the callback's
        t_prev = time.time()
        cb.pinfo = pinfo
        if isinstance(problem.tracers, dict):
            problem.tracers["epoch"] = epoch
        if epoch_func is not None:
            epoch_func(problem, state, epoch, cb)
        now = time.time()
        cb.time_callback += now - t_prev
        t_prev = now
        walltime = now - cb.time_start - cb.time_callback

        if cb.task_report:
            printlog("\nepoch={:05d}".format(epoch))
            if pinfo and "norms" in pinfo:
                printlog("residual: " + ", ".join("{}:{:.5g}".format(name or str(i), float(np.array(norm)))
                                                  for i, (norm, name) in enumerate(zip(pinfo["norms"], pinfo["names"]))))
            if report_func is not None:
                report_func(problem, state, epoch, cb)
            gpu_used, gpu_pool = get_gpu_memory_usage_kb()
            printlog("memory: {:} MiB, gpu_used: {:} MiB, gpu_pool: {:} MiB".format(
                get_memory_usage_kb() // 1024, gpu_used // 1024, gpu_pool // 1024))
            if epoch > cb.epoch:
                wte = (walltime - cb.walltime) / (epoch - cb.epoch)
                thr = np.prod(domain.cshape) / wte if wte > 0 else 0
            else:
                wte = thr = 0
            printlog("walltime: {:.3f} s".format(walltime)
                     + ", walltime+callback: {:.3f} s".format(walltime + cb.time_callback)
                     + ", walltime/epoch: {:.3f} ms".format(wte * 1000))
            printlog("throughput: {:.3f} Mcells/s".format(thr / 1e6))
            cb.walltime = walltime
            cb.epoch = epoch
            cb.throughput = thr / 1e6

        if cb.task_history:
            gpu_used, gpu_pool = get_gpu_memory_usage_kb()
            history.append("epoch", epoch)
            history.append("frame", cb.frame)
            if pinfo and "norms" in pinfo:
                for i, (norm, name) in enumerate(zip(pinfo["norms"], pinfo["names"])):
                    history.append("norm_{:}".format(name or str(i)), np.array(norm))
            if pinfo and "loss" in pinfo:
                history.append("loss", np.array(pinfo["loss"]))
            if getattr(args, "linsolver_history", 0) and "linsolver" in pinfo:
                for key, val in pinfo["linsolver"].items():
                    if isinstance(val, (int, float, str, np.floating)):
                        history.append("lin_" + key, val)
            history.append("walltime", np.round(walltime, 3))
            history.append("memory", get_memory_usage_kb() // 1024)
            history.append("gpu_used", gpu_used // 1024)
            history.append("gpu_pool", gpu_pool // 1024)
            if history_func is not None:
                history_func(problem, state, epoch, history, cb)
            history.write()

        if cb.task_plot:
            if plot_func is not None:
                plot_func(problem, state, epoch, cb.frame, cb)
            cb.frame += 1

        if cb.task_checkpoint:
            if checkpoint_func is not None:
                checkpoint_func(problem, state, epoch, cb)
            else:
                from .core import checkpoint_save

                path = "checkpoint_{:06d}.pickle".format(epoch)
                printlog(path)
                checkpoint_save(domain, state, path)

        cb.time_callback += time.time() - t_prev

    callback.cbinfo = cb
    return callback
