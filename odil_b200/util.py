"""
Driver layer around the engine: the command-line flags every ODIL problem script shares, `optimize()` and its two
back ends (gradient optimizers, Newton), output-directory set-up, and the per-epoch callback that writes
`train.log` / `train.csv`, plots, checkpoints and reports throughput.  Public names, flag names and defaults, log
lines and file formats follow the reference (src/odil/util.py: flags :70-149, `optimize_newton` :152-187,
`optimize_grad` :190-240, `setup_outdir` :281-334, `make_callback` :337-466) so that scripts and the tools that
read their output keep working.

The throughput line is the headline metric of this repository:
    Mcells/s = prod(domain.cshape) / (wall time per epoch, callback time excluded) / 1e6      (util.py:408-419)
The device runs ahead of the host here (nothing in an epoch synchronises), so the callback drains the device before
it reads the clock whenever it is about to report, record, plot or checkpoint.
"""
import argparse
import contextlib
import json
import os
import sys
import time

import numpy as np
import psutil

from .history import History
from .optimizer import Optimizer, make_optimizer


# --------------------------------------------------------------------------------------------------
# Logging and small helpers
# --------------------------------------------------------------------------------------------------
class _Log:
    stream = sys.stderr
    echo = False


def set_log_file(f=None, echo=None):
    """Redirects `printlog` to the open file `f`; `echo` also copies every line to stderr."""
    if f is not None:
        _Log.stream = f
    if echo is not None:
        _Log.echo = echo


def printlog(*msg):
    text = " ".join(str(m) for m in msg) + "\n"
    targets = [_Log.stream]
    if _Log.echo and _Log.stream is not sys.stderr:
        targets.insert(0, sys.stderr)
    for t in targets:
        t.write(text)
        t.flush()


def assert_equal(first, second, msg=""):
    if first != second:
        raise ValueError("Expected equal '{:}' and '{:}'{}".format(first, second, msg))


def get_error(u, v):
    """(mean absolute, root mean square, maximum) difference of two arrays."""
    d = np.abs(np.asarray(u) - np.asarray(v))
    return np.mean(d), np.sqrt(np.mean(d * d)), np.max(d)


class Timer:
    """Named wall-clock totals.  `push(key)` / `pop(key)` bracket a section (sections nest); `section(key)` does
    the same as a context manager; `append(other)` adds another timer's totals."""

    def __init__(self):
        self.counters = {}
        self._open = []

    def push(self, key=None):
        self._open.append((key, time.time()))

    def pop(self, key=None):
        opened, t0 = self._open.pop()
        if opened is not None and key is not None and opened != key:
            raise AssertionError("Inconsistent keys passed to push() and pop(): {:} and {:}".format(opened, key))
        name = opened if key is None else key
        self.counters[name] = self.counters.get(name, 0.0) + (time.time() - t0)

    @contextlib.contextmanager
    def section(self, key):
        self.push(key)
        try:
            yield self
        finally:
            self.pop(key)

    def append(self, timer):
        for name, total in timer.counters.items():
            self.counters[name] = self.counters.get(name, 0.0) + total


# --------------------------------------------------------------------------------------------------
# Command line
# --------------------------------------------------------------------------------------------------
# (flag, type, default, help); a default of ... stands for "none given"
_FLAGS = [
    ("epochs", int, None, "number of epochs; default: frames * plot_every"),
    ("every_factor", float, 1, "scales plot_every, report_every and history_every"),
    ("plot_every", int, 5, "plot every this many epochs"),
    ("report_every", int, 10, "print a report every this many epochs"),
    ("history_every", int, 1, "add a row to train.csv every this many epochs"),
    ("checkpoint_every", int, 0, "write checkpoint_*.pickle every this many epochs (0: never)"),
    ("frames", int, 10, "number of plotted frames (0 also drops the frame of the initial state)"),
    ("outdir", str, ".", "directory for args.json, train.log, train.csv, plots and checkpoints"),
    ("optimizer", str, "adamn", "adam / adamn, gd, lbfgsb / lbfgs, newton"),
    ("seed", int, 1000, "random seed (numpy and the backend)"),
    ("plot_title", int, 0, "put a title on plots"),
    ("plotext", str, "pdf", "file extension of plots"),
    ("history_full", int, 0, "record every epoch in train.csv below this epoch"),
    ("montage", int, 1, "assemble plots with montage"),
    ("double", int, None, "1: float64, 0: float32; default: the runtime's dtype"),
    ("echo", int, 0, "copy the log to stderr"),
    ("epoch_start", int, 0, "epoch counter of the initial state"),
    ("frame_start", int, 0, "frame counter of the initial state"),
    ("checkpoint", str, ..., "start from this state_*.pickle / checkpoint_*.pickle"),
    ("checkpoint_train", str, ..., "history pickle to continue train.csv from; default: derived from --checkpoint, "
                                   "'' disables"),
    ("callback_update_state", int, 0, "let the callback modify the state the optimizer continues from"),
    ("bfgs_m", int, 50, "L-BFGS: number of correction pairs"),
    ("bfgs_maxls", int, 50, "L-BFGS: line-search evaluations per iteration"),
    ("bfgs_pgtol", float, None, "L-BFGS-B: projected-gradient tolerance"),
    ("adam_epsilon", float, ..., "Adam: epsilon"),
    ("adam_beta_1", float, ..., "Adam: beta_1"),
    ("adam_beta_2", float, ..., "Adam: beta_2"),
    ("multigrid", int, 0, "represent fields as multigrid hierarchies"),
    ("mg_interp", str, "stack", "multigrid interpolation formulation (one CUDA kernel serves both)"),
    ("dump_data", int, 1, "write data_*.pickle with every plot"),
    ("jac_nsmp0", int, 50, "Jacobi optimizer: samples at initialisation"),
    ("jac_nsmp1", int, 1, "Jacobi optimizer: samples per step"),
    ("jac_factor", float, 1, "Jacobi optimizer: factor of the diagonal update"),
    ("jac_epsilon", float, 1e-8, "Jacobi optimizer: epsilon"),
    ("nn_initializer", str, "legacy", "initial weights of neural networks"),
]
_CHOICES = {"mg_interp": ["conv", "stack"], "nn_initializer": ["legacy", "glorot", "lecun", "he"]}


def add_arguments(parser):
    """Adds the flags shared by all problem scripts (same names, types and defaults as util.py:70-149)."""
    for name, typ, default, text in _FLAGS:
        kw = {"type": typ, "help": text}
        if default is not ...:
            kw["default"] = default
        if name in _CHOICES:
            kw["choices"] = _CHOICES[name]
        parser.add_argument("--" + name, **kw)


# --------------------------------------------------------------------------------------------------
# Optimization entry points
# --------------------------------------------------------------------------------------------------
def _progress_info(problem, state):
    loss, _, terms, names, norms = problem.eval_loss_grad(state)
    return {"terms": terms, "names": names, "norms": norms, "loss": loss}


def optimize_newton(args, problem, state, callback=None, **kwargs):
    """Newton iterations  state += solve(J, -F)  with the linear solver named by `args.linsolver`; the callback
    sees the initial state as epoch `epoch_start` and every iterate after it."""
    from .linsolver import solve

    domain = problem.domain
    opt = Optimizer(name="newton", displayname="Newton")
    printlog("Running {} optimizer".format(opt.displayname))
    if callback:
        callback(state, args.epoch_start, _progress_info(problem, state))
    for epoch in range(args.epoch_start + 1, args.epochs + 1):
        residual, jacobian = problem.linearize(state)
        opt.evals += 1
        status = {}
        step = solve(jacobian, -residual, args, status, args.linsolver)
        if args.linsolver_verbose:
            printlog(status)
        domain.unpack_state(domain.pack_state(state) + step, state)
        if callback:
            info = _progress_info(problem, state)
            info["linsolver"] = status
            callback(state, epoch, info)
    return domain.arrays_from_state(state), argparse.Namespace(epochs=args.epochs, evals=args.epochs)


_OPTIMIZER_FLAGS = {"bfgs_m": "m", "bfgs_pgtol": "pgtol", "bfgs_maxls": "maxls", "adam_epsilon": "epsilon",
                    "adam_beta_1": "beta_1", "adam_beta_2": "beta_2"}


def optimize_grad(args, optname, problem, state, callback=None, **kwargs):
    """Runs the gradient optimizer `optname` on `state` for epochs `epoch_start + 1 .. epochs`.  The callback first
    sees the initial state (its loss is the first row of train.csv), then every epoch."""
    domain = problem.domain

    def loss_grad(arrays):
        domain.arrays_to_state(arrays, state)
        loss, grads, terms, names, norms = problem.eval_loss_grad(state)
        return loss, grads, {"terms": terms, "names": names, "norms": norms, "loss": loss}

    def on_epoch(arrays, epoch, pinfo):
        domain.arrays_to_state(arrays, state)
        callback(state, epoch, pinfo)
        if args.callback_update_state:
            arrays[:] = domain.arrays_from_state(state)

    def graph_safe():
        """True if replaying the captured epoch is equivalent to running it: the operator lowered to static stencil
        plans (no tracer value inside its tables, no run-time parameters of generated kernels)."""
        engine = getattr(problem, "_cache_eval_loss_grad", {}).get("func")
        view = getattr(engine, "tracer_view", None)
        return (engine is not None and not hasattr(engine, "jacobian") and view is not None and not view.reads
                and not args.callback_update_state)

    for flag, name in _OPTIMIZER_FLAGS.items():
        if getattr(args, flag, None) is not None:
            kwargs[name] = getattr(args, flag)
    opt = make_optimizer(optname, dtype=domain.dtype, mod=domain.mod, **kwargs)
    printlog("Running {} optimizer".format(opt.displayname))
    arrays = domain.arrays_from_state(state)
    if callback:
        callback(state, args.epoch_start, loss_grad(arrays)[2])
    else:
        loss_grad(arrays)  # builds the engine before the optimizer's first timed epoch
    loss_grad.graph_safe = graph_safe()
    engine = getattr(problem, "_cache_eval_loss_grad", {}).get("func")
    # Opt-in (ODIL_B200_FUSE_ADAM=1): the fused kernel is bit-identical to the pair it replaces but, measured at 512^3
    # fp32 on B200, slower (0.874 ms vs 0.179 + 0.575 ms: at 168 registers it runs 12 warps per SM and is bound by
    # latency, not by the 4 bytes per cell it saves) -- profiles/README.md, round 2.
    if hasattr(engine, "request_fused_adam") and not hasattr(engine, "jacobian") \
            and os.environ.get("ODIL_B200_FUSE_ADAM", "0") not in ("", "0"):
        loss_grad.fuse_adam = engine.request_fused_adam
    # Default (ODIL_B200_FUSE_SYNTH=0 disables): the Adam update of the finest multigrid term also writes the regular
    # field of the next evaluation (engine.adam_synth_step, odil_b200_adam_synth), also on slabs.  Not with a callback that
    # rewrites the state between epochs, not for the generated-kernel engine.
    if hasattr(engine, "adam_synth_step") and not hasattr(engine, "jacobian") \
            and not args.callback_update_state and os.environ.get("ODIL_B200_FUSE_SYNTH", "1") not in ("", "0"):
        loss_grad.adam_synth = engine.adam_synth_step
    arrays, optinfo = opt.run(arrays, loss_grad=loss_grad, epochs=args.epochs - args.epoch_start,
                              callback=on_epoch if callback else None, epoch_start=args.epoch_start, lr=args.lr,
                              **kwargs)
    domain.arrays_to_state(arrays, state)
    return arrays, optinfo


def optimize(args, optname, problem, state, callback, **kwargs):
    """`optname` "newton" runs Newton iterations, anything else is a gradient optimizer of `make_optimizer`."""
    if optname == "newton":
        return optimize_newton(args, problem, state, callback, **kwargs)
    return optimize_grad(args, optname, problem, state, callback, **kwargs)


# --------------------------------------------------------------------------------------------------
# Process / device bookkeeping
# --------------------------------------------------------------------------------------------------
def get_memory_usage_kb():
    """Resident set size of this process in KiB."""
    return psutil.Process().memory_info().rss // 1024


def _cuda():
    try:
        import torch

        return torch.cuda if torch.cuda.is_available() else None
    except Exception:
        return None


def get_gpu_memory_usage_kb():
    """(allocated, reserved by the caching allocator) on the current device, in KiB; zeros without a GPU."""
    cuda = _cuda()
    return (cuda.memory_allocated() // 1024, cuda.memory_reserved() // 1024) if cuda else (0, 0)


_ENV_KEYS = ("OMP_NUM_THREADS", "CUDA_VISIBLE_DEVICES", "ODIL_WARN", "ODIL_BACKEND", "ODIL_JIT", "ODIL_MT",
             "ODIL_DTYPE")


def _is_writer():
    """False on ranks > 0 of a multi-process (slab) run."""
    try:
        import torch.distributed as dist

        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    except Exception:
        return True


def get_env_config():
    return {k: os.environ.get(k, "") for k in _ENV_KEYS}


def setup_outdir(args, relpath_args=None):
    """
    Prepares a run: creates `args.outdir`, records the configuration in `args.json` (flags, relevant environment,
    runtime settings), makes it the working directory, opens `train.log`, rewrites the path-valued flags named in
    `relpath_args` relative to it, applies `every_factor`, derives `epochs` when it was left open, and seeds NumPy
    and the backend.
    """
    from . import runtime

    os.makedirs(args.outdir, exist_ok=True)
    config = dict(vars(args))
    config.update(get_env_config())
    config.update(runtime_backend=runtime.backend_name, runtime_dtype=runtime.dtype_name,
                  runtime_jit=runtime.enable_jit, runtime_gpu=runtime.enable_gpu)
    if _is_writer():
        with open(os.path.join(args.outdir, "args.json"), "w") as f:
            json.dump(config, f, sort_keys=True, indent=4)
    os.chdir(args.outdir)
    # One process per GPU (slab runs): rank 0 owns args.json, train.log, train.csv and the checkpoints.
    if _is_writer():
        set_log_file(open("train.log", "w"), echo=args.echo)
    else:
        set_log_file(open(os.devnull, "w"), echo=0)
    for name in relpath_args or ():
        if getattr(args, name):
            setattr(args, name, os.path.relpath(getattr(args, name), start=args.outdir))
    for name in ("plot_every", "history_every", "report_every"):
        every = getattr(args, name)
        if every is not None:
            setattr(args, name, max(1, round(every * args.every_factor)))
    if args.epochs is None:
        args.epochs = args.frames * args.plot_every
    if args.seed is not None:
        np.random.seed(args.seed)
        runtime.mod.random.set_seed(args.seed)
    printlog(" ".join(sys.argv))


# --------------------------------------------------------------------------------------------------
# Per-epoch callback
# --------------------------------------------------------------------------------------------------
class EpochCallback:
    """
    `callback(state, epoch, pinfo)` for the optimizers.  Depending on the epoch it reports to the log, appends a
    row to train.csv, plots a frame and writes a checkpoint; user hooks receive this object as their last argument
    (`cbinfo` in the reference's examples) and may read `args problem pinfo history frame epoch walltime
    time_start time_callback task_report task_history task_plot task_checkpoint`.

    `walltime` counts optimizer time only: whatever is spent inside the callback (hooks included) is accumulated
    in `time_callback` and subtracted.
    """

    def __init__(self, problem, args, epoch_func=None, report_func=None, history_func=None, checkpoint_func=None,
                 plot_func=None):
        self.problem, self.args = problem, args
        self._hooks = dict(epoch=epoch_func, report=report_func, history=history_func, checkpoint=checkpoint_func,
                           plot=plot_func)
        self.walltime = 0        # optimizer wall time at the last report
        self.epoch = 0           # epoch of the last report
        self.time_callback = 0   # total time spent in here
        self.time_start = time.time()
        self.frame = 0
        self.pinfo = None
        self.throughput = 0.0
        self.history = History(csvpath="train.csv" if _is_writer() else os.devnull, warmup=1) \
            if args.history_every else None
        self.task_report = self.task_history = self.task_plot = self.task_checkpoint = False
        self.cbinfo = self  # `callback.cbinfo` of the reference

    def _schedule(self, epoch):
        a = self.args
        self.task_report = a.report_every and epoch % a.report_every == 0
        self.task_history = self.history is not None and (epoch % a.history_every == 0 or epoch < a.history_full)
        self.task_plot = epoch % a.plot_every == 0 and (epoch or a.frames)
        self.task_checkpoint = a.checkpoint_every and epoch % a.checkpoint_every == 0
        return self.task_report or self.task_history or self.task_plot or self.task_checkpoint

    @staticmethod
    def _named_norms(pinfo):
        if not pinfo or "norms" not in pinfo:
            return []
        return [(name or str(i), norm) for i, (norm, name) in enumerate(zip(pinfo["norms"], pinfo["names"]))]

    def _report(self, state, epoch, walltime):
        printlog("\nepoch={:05d}".format(epoch))
        norms = self._named_norms(self.pinfo)
        if norms:
            printlog("residual: " + ", ".join("{}:{:.5g}".format(k, float(np.array(v))) for k, v in norms))
        if self._hooks["report"] is not None:
            self._hooks["report"](self.problem, state, epoch, self)
        used, pool = get_gpu_memory_usage_kb()
        printlog("memory: {:} MiB, gpu_used: {:} MiB, gpu_pool: {:} MiB".format(
            get_memory_usage_kb() // 1024, used // 1024, pool // 1024))
        per_epoch = (walltime - self.walltime) / (epoch - self.epoch) if epoch > self.epoch else 0
        cells_per_s = np.prod(self.problem.domain.cshape) / per_epoch if per_epoch > 0 else 0
        printlog("walltime: {:.3f} s, walltime+callback: {:.3f} s, walltime/epoch: {:.3f} ms".format(
            walltime, walltime + self.time_callback, per_epoch * 1000))
        printlog("throughput: {:.3f} Mcells/s".format(cells_per_s / 1e6))
        self.walltime, self.epoch, self.throughput = walltime, epoch, cells_per_s / 1e6

    def _record(self, state, epoch, walltime):
        h, pinfo = self.history, self.pinfo
        used, pool = get_gpu_memory_usage_kb()
        h.append("epoch", epoch)
        h.append("frame", self.frame)
        for key, norm in self._named_norms(pinfo):
            h.append("norm_" + key, np.array(norm))
        if pinfo and "loss" in pinfo:
            h.append("loss", np.array(pinfo["loss"]))
        if getattr(self.args, "linsolver_history", 0) and pinfo and "linsolver" in pinfo:
            for key, val in pinfo["linsolver"].items():
                if isinstance(val, (int, float, str, np.floating)):
                    h.append("lin_" + key, val)
        h.append("walltime", np.round(walltime, 3))
        h.append("memory", get_memory_usage_kb() // 1024)
        h.append("gpu_used", used // 1024)
        h.append("gpu_pool", pool // 1024)
        if self._hooks["history"] is not None:
            self._hooks["history"](self.problem, state, epoch, h, self)
        h.write()

    def _checkpoint(self, state, epoch):
        if self._hooks["checkpoint"] is not None:
            self._hooks["checkpoint"](self.problem, state, epoch, self)
            return
        from .core import checkpoint_save

        path = "checkpoint_{:06d}.pickle".format(epoch)
        printlog(path)
        checkpoint_save(self.problem.domain, state, path)

    def __call__(self, state, epoch, pinfo):
        if self._schedule(epoch):
            cuda = _cuda()
            if cuda:
                cuda.synchronize()  # queued epochs belong to the optimizer's time, not to the callback's
        entered = time.time()
        self.pinfo = pinfo
        if isinstance(self.problem.tracers, dict):
            self.problem.tracers["epoch"] = epoch
        if self._hooks["epoch"] is not None:
            self._hooks["epoch"](self.problem, state, epoch, self)
        now = time.time()
        self.time_callback += now - entered
        walltime = now - self.time_start - self.time_callback
        if self.task_report:
            self._report(state, epoch, walltime)
        if self.task_history:
            self._record(state, epoch, walltime)
        if self.task_plot:
            if self._hooks["plot"] is not None:
                self._hooks["plot"](self.problem, state, epoch, self.frame, self)
            self.frame += 1
        if self.task_checkpoint:
            self._checkpoint(state, epoch)
        self.time_callback += time.time() - now


def make_callback(problem, args=None, epoch_func=None, report_func=None, history_func=None, checkpoint_func=None,
                  plot_func=None):
    """Callback for `optimize`: `report_func(problem, state, epoch, cbinfo)`, `history_func(problem, state, epoch,
    history, cbinfo)`, `plot_func(problem, state, epoch, frame, cbinfo)`, `checkpoint_func` / `epoch_func(problem,
    state, epoch, cbinfo)` are called when their task is due (util.py:337-466)."""
    return EpochCallback(problem, args, epoch_func=epoch_func, report_func=report_func, history_func=history_func,
                         checkpoint_func=checkpoint_func, plot_func=plot_func)
