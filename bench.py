#!/usr/bin/env python3
"""
bench.py -- headline benchmark of the ODIL residual-and-gradient hot path on B200.

Metric (BASELINE.json): Mcells/s = prod(domain.cshape) / (time per epoch) / 1e6, one epoch = one loss+gradient
evaluation (multigrid synthesis -> fused stencil residual/loss/adjoint -> multigrid adjoint) plus the optimizer
update of every unknown -- the reference's own throughput definition (src/odil/util.py:383-386, :408-419), callback
time excluded.  The timed loop is the SHIPPED one: `odil.util.optimize_grad(args, optimizer, problem, state, callback)`;
the callback only places the CUDA events.

  --config 3 (default)  BASELINE configs[3] on N GPUs: 3-D Poisson 512^3, 4-level multigrid, Adam, fp32.
                        N > 1: slabs along axis 0; --scaling weak (default: 512^3 cells PER GPU) or strong (512^3
                        TOTAL, the configuration BASELINE.json names); the weak line also carries a "strong" block.
  --config 1            configs[1]: 2-D Poisson 1024^2, 3-level multigrid, Adam, fp32 (L2-resident: epochs/s matter)
  --config 2            configs[2]: wave inverse (t, x, y) = 256 x 512 x 512, L-BFGS m=50, fp32
  --config 4            configs[4]: 3-D heat inverse 256^3, Newton + matrix-free CG, fp64

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C] [--scaling weak|strong]
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.

--impl reference: the UNMODIFIED reference (core.py + examples/poisson/poisson.py:operator + optimizer.py Adam) on the
host cores through oracle/ref_arm.py (separate process), same config / metric / unit.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mcells/s (residual+grad+optimizer epoch)"
NEWTON_CG = 20  # CG iterations per Newton step in the configs[4] workload


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    p.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4], help="index into BASELINE.json configs")
    p.add_argument("--scaling", type=str, default="weak", choices=["weak", "strong"])
    p.add_argument("--size", type=int, default=None, help="override cells per axis (parity / smoke runs)")
    p.add_argument("--levels", type=int, default=None)
    p.add_argument("--dtype", type=str, default=None, choices=["f32", "f64"])
    p.add_argument("--lr", type=float, default=0.005)
    p.add_argument("--e2e_steps", type=int, default=5)
    p.add_argument("--cpu_steps", type=int, default=2)
    p.add_argument("--cpu_budget_s", type=float, default=420.0, help="wall-time budget of the reference arm")
    p.add_argument("--no_cpu_baseline", action="store_true")
    p.add_argument("--no_strong", action="store_true", help="skip the extra strong-scaling block at N > 1")
    p.add_argument("--extra_configs", type=str, default="1,4,2",
                   help="N=1, default config only: also measure these BASELINE configs (own processes) and embed "
                        "their summaries in the line; empty string disables")
    p.add_argument("--profile", action="store_true",
                   help="short run for ncu: no burn-in, no clock sampler, no e2e / CPU / strong-scaling legs")
    p.add_argument("--graph", type=int, default=None, help="CUDA-graph replay of the Adam epoch (default: config 1)")
    return p.parse_args()


def workload(args, world):
    """Static description of the workload: identical in both arms (the driver compares `config`)."""
    c = args.config
    if c == 3:
        N, L, dt = args.size or 512, args.levels or 4, args.dtype or "f32"
        strong = args.scaling == "strong" and world > 1
        cshape = (N, N, N) if strong else (N * world, N, N)
        per = f"{N}^3 cells in total" if strong else f"{N}^3 cells per GPU"
        return dict(kind="poisson", cshape=cshape, levels=L, dtype=dt, opt="adam",
                    text=f"3D Poisson {cshape[0]}x{N}x{N} ({per}, slabs along axis 0), {L}-level multigrid, "
                         f"Adam lr={args.lr}, {dt}, zero-Dirichlet BC, rhs ~ N(0,1), unknowns start at 0")
    if c == 1:
        N, L, dt = args.size or 1024, args.levels or 3, args.dtype or "f32"
        return dict(kind="poisson", cshape=(N, N), levels=L, dtype=dt, opt="adam",
                    text=f"2D Poisson {N}x{N}, {L}-level multigrid, Adam lr={args.lr}, {dt}, zero-Dirichlet BC, "
                         "rhs ~ N(0,1), unknowns start at 0")
    if c == 2:
        N, dt = args.size or 512, args.dtype or "f32"
        cshape = (N // 2, N, N)
        return dict(kind="wave2", cshape=cshape, levels=0, dtype=dt, opt="lbfgsb",
                    text=f"2D wave inverse (t,x,y) = {cshape[0]}x{N}x{N}, L-BFGS m=50, {dt}, Dirichlet data + initial "
                         "u, u_t from an exact plane-wave solution, unknowns start at 0")
    N, dt = args.size or 256, args.dtype or "f64"
    return dict(kind="heat3", cshape=(N, N, N), levels=0, dtype=dt, opt="newton",
                text=f"3D heat inverse (t,x,y) = {N}^3, k(u) = 0.02 exp(-20 (u-0.5)^2), Newton + {NEWTON_CG} "
                     f"iterations of matrix-free CG on the normal equations per step, {dt}")


def config_block(args, world, wl):
    return {"workload": wl["text"],
            "l2": "inputs exceed L2 (one field is larger than the 126 MB L2; no explicit flush)"
                  if np.prod(wl["cshape"]) * (4 if wl["dtype"] == "f32" else 8) > 126e6 else
                  "working set is L2-resident (126 MB L2): launch-latency-bound, HBM fraction not meaningful",
            "api": "odil.Domain / odil.Problem(operator) / odil.util.optimize_grad (optimize_newton for config 4)",
            "parallelism": f"slab{world}" if world > 1 else "single",
            "baseline_config": args.config}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# Reference arm / CPU baseline: the unmodified reference on the host cores (oracle/ref_arm.py, own process)
# --------------------------------------------------------------------------------------------------
def reference_epochs(size, ndim, levels, dtype, steps, warmup, lr, threads=0, timeout=1500):
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_arm.py"), "--size", str(size), "--ndim", str(ndim),
           "--levels", str(levels), "--dtype", dtype, "--steps", str(steps), "--warmup", str(warmup), "--lr", str(lr),
           "--threads", str(threads)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS"):
        env.pop(k, None)
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("oracle/ref_arm.py failed: " + (r.stderr or r.stdout)[-1500:])
    return json.loads(lines[-1])


def reference_sample(args, wl, steps, warmup):
    """Runs the reference on the workload's own grid when `steps + warmup` epochs fit the time budget, else on the
    largest halved grid that does; returns (result, size, same_size)."""
    if wl["kind"] != "poisson":
        raise RuntimeError("the reference arm covers the Poisson configurations (configs[1], configs[3])")
    ndim = len(wl["cshape"])
    full = wl["cshape"][-1]
    # a one-epoch probe on a small grid fixes the thread count and predicts the cost of the full grid
    probe_n = max(32, min(full, 128 if ndim == 3 else 1024))
    probe = reference_epochs(probe_n, ndim, wl["levels"], wl["dtype"], 1, 1, args.lr)
    per_cell = probe["ms_min"] * 1e-3 / probe_n ** ndim
    size = full
    while size > probe_n and per_cell * size ** ndim * (steps + warmup) * 1.3 > args.cpu_budget_s:
        size //= 2
    res = reference_epochs(size, ndim, wl["levels"], wl["dtype"], steps, warmup, args.lr, threads=probe["threads"])
    return res, size, size == full


def cpu_baseline_block(res, size, same, wl):
    ndim = len(wl["cshape"])
    grid = "x".join([str(size)] * ndim)
    return {"value": res["mcells_per_s"], "unit": "Mcells/s", "cores": res["threads"], "kind": "reference",
            "host_cores": res["host_cores"], "cpu_model": res["cpu_model"], "ms_min": res["ms_min"],
            "ms_median": res["ms_median"], "same_grid_as_workload": bool(same),
            "sample": f"{res['steps']} epochs (+{res['warmup']} warm-up) of {ndim}-D Poisson {grid}, "
                      f"{wl['levels']}-level multigrid, Adam, {wl['dtype']}: the UNMODIFIED reference core.py "
                      f"(multigrid_to_regular / interp 'stack', Context.field) + examples/poisson/poisson.py:operator "
                      f"+ optimizer.py AdamNativeOptimizer from {res['origin']}, arrays on torch-CPU "
                      f"({res['threads']} threads), torch.autograd for jax.value_and_grad; JAX/TF are not installable "
                      "in this image"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    wl = workload(args, world)
    res, size, same = reference_sample(args, wl, args.steps, args.warmup)
    value = res["mcells_per_s"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": config_block(args, world, wl),
        "cpu_baseline": cpu_baseline_block(res, size, same, wl),
        "e2e": {"value": value, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def poisson_operator(ctx):
    """Zero-Dirichlet Poisson residual, written against the ODIL API like the reference example
    (examples/poisson/poisson.py:57-68, :89-123)."""
    import odil

    mod, ndim = ctx.mod, ctx.domain.ndim
    h, idx, n = ctx.step(), ctx.indices(), ctx.size()
    u = ctx.field("u")
    zero = mod.cast(0, u.dtype)
    ex = odil.core.extrap_quadh
    res = -ctx.extra.rhs
    for a in range(ndim):
        e = [1 if b == a else 0 for b in range(ndim)]
        um, up = ctx.field("u", *[-s for s in e]), ctx.field("u", *e)
        um2 = mod.where(idx[a] == 0, ex(up, u, zero), um)
        up2 = mod.where(idx[a] == n[a] - 1, ex(um, u, zero), up)
        res = res + (up2 - 2 * u + um2) / h[a] ** 2
    return [res]


def run_args(**kw):
    d = dict(epochs=10, epoch_start=0, lr=0.005, callback_update_state=0, bfgs_m=None, bfgs_pgtol=None,
             bfgs_maxls=None, adam_epsilon=None, adam_beta_1=None, adam_beta_2=None)
    d.update(kw)
    return argparse.Namespace(**d)


def make_problem(wl, lr):
    import torch

    import odil

    npdt = np.float32 if wl["dtype"] == "f32" else np.float64
    tdt = torch.float32 if wl["dtype"] == "f32" else torch.float64
    cshape = wl["cshape"]
    if wl["kind"] == "poisson":
        L = wl["levels"]
        domain = odil.Domain(cshape=cshape, dimnames=["x", "y", "z"][:len(cshape)], multigrid=L > 0,
                             mg_nlvl=L if L > 0 else None, dtype=npdt)
        gen = torch.Generator(device="cuda").manual_seed(0)
        rhs = odil.backend.Known(torch.randn(cshape, dtype=tdt, device="cuda", generator=gen))
        state = odil.State()
        state.fields["u"] = None
        state = domain.init_state(state)
        problem = odil.Problem(poisson_operator, domain, argparse.Namespace(rhs=rhs))
        problem._engine(state)  # trace + plans now; drops the global constant in slab mode
        del rhs
        problem.extra.rhs = None
        torch.cuda.empty_cache()
        return problem, state, {}
    if wl["kind"] == "wave2":
        from tests import operators as ops

        problem, state = ops.make_wave2(cshape, npdt)
        return problem, state, dict(bfgs_m=50)
    if wl["kind"] == "heat3":
        from tests import nonaffine_cases as cases

        operator, domain, state, extra, tracers = cases.make_heat3(odil, odil.runtime.mod, npdt, cshape,
                                                                   device_data=True)
        problem = odil.Problem(operator, domain, extra, tracers=dict(tracers))
        # one step = one Newton iteration: linearize (generated kernels) + NEWTON_CG iterations of matrix-free CG on
        # the normal equations (two generated Jacobian-product kernels each) + state update + loss evaluation
        return problem, state, dict(linsolver="cg_b200", linsolver_tol=0.0, linsolver_maxiter=NEWTON_CG,
                                    linsolver_damp=0.0, linsolver_verbose=0, linsolver_history=0)
    raise SystemExit(f"workload {wl['kind']} is not available in this build")


class Stepper:
    """Runs `odil.util.optimize_grad` once for warm-up + timed epochs; its callback brackets the timed epochs with
    barrier + synchronize + CUDA events and (e2e mode) performs the per-step host<->device copies."""

    def __init__(self, problem, state, wl, lr, extra_args, dist):
        self.problem, self.state, self.wl, self.lr, self.extra, self.dist = problem, state, wl, lr, extra_args, dist

    def sync_all(self):
        import torch

        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize()

    def run(self, warmup, steps, on_warm=None, e2e_host=None, graph=False):
        import torch

        import odil

        domain = self.problem.domain
        ev = {}
        info = {"loss": None, "wall0": None, "wall1": None}

        def callback(state, epoch, pinfo):
            if epoch == warmup:
                if on_warm is not None:
                    on_warm()
                self.sync_all()
                ev["e0"] = torch.cuda.Event(enable_timing=True)
                ev["e0"].record()
                info["wall0"] = time.perf_counter()
            if e2e_host is not None and epoch >= warmup:
                if epoch > warmup:
                    info["loss"] = float(pinfo["loss"])  # the finished step's result goes back to the host
                if epoch < warmup + steps:
                    # the next step's inputs arrive from pinned host memory
                    for d, h in zip(domain.arrays_from_state(state), e2e_host):
                        d.copy_(h, non_blocking=True)
            if epoch == warmup + steps:
                ev["e1"] = torch.cuda.Event(enable_timing=True)
                ev["e1"].record()
                torch.cuda.synchronize()
                info["wall1"] = time.perf_counter()
                info["pinfo"] = pinfo
                self.sync_all()

        a = run_args(epochs=warmup + steps, lr=self.lr, **self.extra)
        if graph is None:
            os.environ.pop("ODIL_B200_GRAPH", None)  # the package's own default (replay where launch-bound and safe)
        else:
            os.environ["ODIL_B200_GRAPH"] = "1" if graph else "0"
        try:
            if self.wl["opt"] == "newton":
                odil.util.optimize_newton(a, self.problem, self.state, callback)
            else:
                odil.util.optimize_grad(a, self.wl["opt"], self.problem, self.state, callback)
        except odil.EarlyStopError:
            pass
        if "e1" not in ev:
            raise RuntimeError("the optimizer stopped before the timed epochs completed")
        return ev["e0"].elapsed_time(ev["e1"]) / steps, (info["wall1"] - info["wall0"]) / steps * 1e3, info


def slab_parity(world, rank):
    """Small slab-vs-oracle check run before timing when WORLD_SIZE > 1: loss and every multigrid-level gradient of
    a (32*W, 16, 24) 3-level Poisson problem evaluated on the slabs against oracle.eval_loss_grad_plan on rank 0."""
    import torch

    import odil
    from oracle import odil_oracle as orc
    from tests import operators as ops

    out = {"world": world, "cshape": [32 * world, 16, 24], "levels": 3}
    for dt, key in [(np.float64, "max_rel_err_f64"), (np.float32, "max_rel_err_f32")]:
        cshape, nlvl = (32 * world, 16, 24), 3
        problem, state = ops.make_poisson(cshape, nlvl, dt)
        domain = problem.domain
        rng = np.random.default_rng(0)
        terms = [rng.standard_normal(cs).astype(dt) for cs in domain.mg_cshapes]
        arrays = [domain.slab.scatter(torch.as_tensor(t, device="cuda")) for t in terms]
        domain.arrays_to_state(arrays, state)
        loss, grads, _, _, _ = problem.eval_loss_grad(state)
        gathered = [domain.slab.gather(g).cpu().numpy() for g in grads]
        lossv = float(loss)
        if rank == 0:
            steps = [dt(1) / dt(n) for n in cshape]
            offsets, table, rr = orc.poisson_plan(3, steps)
            rhs = np.asarray(problem.extra.rhs)
            loss_ref, grads_ref, _, _ = orc.eval_loss_grad_plan([t.astype(np.float64) for t in terms], "ccc", offsets,
                                                                table, rr, -rhs.astype(np.float64))
            err = abs(lossv - loss_ref) / abs(loss_ref)
            for g, gr in zip(gathered, grads_ref):
                err = max(err, float(np.max(np.abs(g - gr)) / np.max(np.abs(gr))))
            out[key] = err
    if rank == 0:
        out["ok"] = bool(out["max_rel_err_f64"] < 1e-11 and out["max_rel_err_f32"] < 5e-6)
    return out


def other_configs(which):
    """The other single-GPU configurations of BASELINE.json, each measured by `bench.py --config C` in its own
    process (isolated from the headline measurement) and summarised: secondary lines, not the contract line."""
    steps = {1: ("200", "5"), 2: ("10", "5"), 4: ("3", "3")}
    out = {}
    for c in which:
        if c not in steps:
            continue
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--config", str(c), "--steps", steps[c][0], "--warmup",
               steps[c][1], "--e2e_steps", "2", "--no_cpu_baseline", "--extra_configs", ""]
        try:
            r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=240)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            d = json.loads(lines[-1])
            out[str(c)] = {"workload": d["config"]["workload"], "value": d["value"], "unit": d["unit"],
                           "ms_per_step": d["ms_per_step"], "steps": d["steps"], "dtype": d["dtype"],
                           "graph_replay": d.get("graph_replay"), "gpu_launches": d.get("gpu_launches"),
                           "dominant_kernel": d["roofline"]["kernel"], "dominant_kernel_frac": d["roofline"]["frac"],
                           "e2e": d["e2e"]["value"] if d.get("e2e") else None, "clocks": d.get("clocks"),
                           "final_loss": d.get("final_loss")}
        except Exception as exc:
            out[str(c)] = {"error": repr(exc)[:300]}
    return out


def run_b200(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import odil  # noqa: F401
    from odil_b200 import native

    native.load()
    wl = workload(args, world)
    es = 4 if wl["dtype"] == "f32" else 8
    if world > 1 and wl["kind"] != "poisson":
        raise SystemExit("only the Poisson configurations are slab-decomposed")

    parity = slab_parity(world, rank) if world > 1 else None

    def measure(wl, steps, warmup, with_kernels):
        problem, state, extra = make_problem(wl, args.lr)
        domain = problem.domain
        assert (domain.slab is not None) == (world > 1)
        ncells = int(np.prod(wl["cshape"]))
        stepper = Stepper(problem, state, wl, args.lr, extra, dist)
        from odil_b200 import optimizer as _opt

        graph = bool(args.graph) if args.graph is not None else None
        timers = {}

        def timed(name, fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            timers.setdefault(name, []).append((e0, e1))
            return r

        sampler = ClockSampler(local_rank)
        launches = {}

        def on_warm():
            # keep the device busy until nvidia-smi has produced its first lines, then start counting
            timers.clear()
            launches["n0"] = native.launch_count()

        if rank == 0 and with_kernels and not args.profile:
            sampler.start()
        # burn-in (untimed): traces the operator, builds plans and work lists, and keeps the device under load long
        # enough for nvidia-smi to report clocks before and during the timed region
        ms0, _, _ = stepper.run(1, 2, graph=graph) if args.profile else stepper.run(3, 5, graph=graph)
        if dist is not None:  # every rank must run the same number of epochs (halo exchanges are collective)
            t_ = torch.tensor([ms0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms0 = t_.item()
        if not args.profile:
            stepper.run(1, int(min(400, max(10, 500.0 / max(ms0, 1e-3)))), graph=graph)
        replayed = bool(_opt.LAST_RUN_INFO["graph"]) and wl["opt"] == "adam"
        if with_kernels and not replayed:  # per-call CUDA events cannot be recorded inside a graph capture
            native.set_timer_hook(timed)
        ms, _, info = stepper.run(warmup if args.profile else max(warmup, 3), steps, on_warm=on_warm, graph=graph)
        native.set_timer_hook(None)
        n_launch = native.launch_count() - launches["n0"]
        clocks = sampler.stop() if rank == 0 and with_kernels and not args.profile else None
        if dist is not None:
            t_ = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms = t_.item()
        kern = {}
        for name, evs in timers.items():
            per_step = sum(a.elapsed_time(b) for a, b in evs) / steps
            kern[name] = {"ms_per_step": per_step, "calls_per_step": len(evs) / steps}
        return dict(problem=problem, state=state, stepper=stepper, ms=ms, ncells=ncells, kern=kern, clocks=clocks,
                    launches=n_launch, loss=float(info["pinfo"]["loss"]), graph=replayed)

    m = measure(wl, args.steps, args.warmup, True)
    problem, state, ncells, ms = m["problem"], m["state"], m["ncells"], m["ms"]
    domain = problem.domain
    x = domain.arrays_from_state(state)
    ncells_local = ncells // world
    nunk_local = sum(a.numel() for a in x)
    peak, peak_src = measured_peak()
    kern = m["kern"]
    alg_bytes = {"stencil_fused": 3 * es * ncells_local, "adam_step": 7 * es * nunk_local}
    alg_note = {"stencil_fused": "3*s bytes per cell (read U, read c, write g)"}
    if "adam_synth" in kern:
        # the finest multigrid term is updated by odil_b200_adam_synth (reads t0, m, v, g and 1/8 coarse value per cell,
        # writes t0, m, v, U); odil_b200_adam_step then only covers the coarser terms
        alg_bytes["adam_synth"] = (8 + 1 / 8) * es * ncells_local
        alg_bytes["adam_step"] = 7 * es * max(nunk_local - ncells_local, 0)
        alg_note["adam_synth"] = "(8 + 1/8)*s bytes per finest-level cell (read t0, m, v, g, coarse; write t0, m, v, U)"
    if wl["kind"] == "heat3":
        # generated kernels of the heat operator (one field u, constants imp_mask and imp_u, two outputs)
        alg_bytes.update({"jit:k_g0_jvp": 6 * es * ncells_local, "jit:k_g0_vjp": 6 * es * ncells_local,
                          "jit:k_g0_lossgrad": 4 * es * ncells_local, "jit:k_g0_values": 5 * es * ncells_local})
        alg_note.update({"jit:k_g0_jvp": "6*s bytes per cell (read u, tangent, imp_mask, imp_u; write 2 outputs)",
                         "jit:k_g0_vjp": "6*s bytes per cell (read u, 2 cotangents, imp_mask, imp_u; write g)",
                         "jit:k_g0_lossgrad": "4*s bytes per cell (read u, imp_mask, imp_u; write g)",
                         "jit:k_g0_values": "5*s bytes per cell (read u, imp_mask, imp_u; write 2 outputs)"})
        # Jacobian products from the diagonals stored once per Newton step (11 structurally non-zero (output, load)
        # pairs: 10 for the heat residual, 1 for the imposed-data term)
        alg_bytes.update({"jit:k_g0_jvpd": 14 * es * ncells_local, "jit:k_g0_vjpd": 14 * es * ncells_local,
                          "jit:k_g0_vjpg": 14 * es * ncells_local, "jit:k_g0_jacd": 14 * es * ncells_local})
        alg_note.update({"jit:k_g0_jvpd": "14*s bytes per cell (read 11 stored diagonals and the tangent; write 2 outputs)",
                         "jit:k_g0_vjpd": "14*s bytes per cell (read 11 stored diagonals and 2 cotangents; accumulate g)",
                         "jit:k_g0_vjpg": "14*s bytes per cell (read 11 stored diagonals and 2 cotangents; write g)",
                         "jit:k_g0_jacd": "14*s bytes per cell (read u, imp_mask, imp_u; write 11 diagonals)"})
    for name, k in kern.items():
        if name in alg_bytes and k["ms_per_step"] > 0:
            gbs = alg_bytes[name] / (k["ms_per_step"] / max(k["calls_per_step"], 1) * 1e-3) / 1e9
            k.update({"achieved_GBs": gbs, "frac": gbs / peak})
    # the dominant kernel of the step among those with a byte model (the fused sweep for the Poisson / wave configs)
    cand = [n for n in kern if n in alg_bytes and n != "adam_step"]
    top = "stencil_fused" if "stencil_fused" in kern else (max(cand, key=lambda n: kern[n]["ms_per_step"]) if cand
                                                             else "stencil_fused")
    fused = kern.get(top, {})
    launches_first, clocks_first, loss_first, graph_first = m["launches"], m["clocks"], m["loss"], m["graph"]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "fused_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("cells") == ncells_local and tj.get("dtype") == wl["dtype"] and top == "stencil_fused":
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm",
        "kernel": ("odil_b200_stencil_fused (residual + loss + adjoint gradient in one sweep, + reduce)"
                   if top == "stencil_fused" else f"{top} (NVRTC-generated kernel of the traced operator)"),
        "achieved": fused.get("achieved_GBs"), "peak": peak, "unit": "GB/s", "frac": fused.get("frac"),
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes.get(top),
        "ms_per_launch": (fused.get("ms_per_step") or 0) / max(fused.get("calls_per_step", 1), 1) or None,
        "note": alg_note.get(top, "") + "; time = CUDA events around the C-ABI call on the launch stream, averaged "
                "over the timed steps (rank 0); traffic = dram bytes of one ncu --set full capture of the same launch "
                "(profiles/fused_traffic.json)",
    }

    # e2e: every step the unknowns arrive from pinned host memory and the loss goes back to the host
    e2e = None
    if args.profile:
        args.no_cpu_baseline = args.no_strong = True
    elif wl["opt"] == "adam":
        host = [torch.empty(a.shape, dtype=a.dtype, pin_memory=True).copy_(a) for a in x]
        h2d = sum(a.numel() * a.element_size() for a in host)
        _, wall_ms, _ = m["stepper"].run(2, args.e2e_steps, e2e_host=host)
        if dist is not None:
            t_ = torch.tensor([wall_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            wall_ms = t_.item()
        e2e = {"value": ncells / (wall_ms * 1e-3) / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": wall_ms,
               "note": "wall clock through odil.util.optimize_grad; per step and rank: H2D of all multigrid terms from "
                       "pinned host memory + epoch + D2H of the loss"}
        del host
    else:
        # the optimizer owns its iterate: whole job from host arrays to a host result
        host = [a.detach().cpu().pin_memory() for a in x]
        h2d = sum(a.numel() * a.element_size() for a in host)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for d, h in zip(x, host):
            d.copy_(h, non_blocking=True)
        _, _, info = m["stepper"].run(0, args.e2e_steps)
        back = [a.detach().cpu() for a in domain.arrays_from_state(state)]
        wall_ms = (time.perf_counter() - w0) / args.e2e_steps * 1e3
        e2e = {"value": ncells / (wall_ms * 1e-3) / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": h2d / args.e2e_steps,
               "d2h_bytes_per_step": h2d / args.e2e_steps + 8, "ms_per_step": wall_ms,
               "note": f"wall clock of a whole {args.e2e_steps}-iteration job: H2D of the initial state, the "
                       "iterations with the loss read back every iteration, D2H of the final state (bytes amortised)"}
        del host, back

    strong = None
    if world > 1 and args.config == 3 and args.scaling == "weak" and not args.no_strong:
        # BASELINE configs[3] as named: 512^3 in TOTAL over the N GPUs
        del m, problem, state, x
        torch.cuda.empty_cache()
        a2 = argparse.Namespace(**vars(args))
        a2.scaling = "strong"
        wl2 = workload(a2, world)
        try:
            m2 = measure(wl2, args.steps, args.warmup, False)
            strong = {"value": m2["ncells"] / (m2["ms"] * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": m2["ms"],
                      "workload": wl2["text"], "scaling": "strong", "final_loss": m2["loss"],
                      "graph_replay": m2["graph"]}
        except Exception as exc:  # the extra block must never cost the headline line
            strong = {"error": repr(exc)[:300], "workload": wl2["text"], "scaling": "strong"}

    others = None
    if rank == 0 and world == 1 and args.config == 3 and args.extra_configs and not args.profile \
            and args.size is None:
        others = other_configs([int(c) for c in args.extra_configs.split(",") if c.strip()])
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and wl["kind"] == "poisson":
        try:
            res, size, same = reference_sample(args, wl, args.cpu_steps, 1)
            cpu = cpu_baseline_block(res, size, same, wl)
        except Exception as exc:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "Mcells/s", "cores": 0, "kind": "reference", "sample": f"failed: {exc}"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": ncells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": config_block(args, world, wl),
            "roofline": roofline, "kernels": kern, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches_first, "clocks": clocks_first, "final_loss": loss_first,
            "graph_replay": graph_first,
            "epoch_traffic_model": {"compulsory_bytes_per_cell": (6 * nunk_local / ncells_local + 1) * es,
                                    "epoch_frac_of_peak": (6 * nunk_local + ncells_local) * es / (ms * 1e-3) / 1e9 / peak},
        }
        if others is not None:
            line["other_configs"] = others
        if parity is not None:
            line["parity"] = parity
        if strong is not None:
            line["strong"] = strong
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    main()
