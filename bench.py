#!/usr/bin/env python3
"""
bench.py -- headline benchmark of the ODIL residual-and-gradient hot path on B200.

Metric (BASELINE.json): Mcells/s = prod(domain.cshape) / (time per epoch) / 1e6, one epoch = one
loss+gradient evaluation (multigrid synthesis -> fused stencil residual/loss/adjoint -> multigrid adjoint)
plus the Adam update of every multigrid term -- the reference's own throughput definition
(src/odil/util.py:383-386, :408-419), callback time excluded.

Workload at N=1: 3-D Poisson 512^3, 4-level multigrid, Adam, fp32 (BASELINE.json configs[3] on one GPU;
the configuration the metric is quoted on).  At N>1 the grid is slab-decomposed along axis 0
(weak scaling: 512^3 cells per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 512] [--levels 4]
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
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


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    p.add_argument("--size", type=int, default=512, help="cells per axis (per GPU along axis 0)")
    p.add_argument("--levels", type=int, default=4)
    p.add_argument("--dtype", type=str, default="f32", choices=["f32", "f64"])
    p.add_argument("--lr", type=float, default=0.005)
    p.add_argument("--e2e_steps", type=int, default=5)
    p.add_argument("--cpu_size", type=int, default=192, help="cells per axis of the bounded CPU sample")
    p.add_argument("--cpu_steps", type=int, default=3)
    p.add_argument("--no_cpu_baseline", action="store_true")
    return p.parse_args()


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
# Reference arm / CPU baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_epoch_rate(size, levels, steps, warmup, dtype):
    import torch

    from oracle import ref_port_torch as port

    ncpu = os.cpu_count()
    tdt = torch.float32 if dtype == "f32" else torch.float64
    # Use the thread count that serves the reference best on this host (oversubscribing a many-core box
    # with small elementwise ops is much slower than a moderate count): quick probe on a 96^3 grid.
    probe = port.PoissonAdamEpoch((96,) * 3, min(levels, 4), dtype=tdt)
    best, cores = None, ncpu
    for nthr in sorted({min(ncpu, 64), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True):
        torch.set_num_threads(nthr)
        probe.step()
        t0 = time.perf_counter()
        probe.step()
        dt1 = time.perf_counter() - t0
        if best is None or dt1 < best:
            best, cores = dt1, nthr
    del probe
    torch.set_num_threads(cores)
    ep = port.PoissonAdamEpoch((size,) * 3, levels, dtype=tdt)
    for _ in range(max(1, warmup)):
        ep.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        ep.step()
    dt = (time.perf_counter() - t0) / steps
    return size ** 3 / dt / 1e6, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    value, dt, cores = cpu_epoch_rate(args.cpu_size, args.levels, args.steps, args.warmup, args.dtype)
    sample = (f"3-D Poisson {args.cpu_size}^3 (bounded sample of the {args.size}^3 workload), {args.levels}-level "
              f"multigrid, Adam, {args.dtype}: oracle/ref_port_torch.py = the reference's per-epoch array ops "
              f"(interp 'stack', roll, where, mean(square), reverse-mode AD, Adam) on torch-CPU, {cores} threads; "
              "JAX-CPU itself is not installable in this image")
    line = {
        "impl": "reference", "metric": "Mcells/s (residual+grad+Adam epoch), 3D Poisson, 4-level multigrid",
        "value": value, "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"3D Poisson {args.size}^3/GPU, {args.levels}-level multigrid, Adam, {args.dtype}; "
                               f"timed on a {args.cpu_size}^3 sample"},
        "cpu_baseline": {"value": value, "unit": "Mcells/s", "cores": cores, "kind": "port", "sample": sample},
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


def run_b200(args):
    import torch

    import odil
    from odil_b200 import native
    from odil_b200.optimizer import adam_scalars

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
    native.load()
    npdt = np.float32 if args.dtype == "f32" else np.float64
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    es = 4 if args.dtype == "f32" else 8
    N = args.size
    # weak scaling: every GPU owns a slab of N planes of the (N*world, N, N) grid
    cshape = (N * world, N, N)
    domain = odil.Domain(cshape=cshape, dimnames=["x", "y", "z"], multigrid=True, mg_nlvl=args.levels, dtype=npdt)
    assert (domain.slab is not None) == (world > 1)
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
    x = domain.arrays_from_state(state)
    m = [torch.zeros_like(a) for a in x]
    v = [torch.zeros_like(a) for a in x]
    eps = float(npdt(1e-7))
    ncells = int(np.prod(cshape))
    ncells_local = ncells // world
    nunk_local = sum(a.numel() for a in x)

    timers = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        timers.setdefault(name, []).append((e0, e1))
        return r

    def epoch(t):
        domain.arrays_to_state(x, state)
        loss, grads, terms, names, norms = problem.eval_loss_grad(state)
        alpha, omb1, omb2 = adam_scalars(args.lr, 0.9, 0.999, t, npdt)
        native.adam_step(x, m, v, grads, alpha, omb1, omb2, eps)
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(value):
        if dist is None:
            return value
        t_ = torch.tensor([value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return t_.item()

    native.set_timer_hook(timed)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.2 s to start: sample from the warm-up on, all of it under load
    t = 0
    for _ in range(max(args.warmup, 3)):
        t += 1
        epoch(t)
    sync_all()
    if rank == 0:
        # keep the device busy until the sampler has produced its first lines (not timed)
        t_wait = time.time()
        while len(sampler.lines) < 2 and time.time() - t_wait < 3.0:
            t += 1
            epoch(t)
            torch.cuda.synchronize()
    sync_all()
    timers.clear()
    launches0 = native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        t += 1
        loss = epoch(t)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    launches = native.launch_count() - launches0
    loss_val = float(loss)

    peak, peak_src = measured_peak()
    kern = {}
    alg_bytes = {"stencil_fused": 3 * es * ncells_local, "adam_step": 7 * es * nunk_local}
    for name, evs in timers.items():
        per_step = sum(a.elapsed_time(b) for a, b in evs) / args.steps
        kern[name] = {"ms_per_step": per_step, "calls_per_step": len(evs) / args.steps}
        if name in alg_bytes:
            gbs = alg_bytes[name] / (per_step * 1e-3) / 1e9
            kern[name].update({"achieved_GBs": gbs, "frac": gbs / peak})
    fused = kern.get("stencil_fused", {})
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fused_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("cells") == ncells_local and tj.get("dtype") == args.dtype:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm",
        "kernel": "odil_b200_stencil_fused (k_star8: TMA-fed residual + loss + adjoint gradient in one sweep, + reduce)",
        "achieved": fused.get("achieved_GBs"), "peak": peak, "unit": "GB/s", "frac": fused.get("frac"),
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes["stencil_fused"], "ms_per_launch": fused.get("ms_per_step"),
        "note": "3*s bytes per cell (read U, read c, write g); time = CUDA events around the C-ABI call on the "
                "launch stream, averaged over the timed steps (rank 0)",
    }
    native.set_timer_hook(None)

    # e2e: every step the unknowns arrive from pinned host memory and the loss goes back to the host
    host = [torch.empty(a.shape, dtype=a.dtype, pin_memory=True).copy_(a) for a in x]
    h2d = sum(a.numel() * a.element_size() for a in host)
    sync_all()
    w0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        t += 1
        for d, h in zip(x, host):
            d.copy_(h, non_blocking=True)
        loss = epoch(t)
        _ = float(loss)  # device -> host read of the step's result
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - w0) / args.e2e_steps * 1e3)
    e2e = {"value": ncells / (e2e_ms * 1e-3) / 1e6, "unit": "Mcells/s", "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": 8 * world, "ms_per_step": e2e_ms,
           "note": "per step and rank: H2D of all multigrid terms from pinned host memory + epoch + D2H of the loss"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cv, cdt, cores = cpu_epoch_rate(args.cpu_size, args.levels, args.cpu_steps, 1, args.dtype)
        cpu = {"value": cv, "unit": "Mcells/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} epochs of 3-D Poisson {args.cpu_size}^3, {args.levels}-level multigrid, "
                         f"Adam, {args.dtype}, oracle/ref_port_torch.py on torch-CPU ({cores} threads)"}
    if rank == 0:
        line = {
            "metric": "Mcells/s (residual+grad+Adam epoch), 3D Poisson, 4-level multigrid",
            "value": ncells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"3D Poisson {cshape[0]}x{N}x{N} ({N}^3 cells per GPU, slabs along axis 0), "
                                   f"{args.levels}-level multigrid, Adam lr={args.lr}, {args.dtype}, zero-Dirichlet BC, "
                                   "rhs ~ N(0,1), unknowns start at 0",
                       "l2": f"inputs exceed L2 ({ncells_local * es / 2**20:.0f} MiB per field vs 126 MB L2); "
                             "no explicit flush",
                       "api": "odil.Domain / odil.Problem(operator).eval_loss_grad + odil_b200_adam_step",
                       "parallelism": f"slab{world}" if world > 1 else "single"},
            "roofline": roofline, "kernels": kern, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "final_loss": loss_val,
        }
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
