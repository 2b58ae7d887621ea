#!/usr/bin/env python
"""bench.py -- forward+adjoint design evaluations per second (BASELINE.json metric).

Workload (config.workload): cfg3 of BASELINE.json, the quads_focusing random-initial-guess ensemble
(24x16 quads, contact, pulse drive, n_t=200, rtol 1e-8 / atol 1e-4; designs = initial design +
U(-1,1)*0.15*spacing, one numpy PRNG stream per design).  One "step" = one forward solve + one
adjoint solve (objective: target kinetic energy) of every design of the ensemble.  The ensemble of
`--designs` designs (default 1024, BASELINE.json configs[2]) is SHARDED over the ranks
(`parallel.shard_range`: contiguous slices, no data-path collective): strong scaling, the stated
configuration.  `--designs-per-gpu D` instead gives every rank D designs of its own (weak scaling).

  value      designs/s with all inputs resident in HBM (forward kernel + objective cotangent + adjoint kernel)
  e2e        the same through the public API (DynamicSolver.odeint + torch.autograd) with the per-design leaves
             copied host->device from pinned memory and objective values + gradients read back, every step
  roofline   adjoint kernel: algorithmic FP64 flops (convention W, with the step counts
             the kernel reports) / CUDA-event duration, against the FP64-FMA peak measured in this run;
             frac_executed rescales it to the DFMA/DADD/DMUL instructions the kernel really executes (ratio
             from the committed ncu counters, profiles/r02_fp64_instruction_counts.json)
  cpu_baseline  the C++ CPU oracle (a port of the reference algorithm, NOT the reference) on the host cores

`--impl reference` times the CPU implementation of the path (oracle port: JAX is not installable in
this image) on the same config.
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

METRIC = "forward+adjoint design evaluations per second"
UNIT = "designs/s"


def flops_model(spec, aug_size, n_t, fwd_steps, bwd_steps):
    """SURVEY section 8(d), convention W.  Returns (forward flops, adjoint flops) for the given totals of
    attempted steps (summed over designs; per-design terms scale with the number of designs in them)."""
    f_rhs = 144 * spec.n_bonds + 41 * spec.n_blocks + 3 * spec.n_free
    N = 2 * spec.n_free

    def fwd(steps, designs):
        return steps * (6 * f_rhs + 65 * N) + designs * 50 * N * n_t

    def bwd(steps, designs):
        return steps * (24 * f_rhs + 65 * aug_size) + designs * (n_t - 1) * (8 * f_rhs + 60 * aug_size)
    return fwd, bwd


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line), sampled once a second through
    NVML inside this process; the maximum clock is read once (that query alone takes up to 75 ms, tools/nvml_probe.py) and
    the power every fourth sample.  Management queries share driver locks with the CUDA calls of the measured thread:
    spawning `nvidia-smi` five times a second stalled timed end-to-end steps by hundreds of milliseconds (round 1), four NVML
    queries every 0.2 s still did now and then (tools/e2e_jitter.py: the device time of a step is constant to 0.1 %, the
    waits are on the host).  nvidia-smi remains the fallback when the NVML binding is missing."""

    def __init__(self, index):
        self.rows, self._stop, self.index = [], threading.Event(), index
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES renumbers CUDA devices, NVML does not: map through the UUID-free common case
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self._max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._power, self._k = "", 0
            self.source = "nvml"
        except Exception:
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        if self._k % 4 == 0:
            self._power = f"{n.nvmlDeviceGetPowerUsage(self._h) / 1000.0:.1f}"
        self._k += 1
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        flag = lambda bit: "Active" if r & bit else "Not Active"  # noqa: E731
        return [str(sm), str(self._max), self._power, flag(n.nvmlClocksEventReasonHwSlowdown), flag(n.nvmlClocksEventReasonHwThermalSlowdown),
                flag(n.nvmlClocksEventReasonSwThermalSlowdown), flag(n.nvmlClocksEventReasonSwPowerCap)]

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(1.0 if self._nvml is not None else 2.0)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": self.source}


HORIZON_SCALE = 1.0  # < 1 only for profiling runs (ncu replays every launch dozens of times)


def build_problem(n_designs, seed0):
    import torch
    from difflexmm_b200.problems import QuadsFocusing
    prob = QuadsFocusing()
    if HORIZON_SCALE != 1.0:
        prob.simulation_time *= HORIZON_SCALE
        prob.n_timepoints = max(2, int(round(prob.n_timepoints * HORIZON_SCALE)))
    spec, drive = prob.lower()
    hs, vs = prob.random_ensemble(n_designs, noise=0.15, seed0=seed0)
    leaves, pb, dpd, aug, y0, ts = prob.boundary_inputs((hs, vs), batch=n_designs, device="cpu")
    prob.ensemble = (hs, vs)
    return prob, spec, drive, leaves, pb, dpd, aug, y0, ts


def target_free_index(prob, spec):
    """free-DOF indices of the 3 DOFs of every target block (all free in this configuration)."""
    free_of = -np.ones(3 * spec.n_blocks, dtype=np.int64)
    free_of[spec.free_dofs] = np.arange(spec.n_free)
    idx = free_of[(prob.target_blocks()[:, None] * 3 + np.arange(3)[None]).reshape(-1)]
    assert (idx >= 0).all()
    return idx


def run_cpu(args, rank_out=True):
    """CPU implementation of the path (C++ oracle port of the reference algorithm), all host cores."""
    import oracle
    from oracle import Oracle
    fast = oracle.use_fast_build(True)  # -O3 -march=x86-64-v3 (AVX2 + FMA) when the host has them, else the -O2 checker build
    cores = os.cpu_count() or 1
    # 64 designs (four per host thread on the 16-thread boxes): 10-25 s of CPU work per step; a long reference run
    # (--impl reference with more than 10 steps + warm-ups) takes 32 per step so that the whole run stays within a few minutes
    long_run = args.impl == "reference" and args.steps + args.warmup > 10
    n = args.cpu_designs if args.cpu_designs > 0 else max(32 if long_run else 64, 2 * max(cores, 1))
    prob, spec, drive, leaves, pb, dpd, aug, y0, ts = build_problem(n, seed0=0)
    orc = Oracle(spec)
    lv = {k: v.numpy() for k, v in leaves.items()}
    ps = orc.params(n, lv, pb, dpd)
    nf = spec.n_free
    tidx = target_free_index(prob, spec)

    def step():
        ys, _ = orc.forward(ps, y0.numpy(), ts.numpy(), prob.rtol, prob.atol, n_threads=cores)
        g = np.zeros_like(ys)
        g[:, :, nf + tidx] = ys[:, :, nf + tidx] * lv["inertia"][:, None, tidx]
        orc.adjoint(ps, ys, ts.numpy(), g, prob.rtol, prob.atol, aug, n_threads=cores)
    step.build = "g++ -O3 -march=x86-64-v3 (AVX2, FMA)" if fast else "g++ -O2"
    return step, n, cores


def run_cfg4(args, rank, world, local_rank):
    """BASELINE.json configs[3]: quads 24x18 static tuning, two tasks (compressive strain 0.01 / 0.08, weights 0.75 / -0.25)
    of ONE design; the tasks are dealt to the ranks, weights applied locally, one NCCL all-reduce of [objective | design
    gradient] per evaluation (reference problems/quads_kinetic_energy_static_tuning.py:473-478).  A step = one weighted
    value-and-gradient evaluation through the public API with the design in pinned host memory (so value == e2e)."""
    import torch
    import torch.distributed as dist
    from difflexmm_b200.parallel import multitask_value_and_grad, shard_range
    from difflexmm_b200.problems import QuadsStaticTuning
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tasks, weights = [dict(compressive_strain=0.01), dict(compressive_strain=0.08)], [0.75, -0.25]
    kw = {}
    if HORIZON_SCALE != 1.0:
        kw = dict(simulation_time_dynamic=2.0 / 30.0 * HORIZON_SCALE, n_timepoints=max(3, int(200 * HORIZON_SCALE)),
                  compressive_strain_rate=0.25 / HORIZON_SCALE)
    probs = [QuadsStaticTuning(**t, **kw) for t in tasks]
    b, e = shard_range(len(tasks), rank, world)
    for i in range(b, e):  # a rank only sets up the solvers of its own tasks
        probs[i].setup(device=dev)
    hs, vs = QuadsStaticTuning().make_geometry().get_design_from_rotated_square(QuadsStaticTuning().initial_angle)
    design_pinned = [hs.contiguous().pin_memory(), vs.contiguous().pin_memory()]
    grad_pinned = [torch.empty_like(d).pin_memory() for d in design_pinned]

    def task_vg(design, p, weight):
        d = [x.clone().requires_grad_(True) for x in design]
        J = weight * p.target_kinetic_energy(d, fused=True)
        J.backward()
        return J.detach(), [x.grad for x in d]

    timings = {}

    def step():
        design = [d.to(dev, non_blocking=True) for d in design_pinned]
        J, grads = multitask_value_and_grad(task_vg, design, probs, weights, timings)
        for o, g in zip(grad_pinned, grads):
            o.copy_(g, non_blocking=True)
        Jh = J.item()  # device -> host read of the objective (synchronises)
        return Jh

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    task_ms, ar_ms = {}, []
    with ClockSampler(local_rank) as clocks:
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(args.steps):
            J = step()
            torch.cuda.synchronize()
            for i, t0, t1 in timings["task_events"]:
                task_ms.setdefault(i, []).append(t0.elapsed_time(t1))
            ar_ms.append(timings["allreduce_events"][0].elapsed_time(timings["allreduce_events"][1]))
        stop.record()
        barrier()
        ms_total = start.elapsed_time(stop)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    per_task = torch.zeros(len(tasks), dtype=torch.float64, device=dev)
    for i, v in task_ms.items():
        per_task[i] = float(np.mean(v))
    # the collective's own duration is what the LAST rank to arrive sees (the others also wait for it): min over ranks
    ar = torch.tensor([float(np.mean(ar_ms))], dtype=torch.float64, device=dev)
    ar_wait = ar.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_task)
        dist.all_reduce(ar, op=dist.ReduceOp.MIN)
        dist.all_reduce(ar_wait, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_per_step = t.item() / args.steps
        pt = per_task.tolist()
        n_design = sum(d.numel() for d in design_pinned)
        val = len(tasks) / (ms_per_step * 1e-3)
        cfg = {"workload": "quads 24x18 static tuning (cfg4): 2 tasks (compressive strain 0.01 / 0.08, weights 0.75 / -0.25) of one "
                           "design, forward + adjoint per task, n_t=201, rtol=1e-8, atol=1e-4",
               "parallelism": f"tasks dealt to {world} rank(s); one all-reduce (sum, f64) of {1 + n_design} doubles per evaluation",
               "task_ms": pt, "allreduce_ms": ar.item(), "allreduce_ms_including_wait_for_the_slowest_rank": ar_wait.item(),
               "imbalance": f"task shares of the work {[round(x / sum(pt), 3) for x in pt]}: the 0.08-strain task bounds the 2-GPU time",
               "objective": J}
        if HORIZON_SCALE != 1.0:
            cfg["PROFILING_ONLY_horizon_scale"] = HORIZON_SCALE
        print(json.dumps({"metric": "forward+adjoint task evaluations per second", "value": val, "unit": "tasks/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                          "clocks": clocks.summary(),
                          "e2e": {"value": val, "unit": "tasks/s", "h2d_bytes_per_step": 8 * n_design * world,
                                  "d2h_bytes_per_step": 8 * (n_design + 1) * world,
                                  "note": "the timed step IS the public-API call with host buffers"},
                          "gpu_launches": args.steps * len(tasks) * 5, "roofline": None, "cpu_baseline": None}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--designs", type=int, default=1024, help="designs of the ensemble, sharded over the ranks (strong scaling)")
    ap.add_argument("--designs-per-gpu", type=int, default=0, help="weak scaling instead: this many designs on every rank")
    ap.add_argument("--cpu-designs", type=int, default=0,
                    help="designs per step of the CPU legs (default: two per host thread)")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4"],
                    help="cfg3: the 1024-design ensemble (the headline); cfg4: the two-task static-tuning objective with its NCCL all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--horizon-scale", type=float, default=1.0,
                    help="profiling only: shorten the simulated time and n_timepoints by this factor (invalid as a bench value)")
    args = ap.parse_args()
    global HORIZON_SCALE
    HORIZON_SCALE = args.horizon_scale
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "cfg4":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the CPU arm times the headline workload (cfg3) only"}))
            return
        return run_cfg4(args, rank, world, local_rank)

    weak = args.designs_per_gpu > 0
    total = args.designs_per_gpu * world if weak else args.designs
    from difflexmm_b200.parallel import shard_range
    lo, hi = shard_range(total, rank, world)
    config = {"workload": "quads_focusing 24x16 random-initial-guess ensemble (cfg3): forward + adjoint per design, "
                          "n_t=200, rtol=1e-8, atol=1e-4, contact on, noise 0.15*spacing",
              "designs_total": total, "designs_per_gpu": [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)],
              "parallelism": f"the {total}-design ensemble sharded over {world} rank(s) (contiguous slices), no collective",
              "l2": "inputs larger than L2 (trajectory ys = 3.5 MB per design, x designs_per_gpu)",
              "objective": "target kinetic energy evaluated on the device, cotangent formed inside the adjoint kernel",
              "launch_order": "longest-first, predicted by the step counts of the previous evaluation of the same batch (optimisation-loop "
                              "hint of DynamicSolver; the bench repeats identical designs, so the prediction is exact here; the first "
                              "evaluation, timed in roofline.first_evaluation_without_launch_order_history, has only the forward counts)"}
    if args.horizon_scale != 1.0:
        config["PROFILING_ONLY_horizon_scale"] = args.horizon_scale

    if args.impl == "reference":
        if rank != 0:
            return
        step, n, cores = run_cpu(args)
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
        val = n / dt
        sample = f"{n} designs of the same ensemble per step on {cores} host threads, C++ port of the reference algorithm ({step.build})"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                          "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from difflexmm_b200 import _abi
    from difflexmm_b200.dynamics import DynamicSolver

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = hi - lo
    prob, spec, drive, leaves_h, pb, dpd, aug, y0_h, ts_h = build_problem(B, seed0=lo)
    solver = DynamicSolver(spec, drive, prob.rtol, prob.atol, dev)
    lib = solver._lib
    nf = spec.n_free
    n_t = len(ts_h)
    tidx = torch.as_tensor(target_free_index(prob, spec), device=dev)
    tidx32 = tidx.to(torch.int32)
    ones_w = torch.ones(B, dtype=torch.float64, device=dev)

    leaves_pinned = {k: v.contiguous().pin_memory() for k, v in leaves_h.items()}
    leaves_d = {k: v.to(dev) for k, v in leaves_pinned.items()}
    y0, ts = y0_h.to(dev), ts_h.to(dev)
    ps = _abi.ParamSet(spec, B, leaves_d, pb, dpd)
    config["adjoint_kernel"] = lib.adjoint_plan(solver.handle, ps)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    kernel_ms = {"forward": [], "adjoint": []}
    last = {}

    def step(timed):
        if timed:
            ev[0].record()
        # the forward launch is ordered by the step counts of the previous step's solve of the same batch (as the solver
        # does between the iterations of an optimisation loop): a scheduling hint, the results do not depend on it
        fopt, forder = lib.longest_first(last["st_f"], solver.options, min_batch=297) if "st_f" in last else (solver.options, None)  # noqa: F841
        ys, st_f = lib.forward(solver.handle, ps, y0, ts, prob.rtol, prob.atol, fopt)
        if timed:
            ev[1].record()
        # objective on the device; the adjoint kernel forms its cotangent dJ/dys = m v itself (no g tensor)
        J, ibar, _ = lib.objective_value(solver.handle, ps, ys, tidx32)
        if timed:
            ev[2].record()
        # designs launched longest-first: shortens the tail when designs > SMs.  Predictor as in DynamicSolver
        # .adjoint_launch_options: the adjoint step counts of the previous evaluation of the same batch (the previous
        # iteration of an optimisation loop), else the forward step counts of this one (first evaluation)
        opt, order = lib.longest_first(last["st_b"] if "st_b" in last else st_f, solver.options)  # noqa: F841 (kept alive until the launch is queued)
        y0_bar, ts_bar, grads, st_b = lib.adjoint_objective(solver.handle, ps, ys, ts, tidx32, ones_w, prob.rtol, prob.atol,
                                                          aug, opt)
        if timed:
            ev[3].record()
        last.update(ys=ys, grads=grads, st_f=st_f, st_b=st_b, J=J)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    first_eval = None
    for w in range(args.warmup):
        step(w == 0)
        if w == 0:  # the first evaluation has no history: plain forward order, adjoint ordered by the forward step counts
            torch.cuda.synchronize()
            first_eval = {"forward_ms": ev[0].elapsed_time(ev[1]), "adjoint_ms": ev[2].elapsed_time(ev[3])}
    barrier()
    with ClockSampler(local_rank) as clocks:
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(args.steps):
            step(True)
            torch.cuda.synchronize()
            kernel_ms["forward"].append(ev[0].elapsed_time(ev[1]))
            kernel_ms["adjoint"].append(ev[2].elapsed_time(ev[3]))
        stop.record()
        barrier()
        ms_total = start.elapsed_time(stop)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    value = total / (ms_per_step * 1e-3)

    st_f, st_b = last["st_f"].numpy(), last["st_b"].numpy()
    bad = int((st_f["status"] != 0).sum() + (st_b["status"] != 0).sum())
    fwd_fl, bwd_fl = flops_model(spec, aug, n_t, 0, 0)
    flops_adj = bwd_fl(int(st_b["steps"].sum()), B)
    flops_fwd = fwd_fl(int(st_f["steps"].sum()), B)
    adj_ms = float(np.mean(kernel_ms["adjoint"]))
    fwd_ms = float(np.mean(kernel_ms["forward"]))

    # ---- e2e through the public API with host buffers ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # the call a user makes: value and design gradient of the problem's objective for a batch of designs held in
        # pinned host memory (design -> parameters, forward, objective, adjoint and geometry VJP all inside libdfx)
        prob._solver = solver
        design_pinned = [d.contiguous().pin_memory() for d in prob.ensemble]
        grad_pinned = [torch.empty_like(d).pin_memory() for d in design_pinned]
        obj_pinned = torch.empty(B, dtype=torch.float64).pin_memory()
        h2d = sum(d.numel() * 8 for d in design_pinned)
        d2h = sum(d.numel() * 8 for d in grad_pinned) + obj_pinned.numel() * 8

        def e2e_step():
            design = [d.to(dev, non_blocking=True).requires_grad_(True) for d in design_pinned]
            obj = prob.target_kinetic_energy(design, batch=B, fused=True)
            obj.sum().backward()
            for o, d in zip(grad_pinned, design):
                o.copy_(d.grad, non_blocking=True)
            obj_pinned.copy_(obj.detach(), non_blocking=True)
            # wait for the results by polling an event: on these hosts a blocking cudaDeviceSynchronize returned up to 450 ms
            # after the device had finished (device time per step is constant to 0.1 %, tools/e2e_jitter.py), more often with
            # NVML queries in flight on another thread
            done = torch.cuda.Event()
            done.record()
            while not done.query():
                pass

        for _ in range(max(1, args.warmup)):  # the first calls grow the allocator pools (multi-GB trajectory and workspace blocks)
            e2e_step()
        import gc
        gc.collect()  # autograd graphs of the warm-up steps released now, not by a collection inside the timed region
        gc.freeze()   # ... and the long-lived objects built so far (the ensemble, torch, numpy) leave the collector's
                      # working set: a full collection in the middle of a step stalled the launches by hundreds of ms
        barrier()
        n_e2e = args.steps
        e2e_step_ms = []
        with ClockSampler(local_rank) as e2e_clocks:
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                t1 = time.perf_counter()
                e2e_step()
                e2e_step_ms.append(1e3 * (time.perf_counter() - t1))
            barrier()
            te = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        bytes_t = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(bytes_t)
        e2e = {"value": total / te.item(), "unit": UNIT, "h2d_bytes_per_step": int(bytes_t[0].item()),
               "d2h_bytes_per_step": int(bytes_t[1].item()), "steps": n_e2e,
               "step_ms_median": float(np.median(e2e_step_ms)), "step_ms_max": float(np.max(e2e_step_ms)),
               "step_ms": [round(x, 1) for x in e2e_step_ms], "clocks": e2e_clocks.summary()}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline denominator: FP64 FMA peak measured here (MEASURED_PEAKS.json has no FP64 entry) -------
    peak, peak_src = None, "nominal 148 SMs x 64 DFMA/clk x 2 x sm_max_mhz"
    try:
        import ctypes as C
        lib.lib.dfx_fp64_peak.restype = C.c_double
        peak = float(lib.lib.dfx_fp64_peak(C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        peak_src = "measured in this run: register-resident DFMA chains, all SMs (dfx_fp64_peak)"
    except Exception:
        pass
    cl = clocks.summary()
    if not peak:
        peak = 148 * 64 * 2 * (cl["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
    achieved = flops_adj / (adj_ms * 1e-3) / 1e12
    # executed DFMA (x2) + DADD + DMUL per algorithmic flop of the same launch, from the committed ncu counters of this
    # build's adjoint kernel (tools/fp64_counts.py); None when the file is missing
    exe_ratio, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_fp64_instruction_counts.json")) as f:
            prof = json.load(f)
        exe_ratio = prof["executed_over_model"]
        traffic_note = prof.get("dram_note")
    except Exception:
        pass
    # DRAM bytes of one adjoint launch of exactly this workload (1024 designs on one GPU, full horizon) from the committed ncu
    # capture of this build; a counter cannot be read outside a profiler, so other configurations report null
    traffic = None
    try:
        if B == 1024 and world == 1 and HORIZON_SCALE == 1.0:
            with open(os.path.join(ROOT, "profiles", "r02_adjoint_traffic.json")) as f:
                tr = json.load(f)
            traffic = tr["dram_bytes"]
            traffic_note = "profiles/r02_adjoint_traffic.json (ncu capture of this build and workload): " + tr["note"]
    except Exception:
        pass
    roofline = {"kernel": config["adjoint_kernel"], "bound": "fp64_fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "frac_of_nominal_37.2": achieved / 37.2,
                "frac_executed": None if exe_ratio is None else achieved * exe_ratio / peak,
                # DRAM traffic is not measured in this run (it needs ncu); the capture of this build is cited instead
                "traffic": traffic, "traffic_unit": "DRAM bytes per launch" if traffic else None, "traffic_profile": traffic_note,
                "peak_source": peak_src,
                "note": "HBM and tensor rooflines do not bind this path (10.7 MB and 10 Gflop per design, no dense contraction)",
                "forward_kernel": {"achieved": flops_fwd / (fwd_ms * 1e-3) / 1e12, "ms": fwd_ms}, "adjoint_ms": adj_ms,
                "first_evaluation_without_launch_order_history": None if first_eval is None else dict(
                    first_eval, designs_per_s_this_rank=B / (1e-3 * (first_eval["forward_ms"] + first_eval["adjoint_ms"])),
                    frac=(flops_adj / (first_eval["adjoint_ms"] * 1e-3) / 1e12) / peak),
                "steps_fwd_mean": float(st_f["steps"].mean()), "steps_bwd_mean": float(st_b["steps"].mean())}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cstep, n, cores = run_cpu(args)
        t0 = time.perf_counter()
        cstep()
        dtc = time.perf_counter() - t0
        cpu_baseline = {"value": n / dtc, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n} designs of the same ensemble on {cores} host threads, C++ port of the reference "
                                  f"algorithm ({cstep.build}; not the JAX reference)"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
           "dtype": "f64",
           "data": "synthetic", "config": config, "clocks": cl, "e2e": e2e, "gpu_launches": 3 * args.steps,
           "roofline": roofline, "cpu_baseline": cpu_baseline, "failed_designs": bad}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
