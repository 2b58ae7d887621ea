"""Forward + adjoint timings of the five BASELINE.json configurations on one B200 (CUDA events), with the C++ oracle
port timed on one host core beside the single-design cases.  One JSON line per configuration.
Usage: python tools/config_timings.py [cfg1,cfg2,cfg4,cfg5]     (cfg3 is bench.py)"""
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def gpu_value_and_grad(P, design, reps=3):
    """latency of value_and_grad(target kinetic energy)(design) through the fused path; -> ms (median), J, stats"""
    times = []
    for r in range(reps + 1):
        d = [x.clone().cuda().requires_grad_(True) for x in design]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        J = P.target_kinetic_energy(d, fused=True)
        J.backward()
        e1.record()
        torch.cuda.synchronize()
        if r:
            times.append(e0.elapsed_time(e1))
    s = P.solver
    f, b = s.last_forward_stats.numpy()[0], s.last_adjoint_stats.numpy()[0]
    return float(np.median(times)), float(J.detach()), int(f["steps"]), int(b["steps"]), int(f["status"]) | int(b["status"])


def cpu_port(P, design):
    """same evaluation with the C++ oracle on one host thread -> seconds"""
    from oracle import Oracle
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(design, device="cpu")
    orc = Oracle(P.spec)
    ph = orc.params(1, {k: v.numpy() for k, v in leaves.items()}, pb, dpd)
    nf = P.spec.n_free
    t0 = time.perf_counter()
    ys, _ = orc.forward(ph, y0.numpy(), ts.numpy(), P.rtol, P.atol)
    g = np.zeros_like(ys)
    g[:, :, nf:] = ys[:, :, nf:] * leaves["inertia"].numpy()
    orc.adjoint(ph, ys, ts.numpy(), g, P.rtol, P.atol, aug)
    return time.perf_counter() - t0


def main():
    which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg1", "cfg2", "cfg4", "cfg5"]
    from difflexmm_b200.problems import KagomeFocusing, QuadsFocusing, QuadsStaticTuning
    if "cfg1" in which:
        P = QuadsFocusing()
        P.setup()
        design = P.initial_design()
        ms, J, sf, sb, st = gpu_value_and_grad(P, design)
        cpu = cpu_port(P, design)
        print(json.dumps({"config": "cfg1 quads_focusing 24x16, one design, value_and_grad", "gpu_ms": ms, "objective": J,
                          "fwd_steps": sf, "bwd_steps": sb, "status": st, "cpu_port_s_1core": cpu,
                          "latency_ratio": cpu * 1e3 / ms}), flush=True)
    if "cfg2" in which:
        P = KagomeFocusing()
        P.setup()
        design = P.initial_design()
        ms, J, sf, sb, st = gpu_value_and_grad(P, design)
        cpu = cpu_port(P, design)
        print(json.dumps({"config": "cfg2 kagome_focusing 20x12, one design, value_and_grad", "gpu_ms": ms, "objective": J,
                          "fwd_steps": sf, "bwd_steps": sb, "status": st, "cpu_port_s_1core": cpu,
                          "latency_ratio": cpu * 1e3 / ms}), flush=True)
    if "cfg4" in which:
        from difflexmm_b200.parallel import multitask_value_and_grad
        tasks, weights = [dict(compressive_strain=0.01), dict(compressive_strain=0.08)], [0.75, -0.25]
        probs = [QuadsStaticTuning(**t) for t in tasks]
        for p in probs:
            p.setup()
        hs, vs = probs[0].initial_design()

        def task_vg(design, p, weight):
            d = [x.clone().requires_grad_(True) for x in design]
            J = weight * p.target_kinetic_energy(d, fused=True)
            J.backward()
            return J.detach(), [x.grad for x in d]

        design = [hs.cuda(), vs.cuda()]
        times = []
        for r in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            J, grads = multitask_value_and_grad(task_vg, design, probs, weights)
            torch.cuda.synchronize()
            if r:
                times.append(time.perf_counter() - t0)
        steps = [(int(p.solver.last_forward_stats.numpy()[0]["steps"]), int(p.solver.last_adjoint_stats.numpy()[0]["steps"])) for p in probs]
        print(json.dumps({"config": "cfg4 static tuning 24x18, 2 tasks (strains 0.01 / 0.08, weights 0.75 / -0.25) on ONE GPU, summed gradient",
                          "gpu_ms": 1e3 * float(np.median(times)), "objective": float(J), "steps_fwd_bwd_per_task": steps,
                          "allreduce_doubles": int(sum(g.numel() for g in grads)) + 1}), flush=True)
    if "cfg5" in which:
        periods = 2.0
        P = QuadsFocusing(n1_blocks=100, n2_blocks=100, simulation_time=periods / 30.0, n_timepoints=16, target_shift=(2, 2),
                          min_angle=15 * math.pi / 180, cutoff_angle=45 * math.pi / 180)
        P.setup()
        design = P.initial_design()
        ms, J, sf, sb, st = gpu_value_and_grad(P, design, reps=1)
        print(json.dumps({"config": f"cfg5 quads 100x100 contact active, {periods} drive periods, one lattice over all SMs (cooperative group, software barrier), value_and_grad",
                          "gpu_ms": ms, "objective": J, "fwd_steps": sf, "bwd_steps": sb, "status": st,
                          "us_per_step_fwd_plus_bwd": 1e3 * ms / max(1, sf + sb)}), flush=True)


if __name__ == "__main__":
    main()
