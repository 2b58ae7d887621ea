#!/usr/bin/env python
"""A/B timing of adjoint kernel variants on the bench workload (cfg3 ensemble): one forward solve, then the adjoint of
every variant on the same trajectories; prints one JSON line per variant with the CUDA-event time of the launch and
the largest deviation of its gradients from the first variant.

  python tools/adjoint_ab.py --designs 296 --variants v2,v3 [--horizon-scale 0.25]

A variant is a value of DFX_ADJOINT_KERNEL ("v3" = unset = the default choice of the library)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--designs", type=int, default=296)
    ap.add_argument("--variants", default="v2,v3")
    ap.add_argument("--horizon-scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    bench.HORIZON_SCALE = args.horizon_scale
    from difflexmm_b200 import _abi
    from difflexmm_b200.dynamics import DynamicSolver
    dev = torch.device("cuda", 0)
    B = args.designs
    prob, spec, drive, leaves_h, pb, dpd, aug, y0_h, ts_h = bench.build_problem(B, seed0=0)
    solver = DynamicSolver(spec, drive, prob.rtol, prob.atol, dev)
    lib = solver._lib
    tidx32 = torch.as_tensor(bench.target_free_index(prob, spec), device=dev).to(torch.int32)
    ones_w = torch.ones(B, dtype=torch.float64, device=dev)
    leaves_d = {k: v.to(dev) for k, v in leaves_h.items()}
    y0, ts = y0_h.to(dev), ts_h.to(dev)
    ps = _abi.ParamSet(spec, B, leaves_d, pb, dpd)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ys, st_f = lib.forward(solver.handle, ps, y0, ts, prob.rtol, prob.atol, solver.options)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"forward_ms": e0.elapsed_time(e1), "designs": B, "steps_fwd_mean": float(st_f.numpy()["steps"].mean())}), flush=True)
    ref = None
    for var in args.variants.split(","):
        if var == "v3":
            os.environ.pop("DFX_ADJOINT_KERNEL", None)
        else:
            os.environ["DFX_ADJOINT_KERNEL"] = var
        times = []
        for _ in range(args.reps + 1):
            e0.record()
            y0_bar, ts_bar, grads, st_b = lib.adjoint_objective(solver.handle, ps, ys, ts, tidx32, ones_w, prob.rtol, prob.atol,
                                                              aug, solver.options)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        sb = st_b.numpy()
        out = {"variant": var, "adjoint_ms": min(times[1:]), "all_ms": times, "steps_bwd_mean": float(sb["steps"].mean()),
               "bad": int((sb["status"] != 0).sum())}
        cur = {k: v.double().cpu().numpy() for k, v in grads.items() if v is not None}
        cur["y0_bar"] = y0_bar.cpu().numpy(); cur["ts_bar"] = ts_bar.cpu().numpy(); cur["steps"] = sb["steps"].astype(np.float64)
        if ref is None:
            ref = cur
        else:
            dev_ = {}
            for k in ref:
                a, b = ref[k].reshape(B, -1), cur[k].reshape(B, -1)
                num = np.linalg.norm(a - b, axis=1)
                den = np.maximum(np.linalg.norm(a, axis=1), 1e-300)
                dev_[k] = float(np.nanmax(num / den)) if np.isfinite(num).all() else float("nan")
            out["max_rel_l2_vs_first"] = dev_
            out["steps_equal"] = bool((ref["steps"] == cur["steps"]).all())
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
