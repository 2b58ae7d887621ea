#!/usr/bin/env python
"""Repeated timing of the forward launch on the bench workload (CUDA events, best of N after a warm-up launch).
  python tools/forward_ab.py --designs 296 [--reps 4]     (DFX_LIB selects another build of the library)"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--designs", type=int, default=296)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    from difflexmm_b200 import _abi
    from difflexmm_b200.dynamics import DynamicSolver
    dev = torch.device("cuda", 0)
    B = args.designs
    prob, spec, drive, leaves_h, pb, dpd, aug, y0_h, ts_h = bench.build_problem(B, seed0=0)
    solver = DynamicSolver(spec, drive, prob.rtol, prob.atol, dev)
    ps = _abi.ParamSet(spec, B, {k: v.to(dev) for k, v in leaves_h.items()}, pb, dpd)
    y0, ts = y0_h.to(dev), ts_h.to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(args.reps + 1):
        e0.record()
        ys, st = solver._lib.forward(solver.handle, ps, y0, ts, prob.rtol, prob.atol, solver.options)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    s = st.numpy()
    print(json.dumps({"forward_ms": min(times[1:]), "all_ms": times, "designs": B, "steps_mean": float(s["steps"].mean()),
                      "bad": int((s["status"] != 0).sum()), "checksum": float(ys.double().abs().sum().item())}))


if __name__ == "__main__":
    main()
