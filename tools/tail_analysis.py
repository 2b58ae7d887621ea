#!/usr/bin/env python
"""How much of the adjoint launch of the bench workload is tail (one CTA per SM, designs of unequal length)?
Runs forward + adjoint once for the cfg3 ensemble, then replays the hardware's list scheduling (next CTA to the first free
SM, in launch order) on the measured adjoint step counts for three launch orders: plain, longest-first by the FORWARD step
counts (what the library does), longest-first by the adjoint's own counts (a perfect predictor).  Prints the makespans
relative to the ideal sum / n_sm.   python tools/tail_analysis.py [--designs 1024] [--sms 148]"""
import argparse
import heapq
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def makespan(cost, order, n_sm):
    free = [0.0] * n_sm
    heapq.heapify(free)
    end = 0.0
    for i in order:
        t = heapq.heappop(free) + cost[i]
        end = max(end, t)
        heapq.heappush(free, t)
    return end


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--designs", type=int, default=1024)
    ap.add_argument("--sms", type=int, default=148)
    args = ap.parse_args()
    from difflexmm_b200 import _abi
    from difflexmm_b200.dynamics import DynamicSolver
    dev = torch.device("cuda", 0)
    B = args.designs
    prob, spec, drive, leaves_h, pb, dpd, aug, y0_h, ts_h = bench.build_problem(B, seed0=0)
    solver = DynamicSolver(spec, drive, prob.rtol, prob.atol, dev)
    lib = solver._lib
    tidx32 = torch.as_tensor(bench.target_free_index(prob, spec), device=dev).to(torch.int32)
    ps = _abi.ParamSet(spec, B, {k: v.to(dev) for k, v in leaves_h.items()}, pb, dpd)
    y0, ts = y0_h.to(dev), ts_h.to(dev)
    ys, st_f = lib.forward(solver.handle, ps, y0, ts, prob.rtol, prob.atol, solver.options)
    opt, order = lib.longest_first(st_f, solver.options)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        e0.record()
        _, _, _, st_b = lib.adjoint_objective(solver.handle, ps, ys, ts, tidx32, torch.ones(B, dtype=torch.float64, device=dev),
                                              prob.rtol, prob.atol, aug, opt)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    sf, sb = st_f.numpy()["steps"].astype(np.float64), st_b.numpy()["steps"].astype(np.float64)
    ideal = sb.sum() / args.sms
    out = {"designs": B, "sms": args.sms, "adjoint_ms": ms, "steps_bwd_mean": sb.mean(), "steps_bwd_max": sb.max(), "steps_bwd_min": sb.min(),
           "corr_fwd_bwd_steps": float(np.corrcoef(sf, sb)[0, 1]),
           "makespan_over_ideal": {"plain order": makespan(sb, range(B), args.sms) / ideal,
                                   "longest first by forward steps": makespan(sb, np.argsort(-sf, kind="stable"), args.sms) / ideal,
                                   "longest first by adjoint steps": makespan(sb, np.argsort(-sb, kind="stable"), args.sms) / ideal},
           "ms_per_step_if_ideal": ms / (makespan(sb, np.argsort(-sf, kind="stable"), args.sms))}
    # is there a better predictor of the adjoint's length inside the forward solve's statistics?
    fs = st_f.numpy()
    feats = {"steps": sf, "accepted": fs["accepted"].astype(np.float64), "rejected": sf - fs["accepted"].astype(np.float64),
             "1/last_dt": 1.0 / np.maximum(fs["last_dt"], 1e-300)}
    ysn = ys.detach()
    nf_ = spec.n_free
    feats["max |v|"] = ysn[:, :, nf_:].abs().amax(dim=(1, 2)).cpu().numpy()
    feats["max |u|"] = ysn[:, :, :nf_].abs().amax(dim=(1, 2)).cpu().numpy()
    feats["objective"] = (0.5 * ysn[:, :, nf_:] ** 2).sum(dim=(1, 2)).cpu().numpy()
    out["corr_of_adjoint_steps_with_forward_features"] = {k: float(np.corrcoef(v, sb)[0, 1]) for k, v in feats.items()}
    out["makespan_over_ideal_ordered_by_feature"] = {k: makespan(sb, np.argsort(-v, kind="stable"), args.sms) / ideal for k, v in feats.items()}
    # a short adjoint pre-pass over the last k output intervals as the predictor of the full adjoint's length?
    out["prepass_over_last_intervals"] = {}
    for k in (3, 6, 11, 21):
        ys_k, ts_k = ys[:, -k:].contiguous(), ts[-k:].contiguous()
        e0.record()
        _, _, _, st_k = lib.adjoint_objective(solver.handle, ps, ys_k, ts_k, tidx32, torch.ones(B, dtype=torch.float64, device=dev),
                                              prob.rtol, prob.atol, aug, solver.options)
        e1.record()
        torch.cuda.synchronize()
        sk = st_k.numpy()["steps"].astype(np.float64)
        out["prepass_over_last_intervals"][f"{k - 1} intervals"] = {
            "ms": e0.elapsed_time(e1), "corr_with_full_adjoint_steps": float(np.corrcoef(sk, sb)[0, 1]),
            "makespan_over_ideal": makespan(sb, np.argsort(-sk, kind="stable"), args.sms) / ideal}
    out["forward"] = {"steps_mean": sf.mean(), "steps_max": sf.max(),
                      "makespan_over_ideal_296_slots": {"plain order": makespan(sf, range(B), 2 * args.sms) / (sf.sum() / (2 * args.sms)),
                                                        "longest first": makespan(sf, np.argsort(-sf, kind="stable"), 2 * args.sms) / (sf.sum() / (2 * args.sms))}}
    # an optimisation loop moves the designs a little between two evaluations: how well do the PREVIOUS adjoint step counts
    # order the next launch?  Vertices moved by a random 0.1 % / 1 % of the lattice spacing.
    out["previous_evaluation_as_predictor"] = {}
    cnv0 = leaves_h["centroid_node_vectors"]
    for rel in (0.001, 0.01):
        g = torch.Generator().manual_seed(5)
        lv = dict(leaves_h)
        lv["centroid_node_vectors"] = cnv0 + rel * prob.spacing * torch.randn(cnv0.shape, generator=g, dtype=cnv0.dtype)
        ps1 = _abi.ParamSet(spec, B, {k: v.to(dev) for k, v in lv.items()}, pb, dpd)
        ys1, st_f1 = lib.forward(solver.handle, ps1, y0, ts, prob.rtol, prob.atol, solver.options)
        _, _, _, st_b1 = lib.adjoint_objective(solver.handle, ps1, ys1, ts, tidx32, torch.ones(B, dtype=torch.float64, device=dev),
                                               prob.rtol, prob.atol, aug, opt)
        sb1 = st_b1.numpy()["steps"].astype(np.float64)
        sf1 = st_f1.numpy()["steps"].astype(np.float64)
        ideal1 = sb1.sum() / args.sms
        out["previous_evaluation_as_predictor"][f"vertices moved by {rel:g} x spacing"] = {
            "corr_adjoint_steps_prev_next": float(np.corrcoef(sb, sb1)[0, 1]),
            "makespan_over_ideal": {"ordered by the previous adjoint counts": makespan(sb1, np.argsort(-sb, kind="stable"), args.sms) / ideal1,
                                    "ordered by this forward's counts": makespan(sb1, np.argsort(-sf1, kind="stable"), args.sms) / ideal1,
                                    "perfect": makespan(sb1, np.argsort(-sb1, kind="stable"), args.sms) / ideal1}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
