#!/usr/bin/env python
"""Constrained batched MMA on the cfg1 lattice (quads 24 x 16): B random initial guesses optimised in lock-step for the target
kinetic energy under the reference's angle / edge-length constraints (run_optimization_nlopt's switches,
problems/quads_focusing.py:546-652).  The thresholds are set halfway between the starting designs and what the unconstrained
run reaches, so the constraints are active.  Prints / writes one JSON record: objective and largest constraint violation per
evaluation for the unconstrained and the constrained run, seconds per iteration.
  python tools/constrained_mma_demo.py [B] [iterations]  ->  gpurun_out/r02_constrained_mma_cfg1.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    from difflexmm_b200.optimization import OptimizationProblem
    from difflexmm_b200.problems import QuadsFocusing
    P = QuadsFocusing()
    P.setup(max_steps=30000)  # a design the optimiser breaks must not integrate for minutes
    guesses = [g.cuda() for g in P.random_ensemble(B, noise=0.05)]
    opt0 = OptimizationProblem(P)
    x0 = opt0.flatten(guesses).cuda()
    # asymptotes 2 % of the bound span away from the start: candidates stay near valid lattices (spacing 15, hinges 2.25 long)
    bounds = dict(lower_bound=float(x0.min()) - 0.5, upper_bound=float(x0.max()) + 0.5, initial_move=0.02)
    dc0 = opt0.device_constraints(0.0, 0.0, 0.0)

    def minima(x):
        c, _ = dc0(x)
        na = dc0.n_angle_rows
        return [float((-c[:, :na // 2]).min()), float((-c[:, na // 2:na]).min()), float((-c[:, na:]).min())]

    t0 = time.time()
    free = OptimizationProblem(P)
    best_free, f_free = free.run_optimization_mma(guesses, n_iterations=iters, **bounds)
    t_free = time.time() - t0
    start, reached = minima(x0), minima(free.flatten(best_free).cuda())
    thr = [0.5 * (s + r) if r < s else 0.97 * s for s, r in zip(start, reached)]
    con = OptimizationProblem(P)
    t0 = time.time()
    best_con, f_con = con.run_optimization_mma(guesses, n_iterations=iters, **bounds, min_void_angle=thr[0], min_block_angle=thr[1],
                                               min_edge_length=thr[2])
    t_con = time.time() - t0
    dc = con.device_constraints(thr[0], thr[1], thr[2])
    out = {
        "lattice": "quads 24x16 (cfg1)", "instances": B, "evaluations_per_instance": iters,
        "constraint_rows_per_instance": dc.n_rows, "design_variables_per_instance": int(x0.shape[1]),
        "smallest void angle / block angle / edge length": {"start": start, "unconstrained optimum": reached, "thresholds": thr},
        "unconstrained": {"objective_mean_per_evaluation": [float(h.mean()) for h in free.objective_values],
                          "best_objective_mean": float(f_free.mean()),
                          "violation_of_the_thresholds_at_its_optimum": float(dc(free.flatten(best_free).cuda())[0].max()),
                          "seconds_per_evaluation": t_free / iters},
        "constrained": {"objective_mean_per_evaluation": [float(h.mean()) for h in con.objective_values],
                        "largest_violation_per_evaluation": [float(v.max()) for v in con.optimizer.violation_history],
                        "best_objective_mean": float(f_con.mean()), "largest_violation_of_the_best_designs": float(con.optimizer.best_violation.max()),
                        "instances_improved": int((f_con > torch.as_tensor(con.objective_values[0], device=f_con.device) * 1.001).sum()),
                        "seconds_per_evaluation": t_con / iters},
    }
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_constrained_mma_cfg1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
