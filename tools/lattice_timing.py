"""Time forward + adjoint of ONE large lattice (cfg5 of BASELINE.json: quads n x n with contact) for several
cluster sizes of the generic kernels (DFX_CLUSTER).  Usage: python tools/lattice_timing.py [n] [sim_periods] [clusters] [batch]
(batch > 1: that many copies of the lattice in one launch -- throughput of a small batch of large lattices; "d" = library default)
Prints one JSON line per cluster size (CUDA-event times, steps, microseconds per RHS evaluation)."""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    periods = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
    # "16" = thread-block cluster of 16 CTAs (DFX_CLUSTER), "g64" = cooperative group of 64 CTAs (DFX_GROUP)
    clusters = sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "2", "4", "8", "16"]
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    from difflexmm_b200 import _abi
    from difflexmm_b200.problems import QuadsFocusing
    P = QuadsFocusing(n1_blocks=n, n2_blocks=n, simulation_time=periods / 30.0, n_timepoints=8, target_shift=(2, 2),
                      min_angle=15 * math.pi / 180, cutoff_angle=45 * math.pi / 180)
    s = P.setup()
    leaves, pb, dpd, aug, y0, ts = P.boundary_inputs(P.initial_design(), device="cuda")
    ps = _abi.ParamSet(P.spec, batch, {k: v.contiguous() for k, v in leaves.items()}, pb, dpd)
    nf = P.spec.n_free
    ref = None
    for cl in clusters:
        os.environ.pop("DFX_GROUP", None)
        os.environ.pop("DFX_CLUSTER", None)
        if str(cl).startswith("g"):
            os.environ["DFX_GROUP"] = str(cl)[1:]
        elif str(cl) != "d":
            os.environ["DFX_CLUSTER"] = str(cl)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ys_w, _ = s.lib_forward(ps, y0, ts)  # warm-up of both kernels (lazy module loading, allocator pools)
        s.lib_adjoint(ps, ys_w, ts, torch.zeros_like(ys_w), aug)
        torch.cuda.synchronize()
        ev[0].record()
        ys, st = s.lib_forward(ps, y0, ts)
        ev[1].record()
        g = torch.zeros_like(ys)
        g[:, :, nf:] = ys[:, :, nf:] * leaves["inertia"]
        torch.cuda.synchronize()
        ev[1].record()
        y0b, tsb, gr, sb = s.lib_adjoint(ps, ys, ts, g, aug)
        ev[2].record()
        torch.cuda.synchronize()
        f, b = st.numpy()[0], sb.numpy()[0]
        out = {"lattice": f"{n}x{n}", "batch": batch, "cluster": cl, "forward_ms": ev[0].elapsed_time(ev[1]) if False else None,
               "fwd_steps": int(f["steps"]), "bwd_steps": int(b["steps"]), "status": [int(f["status"]), int(b["status"])]}
        # separate timing of the forward (ev[1] was re-recorded above)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.lib_forward(ps, y0, ts); e1.record(); torch.cuda.synchronize()
        out["forward_ms"] = e0.elapsed_time(e1)
        out["adjoint_ms"] = ev[1].elapsed_time(ev[2])
        out["lattices_per_s"] = batch / (1e-3 * (out["forward_ms"] + out["adjoint_ms"]))
        out["us_per_rhs_fwd"] = 1e3 * out["forward_ms"] / max(1, int(f["rhs_evals"]))
        out["us_per_aug_rhs"] = 1e3 * out["adjoint_ms"] / max(1, int(b["rhs_evals"]))
        cur = (ys.cpu().numpy(), gr["centroid_node_vectors"].cpu().numpy())
        if ref is None:
            ref = cur
        out["traj_vs_first"] = float(np.linalg.norm(cur[0] - ref[0]) / np.linalg.norm(ref[0]))
        out["grad_vs_first"] = float(np.linalg.norm(cur[1] - ref[1]) / max(np.linalg.norm(ref[1]), 1e-300))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
