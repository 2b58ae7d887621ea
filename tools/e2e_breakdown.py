"""Where the host-side time of the user-level call goes (diagnostic): cProfile of problem.target_kinetic_energy(...,
fused=True) + backward for the cfg3 ensemble."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    from difflexmm_b200.problems import QuadsFocusing
    P = QuadsFocusing()
    P.setup()
    hs, vs = P.random_ensemble(B, noise=0.15)
    pinned = [hs.contiguous().pin_memory(), vs.contiguous().pin_memory()]

    def step():
        d = [x.to("cuda", non_blocking=True).requires_grad_(True) for x in pinned]
        t0 = time.perf_counter()
        J = P.target_kinetic_energy(d, batch=B, fused=True)
        t1 = time.perf_counter()
        J.sum().backward()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        return t1 - t0, t2 - t1, t3 - t2

    step()
    print("host seconds (forward call, backward call, final sync):", step())
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
