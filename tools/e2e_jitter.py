"""Step-to-step jitter of the user-level call (diagnostic): host segments and device time of
problem.target_kinetic_energy(..., fused=True) + backward for the cfg3 ensemble, many steps.
  python tools/e2e_jitter.py [designs] [steps] [--sampler]   (--sampler: run bench.py's NVML clock sampler beside it)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(args[0]) if len(args) > 0 else 1024
    n = int(args[1]) if len(args) > 1 else 12
    from difflexmm_b200.problems import QuadsFocusing
    P = QuadsFocusing()
    P.setup()
    hs, vs = P.random_ensemble(B, noise=0.15)
    pinned = [hs.contiguous().pin_memory(), vs.contiguous().pin_memory()]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step():
        t0 = time.perf_counter()
        e[0].record()
        d = [x.to("cuda", non_blocking=True).requires_grad_(True) for x in pinned]
        J = P.target_kinetic_energy(d, batch=B, fused=True)
        e[1].record()
        t1 = time.perf_counter()
        J.sum().backward()
        e[2].record()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        return {"host_fwd_call_ms": 1e3 * (t1 - t0), "host_bwd_call_ms": 1e3 * (t2 - t1), "host_sync_ms": 1e3 * (t3 - t2),
                "total_ms": 1e3 * (t3 - t0), "dev_fwd_ms": e[0].elapsed_time(e[1]), "dev_bwd_ms": e[1].elapsed_time(e[2]),
                "alloc_retries": torch.cuda.memory_stats().get("num_alloc_retries", 0),
                "reserved_gb": torch.cuda.memory_reserved() / 2**30, "segments": torch.cuda.memory_stats().get("segment.all.current", 0)}

    sampler = None
    if "--sampler" in sys.argv:
        import bench
        sampler = bench.ClockSampler(0)
        sampler.__enter__()
    for _ in range(3):
        step()
    for i in range(n):
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in step().items()}), flush=True)
    if sampler is not None:
        sampler.__exit__(None, None, None)


if __name__ == "__main__":
    main()
