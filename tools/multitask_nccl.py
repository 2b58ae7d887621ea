"""cfg4 of BASELINE.json under torchrun: the tasks of the static-tuning objective dealt to the ranks (one GPU each),
weights applied locally, ONE NCCL all-reduce of [objective | design gradient] per evaluation.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multitask_nccl.py
Rank 0 prints one JSON line (time = max over ranks, CUDA-synchronised wall clock around the whole evaluation)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from difflexmm_b200.parallel import multitask_value_and_grad, shard_range
    from difflexmm_b200.problems import QuadsStaticTuning
    tasks, weights = [dict(compressive_strain=0.01), dict(compressive_strain=0.08)], [0.75, -0.25]
    b, e = shard_range(len(tasks), rank, world)
    probs = [QuadsStaticTuning(**t) for t in tasks]
    for i in range(b, e):  # a rank only sets up the solvers of its own tasks
        probs[i].setup(device=dev)
    hs, vs = QuadsStaticTuning().make_geometry().get_design_from_rotated_square(QuadsStaticTuning().initial_angle)

    def task_vg(design, p, weight):
        d = [x.clone().requires_grad_(True) for x in design]
        J = weight * p.target_kinetic_energy(d, fused=True)
        J.backward()
        return J.detach(), [x.grad for x in d]

    design = [hs.to(dev), vs.to(dev)]
    times = []
    for r in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        J, grads = multitask_value_and_grad(task_vg, design, probs, weights)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if r:
            times.append(t.item())
    if rank == 0:
        print(json.dumps({"config": "cfg4 static tuning 24x18, 2 tasks over %d GPU(s), one all-reduce of %d doubles" % (world, 1 + sum(g.numel() for g in grads)),
                          "n_gpus": world, "seconds_per_evaluation": float(np.median(times)), "objective": float(J),
                          "grad_l2": float(torch.sqrt(sum((g ** 2).sum() for g in grads)))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
