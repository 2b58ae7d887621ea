"""Multi-GPU layer of the path: one process per GPU, the *design batch* is the
parallel dimension.

  * ensembles / sweeps: every rank integrates its contiguous slice of the designs; there is no
    data-path collective (results are gathered by the caller if it wants them in one place);
  * summed multi-task objectives (the reference's `pmap(..., in_axes=(None, 0, 0))` over tasks,
    `problems/quads_kinetic_energy_static_tuning.py:473-478`, whose VJP sums the per-task design
    cotangents): tasks are dealt to ranks, weights are applied locally, then ONE all-reduce (sum, f64) of
    the design-gradient vector and the objective per optimiser iteration -- NCCL over NVLink on GPUs,
    gloo in the CPU tests.
"""

from typing import Callable, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [begin, end) slice of `n_units` independent units owned by `rank` (sizes differ by <= 1)"""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """in-place sum over ranks (no-op without a process group)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def multitask_value_and_grad(task_value_and_grad: Callable, design: Sequence[torch.Tensor], tasks: Sequence,
                             weights: Sequence[float], timings: dict = None):
    """weights @ [objective(design, task) for task in tasks] and its gradient w.r.t. the shared design.

    `task_value_and_grad(design, task, weight) -> (weight*value, [grad of weight*value per design tensor])`
    evaluates one task (one forward + adjoint solve).  The weight goes INTO the differentiated function, as in the
    reference where the cotangent entering each task's odeint backward is `weight_i * d objective_i / d ys`: the
    adjoint solve is adaptive with an absolute tolerance, so scaling its cotangent afterwards is not the same
    computation (the step sequence differs; results agree only to the integration tolerance).
    Tasks are sharded over the ranks; the partial sums are packed into one flat f64 buffer and all-reduced once.

    `timings` (CUDA devices only): a dict that receives `task_events` = [(task index, start, stop)] and
    `allreduce_events` = (start, stop), torch.cuda.Event pairs recorded on the current stream around every task of this
    rank and around the collective (read them with `start.elapsed_time(stop)` after a synchronisation)."""
    rank, nranks = world()
    b, e = shard_range(len(tasks), rank, nranks)
    dev = design[0].device
    sizes = [d.numel() for d in design]
    buf = torch.zeros(1 + sum(sizes), dtype=torch.float64, device=dev)
    timed = timings is not None and dev.type == "cuda"

    def mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    if timed:
        timings["task_events"] = []
    for i in range(b, e):
        t0 = mark() if timed else None
        v, gs = task_value_and_grad(design, tasks[i], weights[i])
        buf[0] += v
        off = 1
        for g, n in zip(gs, sizes):
            buf[off:off + n] += g.reshape(-1).to(torch.float64)
            off += n
        if timed:
            timings["task_events"].append((i, t0, mark()))
    t0 = mark() if timed else None
    allreduce_sum_(buf)
    if timed:
        timings["allreduce_events"] = (t0, mark())
    grads, off = [], 1
    for d, n in zip(design, sizes):
        grads.append(buf[off:off + n].reshape(d.shape))
        off += n
    return buf[0], grads
