"""Host logic of the multi-GPU layer on CPU: world_size-2 gloo process groups."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from difflexmm_b200.parallel import multitask_value_and_grad, shard_range


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _task_value_and_grad(design, task, weight=1.0):
    # stand-in for one forward + adjoint solve: a smooth function of the shared design and the task's parameters
    hs, vs = design
    amp, rate = task
    v = amp * (hs ** 2).sum() + rate * torch.sin(vs).sum()
    return weight * v, [weight * 2 * amp * hs, weight * rate * torch.cos(vs)]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    design = [torch.randn(5, 4, 2, dtype=torch.float64), torch.randn(4, 5, 2, dtype=torch.float64)]
    tasks = [(1.0, 30.0), (0.5, 20.0), (2.0, 10.0)]
    weights = [0.75, -0.25, 0.5]
    v, g = multitask_value_and_grad(_task_value_and_grad, design, tasks, weights)
    if rank == 0:
        torch.save({"v": v, "g": g}, out)
    dist.destroy_process_group()


def test_multitask_gradient_allreduce_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    design = [torch.randn(5, 4, 2, dtype=torch.float64), torch.randn(4, 5, 2, dtype=torch.float64)]
    tasks = [(1.0, 30.0), (0.5, 20.0), (2.0, 10.0)]
    weights = [0.75, -0.25, 0.5]
    v_ref = sum(w * _task_value_and_grad(design, t)[0] for w, t in zip(weights, tasks))
    g_ref = [sum(w * _task_value_and_grad(design, t)[1][k] for w, t in zip(weights, tasks)) for k in range(2)]
    assert torch.allclose(got["v"], v_ref, rtol=1e-14)
    for a, b in zip(got["g"], g_ref):
        assert torch.allclose(a, b, rtol=1e-14, atol=1e-14)
