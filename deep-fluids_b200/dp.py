"""Data-parallel host logic (new; the reference is single-GPU): one process per GPU, batch sharded by rank, ONE
all-reduce of the flat fp32 gradient buffer per step, 1/world folded into the optimizer's grad_scale.  Backend-agnostic
(NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if world() > 1 else 0


def allreduce_grads_(flat_grad):
    """Sum the flat gradient buffer over ranks in place; returns the grad_scale (1/world) for the optimizer."""
    w = world()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / w


def max_over_ranks(value, device=None):
    """Timing aggregation for bench.py: a step is as slow as its slowest rank."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_seed(seed, r=None):
    """Data seed per rank (weights use the un-shifted seed so replicas start identical)."""
    return int(seed) + (rank() if r is None else int(r))
