"""Data-parallel host logic (new; the reference is single-GPU): one process per GPU, batch sharded by rank, ONE
all-reduce of the flat fp32 gradient buffer per step, 1/world folded into the optimizer's grad_scale.  Backend-agnostic
(NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if world() > 1 else 0


_NATIVE = {"comm": None}


def native_comm():
    """NCCL communicator owned by the C-ABI library (dfl_comm_init), created on first use: rank 0 draws the unique id,
    torch.distributed only carries those 128 bytes to the other ranks.  Selected with DFL_NATIVE_NCCL=1; the all-reduce
    then runs through dfl_allreduce on the step's own stream (and can be captured into the step's CUDA graph)."""
    import ctypes as C
    from . import cabi
    if _NATIVE["comm"] is None:
        lib = cabi.lib()
        idbuf = (C.c_char * 128)()
        if rank() == 0:
            cabi.check(lib.dfl_comm_unique_id(C.cast(idbuf, C.c_void_p)))
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device())
                         if dist.get_backend() == "nccl" else None)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        comm = C.c_void_p()
        cabi.check(lib.dfl_comm_init(C.byref(comm), world(), C.c_char_p(raw), rank()))
        _NATIVE["comm"] = comm
    return _NATIVE["comm"]


def use_native():
    import os
    return os.environ.get("DFL_NATIVE_NCCL", "0") == "1" and world() > 1 and dist.get_backend() == "nccl"


def allreduce_grads_(flat_grad):
    """Sum the flat gradient buffer over ranks in place; returns the grad_scale (1/world) for the optimizer."""
    w = world()
    if w > 1:
        from .kernels import nvtx_range
        with nvtx_range("gradient all-reduce"):
            _allreduce(flat_grad)
    return 1.0 / w


def _allreduce(flat_grad):
    if use_native():
        import ctypes as C
        from . import cabi
        cabi.check(cabi.lib().dfl_allreduce(C.c_void_p(flat_grad.data_ptr()), flat_grad.numel(), cabi.F32, native_comm(),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    else:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)


def max_over_ranks(value, device=None):
    """Timing aggregation for bench.py: a step is as slow as its slowest rank."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_seed(seed, r=None):
    """Data seed per rank (weights use the un-shifted seed so replicas start identical)."""
    return int(seed) + (rank() if r is None else int(r))
