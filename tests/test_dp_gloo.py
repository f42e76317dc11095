"""CPU, world_size=2, gloo: the data-parallel host logic (gradient all-reduce + 1/world scaling + identical optimizer
step on every rank, max-over-ranks timing) gives the same parameters as one process on the concatenated batch."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepfluids_b200 import dp
    from oracle import ref_model as M, ref_train as T
    torch.set_num_threads(2)
    spatial = [16, 8]
    tab, _, _ = M.generator_layout(spatial + [1], filters=8, num_conv=1)
    var = M.init_variables(tab, 123)                       # same weights on every rank
    x, y = T.synthetic_batch(4, spatial, seed=7)           # global batch 4 -> 2 per rank
    xs, ys = x[rank * 2:(rank + 1) * 2], y[rank * 2:(rank + 1) * 2]
    _, _, _, _, _, grads = T.generator_loss_and_grads(ys, xs, var, filters=8, num_conv=1)
    flat = torch.cat([g.reshape(-1) for g in grads.values()])
    scale = dp.allreduce_grads_(flat)
    assert dp.world() == 2 and dp.rank() == rank and scale == 0.5
    assert dp.max_over_ranks(1.0 + rank) == 2.0
    assert dp.rank_seed(123) == 123 + rank
    if rank == 0:
        torch.save({"flat": flat * scale}, out)
    dist.destroy_process_group()


def test_dp_two_ranks_equals_single_process(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    from oracle import ref_model as M, ref_train as T
    spatial = [16, 8]
    tab, _, _ = M.generator_layout(spatial + [1], filters=8, num_conv=1)
    var = M.init_variables(tab, 123)
    x, y = T.synthetic_batch(4, spatial, seed=7)
    _, _, _, _, _, grads = T.generator_loss_and_grads(y, x, var, filters=8, num_conv=1)
    ref = torch.cat([g.reshape(-1) for g in grads.values()])
    got = torch.load(out)["flat"]
    # mean over 2 shards of per-shard means == mean over the whole batch (equal shard sizes)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-7)
