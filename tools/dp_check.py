"""torchrun --nproc-per-node 2 tools/dp_check.py : data-parallel sanity on real GPUs (NCCL).
Each rank trains 3 steps on its own shard; after every step all ranks must hold bit-identical parameters (identical
all-reduced gradients -> identical Adam updates), and the loss must be finite."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from deepfluids_b200 import config as C
from deepfluids_b200.data import BatchManager
from deepfluids_b200.trainer3 import Trainer3
cfg, _ = C.get_config(["--synthetic=true", "--is_3d=true", "--res_x=32", "--res_y=32", "--res_z=32", "--batch_size=2",
                       "--num_conv=2", "--max_step=100"])
bm = BatchManager(cfg, rank=dist.get_rank(), pool=2)
tr = Trainer3(cfg, bm)
for i in range(3):
    tr.train_step()
    tr.update_lr(i)
    p = tr.engine.params.data
    ref = p.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, p))
    flags = torch.tensor([1.0 if same else 0.0], device=p.device)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    loss = tr.losses()[0]
    if dist.get_rank() == 0:
        print("dp_check step %d: params identical on all ranks = %s, loss(rank0) = %.5f" % (i, bool(flags.item()), loss), flush=True)
    assert flags.item() == 1.0 and loss == loss
if dist.get_rank() == 0:
    print("dp_check OK (world=%d)" % dist.get_world_size())
dist.destroy_process_group()
