"""Entry point with the reference's control flow (main.py:10-35):
prepare dirs -> seed -> BatchManager(config) -> Trainer3 if is_3d else Trainer -> train() / test().
Multi-GPU (new): launch with torchrun; each rank binds to LOCAL_RANK and the trainers all-reduce gradients."""
import os

import torch

from .config import get_config
from .util import prepare_dirs_and_logger, save_config


def main(config):
    prepare_dirs_and_logger(config)
    torch.manual_seed(config.random_seed)                    # tf.set_random_seed (main.py:12)

    rank = 0
    if "LOCAL_RANK" in os.environ:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        if not dist.is_initialized():
            dist.init_process_group("nccl")
        rank = dist.get_rank()

    from .trainer import Trainer
    from .trainer3 import Trainer3
    if 'nn' in config.arch:                                  # main.py:14-17
        from .data_nn import BatchManager
        batch_manager = BatchManager(config)
    else:
        from .data import BatchManager
        batch_manager = BatchManager(config, rank=rank)

    if config.is_3d:
        trainer = Trainer3(config, batch_manager)
    else:
        trainer = Trainer(config, batch_manager)

    if config.is_train:
        if rank == 0:
            save_config(config)
        trainer.train()
    else:
        if not config.load_path:
            raise Exception("[!] You should specify `load_path` to load a pretrained model")
        trainer.test()
    return trainer


if __name__ == "__main__":
    config, unparsed = get_config()
    main(config)
