"""arch=nn at the reference's recipe (run.bat:62: --w_size=30 --z_num=16 --filters=512 --batch_size=1024) on a synthetic code file:
time of one train step (roll-out of 30 chained NN calls, loss, backward, Adam), CUDA events, eager launches.
    python tools/nn_bench.py [--json F]"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfluids_b200 import kernels as K  # noqa: E402
from deepfluids_b200.data_nn import BatchManager  # noqa: E402
from deepfluids_b200.trainer import Trainer  # noqa: E402


def main():
    root = tempfile.mkdtemp()
    sims, frames, z, dof = 40, 400, 16, 2
    rng = np.random.RandomState(0)
    c = (rng.randn(sims, frames, z).cumsum(axis=1) * 0.05).astype(np.float64)
    p = rng.randn(sims, frames - 1, dof) * 0.05
    os.makedirs(os.path.join(root, "data")); os.makedirs(os.path.join(root, "code")); os.makedirs(os.path.join(root, "model"))
    open(os.path.join(root, "data", "args.txt"), "w").write("num_dof: %d\n" % dof)
    np.savez_compressed(os.path.join(root, "code", "code%d.npz" % z), x=c[:, :-1].reshape(-1, z), y=c[:, 1:].reshape(-1, z),
                        p=p.reshape(-1, dof), s=sims, f=frames)
    cfg = argparse.Namespace(data_path=os.path.join(root, "data"), code_path=os.path.join(root, "code"), is_3d=False, w_size=30, z_num=z,
                             batch_size=1024, random_seed=123, dataset="synthetic", data_type="velocity", arch="nn", res_x=96, res_y=128,
                             res_z=0, test_batch_size=100, repeat=0, filters=512, num_conv=4, w1=1.0, w2=1.0, use_curl=False,
                             optimizer="adam", beta1=0.5, beta2=0.999, model_dir=os.path.join(root, "model"), load_path="", start_step=0,
                             max_epoch=200, lr_update="decay", lr_min=2.5e-5, lr_max=1e-4, lr_update_step=100, log_step=10,
                             test_step=10, save_sec=3600, is_train=True)
    bm = BatchManager(cfg)
    tr = Trainer(cfg, bm)
    for _ in range(3):
        tr.train_step_nn()
    torch.cuda.synchronize()
    n0 = K.PROF.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        loss = tr.train_step_nn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"workload": "arch=nn, batch 1024, w_size 30, z_num 16, filters 512 (run.bat:62)", "ms_per_step": ms,
           "library_launches_per_step": (K.PROF.launches - n0) / steps, "windows_per_sec": 1024 / (ms * 1e-3), "loss": float(loss),
           "params": int(tr.engine.params.data.numel())}
    print(json.dumps(out))
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
