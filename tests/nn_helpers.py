"""shared fixtures of the arch=nn tests (CPU and GPU): the golden vectors and a synthetic code file"""
import argparse
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nn_wiring.npz")


def _golden():
    z = np.load(GOLD)
    B, F, Z, P, W = [int(v) for v in z["dims"]]
    var = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("var/")}
    masks = [[torch.from_numpy(z["mask/%d/%d" % (i, j)]) for j in range(2)] for i in range(1 + W)]
    return z, (B, F, Z, P, W), var, masks


def _write_codes(root, sims=20, frames=12, z_num=5, dof=2, seed=3):
    rng = np.random.RandomState(seed)
    c = rng.randn(sims, frames, z_num).cumsum(axis=1) * 0.3
    p = rng.randn(sims, frames - 1, dof) * 0.05
    os.makedirs(os.path.join(root, "data"), exist_ok=True)
    os.makedirs(os.path.join(root, "code"), exist_ok=True)
    with open(os.path.join(root, "data", "args.txt"), "w") as f:
        f.write("num_dof: %d\nnum_param: 3\n" % dof)
    np.savez_compressed(os.path.join(root, "code", "code%d.npz" % z_num), x=c[:, :-1].reshape(-1, z_num), y=c[:, 1:].reshape(-1, z_num),
                        p=p.reshape(-1, dof), s=sims, f=frames)
    return c, p


def _cfg(root, **kw):
    d = dict(data_path=os.path.join(root, "data"), code_path=os.path.join(root, "code"), is_3d=False, w_size=4, z_num=5, batch_size=8,
             random_seed=123)
    d.update(kw)
    return argparse.Namespace(**d)


