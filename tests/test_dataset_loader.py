"""CPU test of the real-dataset BatchManager on a tiny synthetic mantaflow-style dataset written in the
reference's on-disk format (scene/smoke_pos_size.py:121-126,215-234): args.txt, v/*.npz {x, y}, v_range.txt."""
import os

import numpy as np
import torch

from deepfluids_b200 import config as C
from deepfluids_b200 import data as D


def _write_dataset(root, n0=3, n1=2, nf=4, H=8, W=6):
    os.makedirs(os.path.join(root, "v"))
    with open(os.path.join(root, "args.txt"), "w") as f:
        for k, v in [("num_param", 3), ("p0", "src_x_pos"), ("p1", "src_radius"), ("p2", "frames"),
                     ("min_src_x_pos", 0.2), ("max_src_x_pos", 0.8), ("num_src_x_pos", n0),
                     ("min_src_radius", 0.04), ("max_src_radius", 0.12), ("num_src_radius", n1),
                     ("min_frames", 0), ("max_frames", nf - 1), ("num_frames", nf), ("resolution_x", W), ("resolution_y", H)]:
            f.write("%s: %s\n" % (k, v))
    rng = np.random.RandomState(0)
    vmin, vmax = 0.0, 0.0
    for i in range(n0):
        for j in range(n1):
            for t in range(nf):
                x = rng.randn(H, W, 2).astype(np.float32) * 3
                vmin, vmax = min(vmin, x.min()), max(vmax, x.max())
                y = [0.2 + 0.6 * i / (n0 - 1), 0.04 + 0.08 * j / (n1 - 1), t]
                np.savez_compressed(os.path.join(root, "v", "%d_%d_%d.npz" % (i, j, t)), x=x, y=y)
    with open(os.path.join(root, "v_range.txt"), "w") as f:
        f.write("%f\n%f" % (vmin, vmax))
    return max(abs(vmin), abs(vmax))


def test_dataset_batch_manager(tmp_path):
    root = str(tmp_path / "data" / "toy")
    x_range = _write_dataset(root)
    cfg, _ = C.get_config(["--dataset=toy", "--data_dir=" + str(tmp_path / "data"), "--res_x=6", "--res_y=8",
                           "--batch_size=4", "--num_worker=2"])
    cfg.data_path = root
    bm = D.BatchManager(cfg, device=torch.device("cpu"))
    assert isinstance(bm, D.DatasetBatchManager)
    assert bm.num_samples == 24 and bm.c_num == 3 and bm.y_num == [3, 2, 4]
    assert abs(bm.x_range - x_range) < 1e-5 and bm.epochs_per_step == 4 / 24.0
    x, y = bm.batch()
    bm.stop_thread()
    assert x.shape == (4, 8, 6, 2) and y.shape == (4, 3)
    assert float(x.abs().max()) <= 1.0 + 1e-6 and float(y.abs().max()) <= 1.0 + 1e-6
    # one sample against a hand normalisation (data.py:329-332)
    xs, ys = D.preprocess(os.path.join(root, "v", "2_1_3.npz"), "velocity", bm.x_range, bm.y_range)
    raw = np.load(os.path.join(root, "v", "2_1_3.npz"))
    assert np.allclose(xs, raw["x"] / x_range) and np.allclose(ys, [1.0, 1.0, 1.0])
    xd, yd = bm.denorm(torch.from_numpy(xs), torch.from_numpy(ys[None]))
    assert np.allclose(xd.numpy(), raw["x"], atol=1e-5) and np.allclose(yd.numpy()[0], raw["y"], atol=1e-5)


def test_factory_falls_back_to_synthetic_without_dataset(tmp_path):
    cfg, _ = C.get_config(["--dataset=missing", "--data_dir=" + str(tmp_path)])
    # no args.txt -> synthetic source class is selected (constructing it needs the GPU, so only the choice is checked)
    root = os.path.join(cfg.data_dir, cfg.dataset)
    assert not os.path.exists(os.path.join(root, "args.txt"))
