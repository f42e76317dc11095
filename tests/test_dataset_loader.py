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


def test_mirror_matches_the_reference_batch_manager(tmp_path):
    """The reference's own data.BatchManager.__init__ / preprocess / denorm (imported unchanged through oracle/tf_shim.py;
    only in the build container, where /root/reference exists) on the same toy dataset: file order, counts, ranges,
    normalisation and de-normalisation must agree with the mirror -- for the `de` layout and the `ae` layout (files sorted
    by scene * num_frames + frame, labels [dof, frames])."""
    import argparse
    import pytest
    if not os.path.exists("/root/reference/data.py"):
        pytest.skip("reference source not present on this machine")
    from oracle import tf_shim
    ref = tf_shim.import_reference_data("/root/reference")
    root = str(tmp_path / "data" / "toy")
    _write_dataset(root)
    for arch in ("de", "ae"):
        if arch == "ae":       # AE scenes: files "<scene>_<frame>.npz", y = [dof, frames] history, args carry num_dof
            import shutil
            shutil.rmtree(os.path.join(root, "v"))
            os.makedirs(os.path.join(root, "v"))
            with open(os.path.join(root, "args.txt"), "a") as f:
                f.write("num_dof: 2\n")
            rng = np.random.RandomState(1)
            for sc in range(11):
                for fr in range(4):
                    np.savez_compressed(os.path.join(root, "v", "%d_%d.npz" % (sc, fr)),
                                        x=rng.randn(8, 6, 2).astype(np.float32), y=rng.rand(2, 4).astype(np.float32) * 2 - 1)
        cfg, _ = C.get_config(["--dataset=toy", "--data_dir=" + str(tmp_path / "data"), "--res_x=6", "--res_y=8", "--batch_size=4",
                               "--num_worker=2", "--arch=" + arch])
        cfg.data_path = root
        rb = ref.BatchManager(argparse.Namespace(**vars(cfg)))
        mb = D.BatchManager(cfg, device=torch.device("cpu"))
        assert isinstance(mb, D.DatasetBatchManager)
        assert [os.path.basename(p) for p in mb.paths] == [os.path.basename(p) for p in rb.paths]
        assert mb.num_samples == rb.num_samples and mb.epochs_per_step == rb.epochs_per_step and mb.c_num == rb.c_num
        assert mb.depth == rb.depth and list(mb.y_num) == list(rb.y_num) and mb.batch_size == rb.batch_size
        assert abs(mb.x_range - float(rb.x_range)) < 1e-12 and [list(r) for r in mb.y_range] == [list(r) for r in rb.y_range]
        if arch == "ae":
            assert mb.dof == rb.dof == 2 and os.path.basename(mb.paths[5]) == "1_1.npz" and os.path.basename(mb.paths[-1]) == "10_3.npz"
        for path in (mb.paths[0], mb.paths[len(mb.paths) // 2], mb.paths[-1]):
            xr, yr = ref.preprocess(path, cfg.data_type, rb.x_range, rb.y_range)
            xm, ym = D.preprocess(path, cfg.data_type, mb.x_range, mb.y_range)
            assert np.asarray(xm).dtype == np.float32 == np.asarray(xr).dtype
            np.testing.assert_array_equal(np.asarray(xm), np.asarray(xr))
            np.testing.assert_array_equal(np.asarray(ym), np.asarray(yr).astype(np.float32))     # TF casts the fed labels to fp32
            if arch == "de":
                xd_r, yd_r = rb.denorm(x=np.array(xr, copy=True), y=np.array(yr, copy=True)[None])
                xd_m, yd_m = mb.denorm(torch.from_numpy(np.array(xm, copy=True)), torch.from_numpy(np.array(ym, copy=True)[None]))
                np.testing.assert_allclose(xd_m.numpy(), xd_r, rtol=1e-6)
                np.testing.assert_allclose(yd_m.numpy(), yd_r, rtol=1e-6)
        mb.stop_thread()
