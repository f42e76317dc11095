"""GPU parity of arch=nn (model.NN, Trainer.build_model_nn / train_nn / test_nn) against the CPU oracle oracle/ref_nn.py,
which is pinned against the reference's own model.py / trainer.py by oracle/make_golden_nn.py.

  * dfl_bn_act_fwd / dfl_bn_act_bwd (training and inference mode; none / lrelu / elu) vs oracle autograd;
  * dfl_dropout: bit-identical masks to the oracle's restatement of the counter-based generator;
  * model.NN forward (inference) and one full roll-out step -- loss, every gradient, moving statistics, Adam update -- on
    the golden inputs of tests/golden/nn_wiring.npz, the oracle being fed the masks the kernel drew;
  * Trainer(arch=nn) end to end on a synthetic code file: train, checkpoint round trip, test_nn -> code_out.npz.
Tolerances (fp32 everywhere, different summation order): 2e-5 relative to max|reference| for activations / gradients.
"""
import argparse
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_nn as N
from oracle import ref_train as T
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from nn_helpers import _golden, _write_codes, _cfg  # noqa: E402


def dev():
    return torch.device("cuda:0")


def close(a, b, tol=2e-5):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("M,n", [(8, 16), (37, 200), (256, 1024), (1, 5)])
def test_bn_act_kernels_vs_oracle(M, n, act):
    K = importlib.import_module("deep-fluids_b200.kernels")
    from oracle import ref_ops as R
    fn = {0: None, 1: R.lrelu, 2: N.elu}[act]
    g = torch.Generator().manual_seed(M * 7 + n + act)
    x = torch.randn(M, n, generator=g) * 2 + 0.5
    gamma, beta = 1 + 0.3 * torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.2
    mm, mv = torch.randn(n, generator=g) * 0.1, 1 + torch.rand(n, generator=g)
    dy = torch.randn(M, n, generator=g)
    for training in (True, False):
        if training and M == 1:
            continue
        xo, go, bo = x.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
        mmo, mvo = mm.double().clone(), mv.double().clone()
        yo = N.batch_norm(xo, go, bo, mmo, mvo, training, 1e-5, 0.9, act=fn)
        mmd, mvd = mm.to(dev()), mv.to(dev())
        y, sm, sr = K.bn_act_fwd(x.to(dev()), gamma.to(dev()), beta.to(dev()), mmd, mvd, 1e-5, 0.9, training, act)
        assert close(y, yo)
        assert close(mmd, mmo) and close(mvd, mvo)
        if training:
            dxo, dgo, dbo = torch.autograd.grad(yo, [xo, go, bo], dy.double())
            dx, dg, db = K.bn_act_bwd(x.to(dev()), y, dy.to(dev()), gamma.to(dev()), sm, sr, act)
            assert close(dx, dxo, 1e-4) and close(dg, dgo, 1e-4) and close(db, dbo, 1e-4)


def test_dropout_kernel_bit_identical_to_oracle_stream():
    K = importlib.import_module("deep-fluids_b200.kernels")
    x = torch.randn(333, 77)
    for seed, off, keep in ((123, 0, 0.1), (5, 1000, 0.5), (2 ** 40 + 3, 2 ** 33, 0.9)):
        y = K.dropout(x.to(dev()), keep, seed, off).cpu()
        m = N.dropout_mask(seed, off, x.shape, keep)
        assert torch.equal(y != 0, m & (x != 0))
        assert torch.equal(y[m], (x * np.float32(1.0 / np.float32(keep)))[m])


def _engine_from_golden():
    oe = importlib.import_module("deep-fluids_b200.ops_engine")
    ops = importlib.import_module("deep-fluids_b200.ops")
    z, (B, F, Z, P, W), var, _ = _golden()
    eng = oe.OpsNNEngine(B, Z + P, F, Z, P, W, float(z["rescale"]), 1.0, dev(), seed=77, dropout=0.5)
    assert eng.variables == list(var.keys())                 # the reference's names, in its creation order
    for k, v in var.items():
        eng.params.p(k).copy_(v.to(dev()))
        assert ops.get_variable(k).data_ptr() == eng.params.p(k).data_ptr()
        assert ops.get_variable(k).requires_grad == N.is_trainable(k)
    return eng, z, (B, F, Z, P, W), var


def test_nn_forward_inference_vs_oracle():
    eng, z, dims, var = _engine_from_golden()
    xt = torch.from_numpy(z["in/xt"])
    before = eng.params.data.clone()
    with torch.no_grad():
        y = eng.net(xt.to(dev()), False)
    assert close(y, N.nn_forward(xt.double(), {k: v.double() for k, v in var.items()}, False))
    assert torch.equal(before, eng.params.data)              # inference does not touch the moving statistics
    xtw = torch.from_numpy(z["in/xtw"])
    with torch.no_grad():
        yw = eng.rollout(xtw.to(dev()), False)
    assert close(yw, N.rollout(xtw.double(), {k: v.double() for k, v in var.items()}, dims[3], float(z["rescale"]), False))


def test_nn_rollout_step_vs_oracle():
    """one sess.run(optim) of trainer.py:645: roll-out loss, gradients, moving statistics, TF-Adam update"""
    ops = importlib.import_module("deep-fluids_b200.ops")
    eng, z, (B, F, Z, P, W), var = _engine_from_golden()
    xw, yw = torch.from_numpy(z["in/xw"]), torch.from_numpy(z["in/yw"])
    keep = 0.5
    ops.dropout_seed(4242)
    # the masks the kernels will draw: call order = window step, then layer; offsets advance by the tensor sizes
    masks, off = [], 0
    for i in range(W):
        pair = []
        for n in (2 * F, F):
            pair.append(N.dropout_mask(4242, off, (B, n), keep))
            off += B * n
        masks.append(pair)
    ovar = {k: v.double().clone() for k, v in var.items()}
    loss_o, grads_o, yw_o = N.nn_loss_and_grads(xw.double(), yw.double(), ovar, P, float(z["rescale"]), masks, keep_prob=keep)

    eng.zero_grad()
    loss = eng.loss_and_grads(xw.to(dev()), yw.to(dev()))
    assert abs(float(loss) - float(loss_o)) <= 2e-5 * abs(float(loss_o))
    # the biases in front of a batch norm have an exactly zero gradient (the batch mean is subtracted again): fp32 leaves
    # summation noise there, so every gradient is compared relative to the largest gradient of the step
    gmax = max(float(g.abs().max()) for g in grads_o.values())
    for k, g in grads_o.items():
        d = float((eng.params.g(k).double().cpu() - g).abs().max())
        assert d <= 1e-4 * max(float(g.abs().max()), 1e-2 * gmax), (k, d)
    for k in ovar:
        if not N.is_trainable(k):
            assert close(eng.params.p(k), ovar[k]), k
            assert float(eng.params.g(k).abs().max()) == 0.0
    # Adam (TF semantics) on the flat buffer: the statistics, which have no gradient, must not move
    names = list(grads_o.keys())
    adam = T.TFAdam({k: ovar[k] for k in names}, 0.5, 0.999)
    adam.step({k: ovar[k] for k in names}, grads_o, 1e-3)
    stats = {k: eng.params.p(k).clone() for k in ovar if not N.is_trainable(k)}
    eng.adam_step(1e-3, 0.5, 0.999, 1e-8, 1.0)
    for k in names:
        if k.endswith("biases") and "fully_connected_2" not in k:
            continue      # zero-gradient variables: Adam normalises the fp32 noise to steps of order lr (in TF as well); BN cancels them
        assert close(eng.params.p(k), ovar[k], 2e-4), k
    for k, v in stats.items():
        assert torch.equal(eng.params.p(k), v), k


def _nn_config(root, **kw):
    d = dict(is_3d=False, dataset="synthetic", data_type="velocity", arch="nn", res_x=8, res_y=8, res_z=0, test_batch_size=8, repeat=0,
             filters=32, num_conv=4, w1=1.0, w2=1.0, use_curl=False, optimizer="adam", beta1=0.5, beta2=0.999,
             model_dir=os.path.join(root, "model"), load_path="", start_step=0, max_epoch=2, lr_update="decay", lr_min=2.5e-5,
             lr_max=1e-3, lr_update_step=100, log_step=10, test_step=10, save_sec=3600, is_train=True)
    d.update(kw)
    cfg = _cfg(root, **d)
    os.makedirs(cfg.model_dir, exist_ok=True)
    return cfg


def test_trainer_nn_end_to_end(tmp_path):
    data_nn = importlib.import_module("deep-fluids_b200.data_nn")
    tr = importlib.import_module("deep-fluids_b200.trainer")
    root = str(tmp_path)
    c, p = _write_codes(root, sims=40, frames=12)
    cfg = _nn_config(root)
    bm = data_nn.BatchManager(cfg, device=dev())
    t = tr.Trainer(cfg, bm)
    assert t.max_step == int(cfg.max_epoch // bm.epochs_per_step) and t.log_step == bm.train_steps
    assert t.var == list(N.nn_layout(5 + 2, 32, 5).keys())
    l0 = t._test_losses_nn()
    first = float(t.train_step_nn())
    t.train()
    l1 = t._test_losses_nn()
    assert np.isfinite(first) and np.isfinite(l1[0]) and np.isfinite(l1[1])
    assert l1[0] < l0[0]                                     # the single-step test loss goes down
    assert os.path.exists(os.path.join(cfg.model_dir, "model.pt"))
    # the TensorFlow bundle carries slim's names; Adam slots only for the trainable variables
    tfc = importlib.import_module("deep-fluids_b200.tf_checkpoint")
    names = {n for n, _, _ in tfc.list_variables(tfc.latest_checkpoint(cfg.model_dir))}
    assert {"NN/BatchNorm/moving_mean", "NN/BatchNorm_1/moving_variance", "NN/fully_connected_2/weights/Adam_1", "beta1_power"} <= names
    assert "NN/BatchNorm/moving_mean/Adam" not in names
    # a fresh trainer restored from the checkpoint integrates identically
    cfg2 = _nn_config(root, is_train=False, load_path=cfg.model_dir)
    bm2 = data_nn.BatchManager(cfg2, device=dev())
    t2 = tr.Trainer(cfg2, bm2)
    assert torch.equal(t2.engine.params.data, t.engine.params.data) and t2.step == t.step
    path = t2.test_nn()
    out = np.load(path)
    sims_test, frames, zn = bm2.num_test_scenes, 12, 5
    assert out["z_out"].shape == (sims_test, frames, zn) and out["z_gt"].shape == (sims_test, frames, zn)
    # z_gt is the de-normalised ground-truth code sequence of the test simulations; frame 0 of z_out is the seed frame
    assert np.allclose(out["z_gt"], c[-sims_test:], atol=1e-5)
    assert np.allclose(out["z_out"][:, 0], c[-sims_test:, 0], atol=1e-5)
    assert np.allclose(out["z_out"][:, :, -2:], out["z_gt"][:, :, -2:], atol=1e-5)       # the p_num tail is overwritten by gt
    assert np.isfinite(out["z_out"]).all()


def test_trainer_nn_rejects_other_optimizers(tmp_path):
    data_nn = importlib.import_module("deep-fluids_b200.data_nn")
    tr = importlib.import_module("deep-fluids_b200.trainer")
    root = str(tmp_path)
    _write_codes(root)
    cfg = _nn_config(root, optimizer="gd")
    with pytest.raises(Exception, match="other than Adam"):
        tr.Trainer(cfg, data_nn.BatchManager(cfg, device=dev()))
