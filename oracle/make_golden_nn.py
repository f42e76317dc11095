"""TEST INFRASTRUCTURE ONLY.  Pins the oracle's arch=nn restatement (oracle/ref_nn.py) against the reference's own
`model.NN` (model.py:218-224) and `Trainer.build_model_nn` (trainer.py:586-629), IN THE BUILD CONTAINER.

Both are imported unchanged and run under the structural shim (oracle/tf_shim.py) extended by `install_nn`: the layer
sequence, slim's default variable scopes (fully_connected[_k], BatchNorm[_k]), the train / inference split, the roll-out
over the window (re-normalisation by out_std / code_std, splice of the next frame's parameters), the loss expression, the
optimizer construction and `minimize(loss, global_step, var_list)` are the reference's code; the arithmetic of
slim.fully_connected / batch_norm / dropout / elu / mean_squared_error is the oracle's restated primitives (TensorFlow is
not installable here), dropout masks are drawn by this script and handed to both sides.  Asserted while generating:
  * NN(x) variable names / shapes / order == ref_nn.nn_layout; outputs (train and inference mode) bit-identical,
  * yw_, ytw_, loss_train_w, l_test, l_test_w of the reference-built graph == ref_nn.rollout / the mse, bit for bit,
  * d loss / d trainable variables (torch autograd through the reference-built expression) == ref_nn.nn_loss_and_grads,
  * moving statistics after the build == the oracle's after the same sequence of training-mode calls,
  * Adam only (trainer.py:619-622), minimize(loss_train_w, global_step=step, var_list=all NN variables).
Writes tests/golden/nn_wiring.npz (inputs, masks, variables, losses, gradients).

    python -m oracle.make_golden_nn
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import tf_shim  # noqa: E402
from oracle import ref_nn as N  # noqa: E402
from oracle import ref_ops as R  # noqa: E402

SEED, B, FILTERS, Z_NUM, P_NUM, W_NUM, KEEP = 20261019, 6, 8, 5, 2, 3, 0.1
OUT_STD, CODE_STD = 0.37, 1.9


def make_inputs():
    g = torch.Generator().manual_seed(SEED)
    tab = N.nn_layout(Z_NUM + P_NUM, FILTERS, Z_NUM)
    var = N.init_variables(tab, SEED)
    for k in var:                                      # non-trivial values everywhere (also the moving statistics)
        if k.endswith("biases") or k.endswith("beta") or k.endswith("moving_mean"):
            var[k] = torch.randn(var[k].shape, generator=g) * 0.1
        elif k.endswith("gamma") or k.endswith("moving_variance"):
            var[k] = 1.0 + 0.2 * torch.rand(var[k].shape, generator=g)
    r = lambda *s: torch.randn(*s, generator=g)
    data = {"x": r(B, Z_NUM + P_NUM), "y": r(B, Z_NUM), "xt": r(B, Z_NUM + P_NUM), "yt": r(B, Z_NUM),
            "xw": r(B, W_NUM, Z_NUM + P_NUM), "yw": r(B, W_NUM, Z_NUM), "xtw": r(B, W_NUM, Z_NUM + P_NUM), "ytw": r(B, W_NUM, Z_NUM)}
    # keep_prob 0.1 on 16 / 8 features: draw until every mask keeps something, so the gradients are not trivially zero
    def mask(n):
        while True:
            m = torch.rand(B, n, generator=g) < 0.5          # the oracle takes masks as inputs; 0.5 keeps the check dense
            if m.any():
                return m
    # build_model_nn makes 1 + w_num training-mode NN calls (self.y_, then the roll-out), 2 masks each
    masks = [[mask(FILTERS * 2), mask(FILTERS)] for _ in range(1 + W_NUM)]
    return tab, var, data, masks


def main(reference_root="/root/reference", out_dir=os.path.join(ROOT, "tests", "golden")):
    tab, var, data, masks = make_inputs()
    leaves = {k: v.clone().requires_grad_(N.is_trainable(k)) for k, v in var.items()}
    store = tf_shim.VariableStore(leaves)
    tf_shim.install_structural(store, R.conv_nd, R.linear)
    record = {}
    tf_shim.install_training(record)
    flat_masks = [m for pair in masks for m in pair]
    tf_shim.install_nn(store, N.batch_norm, N.dropout, flat_masks, lambda a, b: ((a - b) ** 2).mean())
    trainer_mod, _, _ = tf_shim.import_reference_trainers(reference_root)
    model = sys.modules.get("_dfl_reference_model") or tf_shim.import_reference_model(reference_root)

    # --- model.NN alone: names / order / inference output
    # (a reference call in inference mode does not touch the moving statistics)
    out_inf, names = trainer_mod.NN(data["xt"], FILTERS, Z_NUM, train=False, reuse=False)
    assert list(names) == list(tab.keys()), (names, list(tab.keys()))
    assert [s for _, s in store.requested] == [tuple(s) for s in tab.values()]
    assert torch.equal(out_inf.detach(), N.nn_forward(data["xt"], var, False))

    # --- Trainer.build_model_nn on a stub self
    class BM(object):
        out_std, code_std = OUT_STD, CODE_STD

    t = object.__new__(trainer_mod.Trainer)
    for k, v in data.items():
        setattr(t, k, v)
    t.filters, t.z_num, t.p_num, t.w_num, t.batch_manager = FILTERS, Z_NUM, P_NUM, W_NUM, BM()
    t.optimizer, t.g_lr, t.beta1, t.beta2, t.step = "adam", "g_lr-variable", 0.5, 0.999, "step-variable"
    try:
        t.build_model_nn()
        raise AssertionError("build_model_nn did not reach its first placeholder")
    except tf_shim.BuildDone:
        pass                                              # stopped at self.loss_test = tf.placeholder (trainer.py:633)
    assert not flat_masks, "the reference made fewer training-mode dropout calls than expected"
    assert record["optimizer"]["kind"] == "adam" and record["optimizer"]["args"] == ("g_lr-variable",)
    assert record["optimizer"]["kwargs"] == {"beta1": 0.5, "beta2": 0.999}
    assert t.loss is t.loss_train_w and list(t.var) == list(tab.keys())

    # --- the oracle on the same inputs, same call order: NN(self.x) [training], then per window step NN(x_) [training]
    ovar = {k: v.clone() for k, v in var.items()}
    y_ = N.nn_forward(data["x"], ovar, True, masks[0], keep_prob=KEEP)
    rescale = OUT_STD / CODE_STD
    loss, grads, yw_ = N.nn_loss_and_grads(data["xw"], data["yw"], ovar, P_NUM, rescale, masks[1:], keep_prob=KEEP)
    assert torch.equal(t.y_.detach(), y_) and torch.equal(t.yw_.detach(), yw_)
    assert torch.equal(t.loss_train_w.detach(), loss)
    assert torch.equal(t.loss_train.detach(), ((y_ - data["y"]) ** 2).mean())
    for k in tab:
        if not N.is_trainable(k):
            assert torch.equal(leaves[k].detach(), ovar[k]), k       # moving statistics advanced identically
    ref_grads = torch.autograd.grad(t.loss_train_w, [leaves[k] for k in grads])
    worst = 0.0
    for (k, g), rg in zip(grads.items(), ref_grads):
        worst = max(worst, float((g - rg).abs().max()) / max(float(g.abs().max()), 1e-30))
    assert worst <= 1e-6, worst
    # inference-mode graph (yt_, ytw_ use the statistics as they are when evaluated: here, after the build's updates)
    ytw_ = N.rollout(data["xtw"], ovar, P_NUM, rescale, False)
    # the reference's yt_/ytw_ tensors were computed DURING the build, interleaved with the training-mode calls; re-evaluate
    # the reference NN now for an apples-to-apples check of the inference roll-out wiring
    xt_ = data["xtw"][:, 0, :]
    outs = []
    for i in range(W_NUM):
        o, _ = trainer_mod.NN(xt_, FILTERS, Z_NUM, train=False, reuse=True)
        outs.append(o.unsqueeze(1))
        if i < W_NUM - 1:
            xt_ = torch.cat([xt_[:, :-P_NUM] + o * rescale, data["xtw"][:, i + 1, -P_NUM:]], dim=-1)
    assert torch.equal(torch.cat(outs, 1).detach(), ytw_)
    assert t.ytw_.shape == ytw_.shape and t.l_test.dim() == 0

    # training without Adam must raise (trainer.py:621-622)
    t2 = object.__new__(trainer_mod.Trainer)
    t2.__dict__.update(t.__dict__)
    t2.optimizer = "gd"
    flat_masks.extend(m for pair in masks for m in pair)
    try:
        t2.build_model_nn()
        raise AssertionError("non-Adam optimizer accepted")
    except tf_shim.BuildDone:
        raise AssertionError("non-Adam optimizer accepted")
    except Exception as e:
        assert "optimizer other than Adam" in str(e).replace("opimizer", "optimizer"), e

    blob = {"loss": loss.numpy(), "yw_": yw_.numpy(), "y_": y_.numpy(), "ytw_": ytw_.numpy(), "rescale": np.float64(rescale),
            "dims": np.array([B, FILTERS, Z_NUM, P_NUM, W_NUM]), "keep": np.float64(KEEP)}
    for k, v in data.items():
        blob["in/" + k] = v.numpy()
    for i, pair in enumerate(masks):
        for j, m in enumerate(pair):
            blob["mask/%d/%d" % (i, j)] = m.numpy()
    for k, v in var.items():
        blob["var/" + k] = v.numpy()
    for k, v in ovar.items():
        if not N.is_trainable(k):
            blob["stats_after/" + k] = v.numpy()
    for k, g in grads.items():
        blob["grad/" + k] = g.numpy()
    np.savez_compressed(os.path.join(out_dir, "nn_wiring.npz"), **blob)
    print("model.NN / Trainer.build_model_nn: variables, roll-out, loss, moving statistics == oracle "
          "(max relative gradient difference %.1e)" % worst)
    print("written", os.path.join(out_dir, "nn_wiring.npz"))


if __name__ == "__main__":
    main()
