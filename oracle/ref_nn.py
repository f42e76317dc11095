"""TEST INFRASTRUCTURE ONLY -- CPU oracle for arch=nn (the latent-space integrator).

Restates model.py:218-224 (NN: linear -> batch_norm+elu -> dropout, twice, then linear) and the windowed roll-out / loss of
trainer.py:586-629 (build_model_nn).  Backward = torch autograd over the restatement.

slim.batch_norm / slim.dropout / tf.nn.elu / tf.losses.mean_squared_error are TensorFlow library code that is not in
/root/reference (TF 1.15, README.md:18,23).  Restated here from their published definitions:
  * fused batch norm, training: y = gamma (x - mean_B) / sqrt(var_B + eps) + beta with the BIASED batch variance; the moving
    statistics are updated in place (updates_collections=None) as m <- decay m + (1-decay) stat, the variance statistic with
    Bessel's correction B/(B-1) (tf.nn.fused_batch_norm returns the unbiased variance).  torch.nn.functional.batch_norm
    follows the same convention and serves as an independent witness in tests/test_oracle.py.
  * inference: the moving statistics replace the batch statistics.
  * dropout(x, keep_prob): x * mask / keep_prob, mask ~ Bernoulli(keep_prob).  The random stream is not part of the
    reference's contract; masks are INPUTS of the oracle (the product's counter-based generator is restated in
    `dropout_mask` so the GPU tests can feed the oracle the very masks the kernel drew).
  * elu(x) = x if x > 0 else exp(x) - 1.
Parity status: structure (variable names / order, roll-out wiring, loss) pinned against the reference's own model.NN and
Trainer.build_model_nn through the tf shim (oracle/make_golden_nn.py); the four primitives above are unpinned restatements.
"""
from collections import OrderedDict

import numpy as np
import torch


def nn_layout(in_dim, filters, onum, name="NN"):
    """variables in creation order (model.py:218-224 through slim's default scopes)"""
    t = OrderedDict()
    dims = [(in_dim, filters * 2), (filters * 2, filters), (filters, onum)]
    for i, (a, b) in enumerate(dims):
        fc = "fully_connected" if i == 0 else "fully_connected_%d" % i
        t["%s/%s/weights" % (name, fc)] = (a, b)
        t["%s/%s/biases" % (name, fc)] = (b,)
        if i < 2:
            bn = "BatchNorm" if i == 0 else "BatchNorm_%d" % i
            t["%s/%s/beta" % (name, bn)] = (b,)
            t["%s/%s/gamma" % (name, bn)] = (b,)
            t["%s/%s/moving_mean" % (name, bn)] = (b,)
            t["%s/%s/moving_variance" % (name, bn)] = (b,)
    return t


def is_trainable(name):
    return not (name.endswith("/moving_mean") or name.endswith("/moving_variance"))


def init_variables(table, seed=123, dtype=torch.float32):
    """slim defaults: xavier-uniform weights, zero biases / beta / moving_mean, one gamma / moving_variance"""
    from .ref_ops import xavier_uniform_
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, shp in table.items():
        if k.endswith("/weights"):
            out[k] = xavier_uniform_(shp, g, dtype)
        elif k.endswith("/gamma") or k.endswith("/moving_variance"):
            out[k] = torch.ones(shp, dtype=dtype)
        else:
            out[k] = torch.zeros(shp, dtype=dtype)
    return out


def elu(x):
    """tf.nn.elu: x if x > 0 else exp(x) - 1"""
    return torch.nn.functional.elu(x)


def batch_norm(x, gamma, beta, moving_mean, moving_var, train, eps=1e-5, decay=0.9, act=elu):
    """ops.py:26-36 on [B, N]; in training mode moving_mean / moving_var are UPDATED IN PLACE"""
    if train:
        B = x.shape[0]
        mean = x.mean(dim=0)
        var = ((x - mean) ** 2).mean(dim=0)
        with torch.no_grad():
            moving_mean.mul_(decay).add_((1 - decay) * mean.detach())
            moving_var.mul_(decay).add_((1 - decay) * var.detach() * (B / (B - 1.0) if B > 1 else 1.0))
    else:
        mean, var = moving_mean, moving_var
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    return act(y) if act is not None else y


def dropout(x, keep_prob, mask):
    return x * mask.to(x.dtype) / keep_prob


def _splitmix_uniform(seed, ctr):
    """the product's counter-based generator (csrc/dfl_mlp.cu dropout_uniform): splitmix64 of (seed, counter), top 24 bits"""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (ctr.astype(np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def dropout_mask(seed, offset, shape, keep_prob):
    n = int(np.prod(shape))
    u = _splitmix_uniform(seed, np.arange(offset, offset + n, dtype=np.uint64))
    return torch.from_numpy((u < np.float32(keep_prob)).reshape(shape))


def nn_forward(x, var, train, masks=None, name="NN", keep_prob=0.1, eps=1e-5, decay=0.9):
    """model.py:218-224.  `masks`: the two dropout masks of this call (training mode only)."""
    h = x
    for i in range(2):
        fc = name + ("/fully_connected" if i == 0 else "/fully_connected_%d" % i)
        bn = name + ("/BatchNorm" if i == 0 else "/BatchNorm_%d" % i)
        h = h @ var[fc + "/weights"] + var[fc + "/biases"]
        h = batch_norm(h, var[bn + "/gamma"], var[bn + "/beta"], var[bn + "/moving_mean"], var[bn + "/moving_variance"],
                       train, eps, decay)
        if train:
            h = dropout(h, keep_prob, masks[i])
    return h @ var[name + "/fully_connected_2/weights"] + var[name + "/fully_connected_2/biases"]


def rollout(xw, var, p_num, rescale, train, masks=None, name="NN", keep_prob=0.1):
    """trainer.py:590-615: w_num chained predictions; `masks` = [w_num][2] in call order"""
    w_num = xw.shape[1]
    x_ = xw[:, 0, :]
    outs = []
    for i in range(w_num):
        y_ = nn_forward(x_, var, train, masks[i] if train else None, name, keep_prob)
        outs.append(y_.unsqueeze(1))
        if i < w_num - 1:
            x_ = torch.cat([x_[:, :-p_num] + y_ * rescale, xw[:, i + 1, -p_num:]], dim=-1)
    return torch.cat(outs, dim=1)


def nn_loss_and_grads(xw, yw, var, p_num, rescale, masks, name="NN", keep_prob=0.1):
    """loss = tf.losses.mean_squared_error(yw, yw_) (trainer.py:627,629) and its gradient w.r.t. the trainable variables;
    the moving statistics in `var` are advanced by the w_num training-mode calls, as one sess.run(optim) does"""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(is_trainable(k))) for k, v in var.items())
    yw_ = rollout(xw, leaves, p_num, rescale, True, masks, name, keep_prob)
    loss = ((yw_ - yw) ** 2).mean()
    names = [k for k in leaves if is_trainable(k)]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    for k in var:
        if not is_trainable(k):
            var[k].copy_(leaves[k].detach())
    return loss.detach(), OrderedDict(zip(names, grads)), yw_.detach()
