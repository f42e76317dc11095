"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the deep-fluids train step (loss, Adam, LR schedule).

Restates trainer.py:136-184 / trainer3.py:14-63 (generator `build_model`: G_s -> curl -> jacobian ->
L1 + Jacobian-L1 loss -> Adam.minimize), trainer.py:357-396 / trainer3.py:240-279 (AE variant),
trainer.py:69-80,284-288 (LR schedule).  Backward = torch autograd over the fp32/fp64 restatement.
Parity status: stencils/loss pinned via the tf shim (oracle/make_golden.py); conv/FC/Adam unpinned
(TensorFlow 1.15 not available) -- see oracle/ref_ops.py.
"""
import math
from collections import OrderedDict

import torch

from . import ref_ops as R
from . import ref_model as M


def stencil_loss(pot_or_vel, x, w1=1.0, w2=1.0, use_curl=True):
    """Loss of the de/ae path given the network output and the target velocity x.

    2D: pot [B,H,W,>=1] -> G_=curl(pot) (trainer.py:140);  3D: pot [B,D,H,W,3] -> G_=jacobian3(pot)[1]
    (trainer3.py:18).  loss = w1*mean|G_-x| + w2*mean|J(G_)-J(x)| (trainer.py:170-172, trainer3.py:49-51).
    Returns (loss, l1, j_l1, G_)."""
    is3d = x.dim() == 5
    if use_curl:
        g = R.curl3(pot_or_vel) if is3d else R.curl(pot_or_vel)
    else:
        g = pot_or_vel
    jac = R.jacobian3 if is3d else R.jacobian
    l1 = (g - x).abs().mean()
    jl1 = (jac(g)[0] - jac(x)[0]).abs().mean()
    return w1 * l1 + w2 * jl1, l1, jl1, g


class TFAdam(object):
    """tf.train.AdamOptimizer(lr, beta1, beta2, epsilon=1e-8) semantics (trainer.py:160-162):
        lr_t = lr*sqrt(1-b2^t)/(1-b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
        theta -= lr_t * m / (sqrt(v) + eps)          (epsilon NOT bias-corrected: differs from torch.optim.Adam)
    """

    def __init__(self, var, beta1=0.5, beta2=0.999, eps=1e-8):
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.t = 0
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in var.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in var.items())

    def step(self, var, grads, lr):
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k in var:
            g = grads[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            var[k].sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


def lr_decay(step, max_step, lr_max=1e-4, lr_min=2.5e-6):
    """g_lr_update of trainer.py:74-75, evaluated with the already-incremented global step
    (it runs after the optimizer op, trainer.py:284-288; the first step uses lr_max)."""
    return lr_min + 0.5 * (lr_max - lr_min) * (math.cos(step * math.pi / max_step) + 1)


def lr_step(lr, lr_min=2.5e-6):
    """'step' schedule, trainer.py:77-78."""
    return max(lr * 0.5, lr_min)


def generator_loss_and_grads(z, x, var, filters=128, num_conv=4, repeat=0, w1=1.0, w2=1.0,
                             use_curl=True, name="G"):
    """One forward+backward of `build_model` (arch=de). Returns (loss, l1, jl1, G_, grads dict)."""
    is3d = x.dim() == 5
    cout = (3 if is3d else 1) if use_curl else x.shape[-1]
    output_shape = list(x.shape[1:-1]) + [cout]
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in var.items())
    pot = M.generator_forward(z, leaves, output_shape, filters, num_conv, repeat, name)
    loss, l1, jl1, g = stencil_loss(pot, x, w1, w2, use_curl)
    gs = torch.autograd.grad(loss, list(leaves.values()))
    grads = OrderedDict((k, gi) for k, gi in zip(leaves.keys(), gs))
    return loss.detach(), l1.detach(), jl1.detach(), g.detach(), pot.detach(), grads


def dg_losses_and_grads(z, x, g_var, d_var, filters=128, num_conv=4, repeat=0, w1=1.0, w2=1.0, w3=1.0):
    """One `sess.run([g_optim, d_optim])` worth of losses and gradients of arch=dg (trainer.py:138-184 /
    trainer3.py:16-63): D sees concat(velocity, vorticity) of the target and of the generated field;
        g_loss = w1*L1 + w2*L1(J) + w3*mean((D_G - 1)^2)            minimised over the generator variables,
        d_loss = mean((D_x - 1)^2) + mean(D_G^2)                    minimised over the discriminator variables.
    Returns (dict of loss scalars, generator grads, discriminator grads)."""
    is3d = x.dim() == 5
    gl = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in g_var.items())
    dl = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in d_var.items())
    pot = M.generator_forward(z, gl, list(x.shape[1:-1]) + [3 if is3d else 1], filters, num_conv, repeat, "G")
    loss, l1, jl1, g = stencil_loss(pot, x, w1, w2, True)
    jac = R.jacobian3 if is3d else R.jacobian
    d_x = M.discriminator_forward(torch.cat([x, jac(x)[1]], dim=-1), dl)
    d_g = M.discriminator_forward(torch.cat([g, jac(g)[1]], dim=-1), dl)
    g_loss_real = ((d_g - 1) ** 2).mean()
    d_loss_fake = (d_g ** 2).mean()
    d_loss_real = ((d_x - 1) ** 2).mean()
    g_loss = loss + w3 * g_loss_real
    d_loss = d_loss_real + d_loss_fake
    gg = torch.autograd.grad(g_loss, list(gl.values()), retain_graph=True)
    dg = torch.autograd.grad(d_loss, list(dl.values()))
    losses = {"g_loss": g_loss.detach(), "g_loss_l1": l1.detach(), "g_loss_j_l1": jl1.detach(), "g_loss_real": g_loss_real.detach(),
              "d_loss_fake": d_loss_fake.detach(), "d_loss_real": d_loss_real.detach(), "d_loss": d_loss.detach(),
              "D_x": d_x.detach(), "D_G": d_g.detach(), "pot": pot.detach(), "G_": g.detach()}
    return losses, OrderedDict(zip(gl.keys(), gg)), OrderedDict(zip(dl.keys(), dg))


def ae_loss_and_grads(x, y_last, var, p_num, filters=128, z_num=16, num_conv=4, repeat=0,
                      w1=1.0, w2=1.0, w4=1.0, use_curl=True, name="AE", use_sparse=False, sparsity=0.01, w5=1.0):
    """One forward+backward of `build_model_ae` (trainer.py:357-396 / trainer3.py:240-279).
    y_last = y[:,:,-1] [B,p_num]; loss_p = mean((y_last - z[:,-p_num:])^2)."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in var.items())
    s, z = M.ae_forward(x, leaves, filters, z_num, num_conv, repeat, name, use_sparse=use_sparse)
    loss, l1, jl1, g = stencil_loss(s, x, w1, w2, use_curl)
    loss_p = ((y_last - z[:, -p_num:]) ** 2).mean()
    total = loss + w4 * loss_p
    if use_sparse:   # trainer.py:389-394: sum_j KL(Bernoulli(rho) || Bernoulli(mean_b z[:, j])) over the non-supervised dims
        m = z[:, :-p_num].mean(0)
        rho = torch.tensor(sparsity, dtype=z.dtype)
        kl = rho * (rho.log() - m.log()) + (1 - rho) * ((1 - rho).log() - (1 - m).log())
        total = total + w5 * kl.sum()
    gs = torch.autograd.grad(total, list(leaves.values()))
    grads = OrderedDict((k, gi) for k, gi in zip(leaves.keys(), gs))
    return total.detach(), l1.detach(), jl1.detach(), loss_p.detach(), g.detach(), z.detach(), grads


def synthetic_batch(batch, spatial, seed=123, z_dim=3, dtype=torch.float32, smooth=2):
    """params y~U[-1,1] [B,z_dim]; target velocity = reference-curl of a smoothed
    N(0,1) potential scaled to max|x|=1 (mirrors x/=x_range, data.py:329)."""
    g = torch.Generator().manual_seed(seed)
    y = torch.rand(batch, z_dim, generator=g, dtype=torch.float64) * 2 - 1
    nd = len(spatial)
    pot = torch.randn([batch] + list(spatial) + [1 if nd == 2 else 3], generator=g, dtype=torch.float64)
    for _ in range(smooth):  # cheap separable box smoothing, band-limits the field
        for ax in range(1, nd + 1):
            pot = (pot + torch.roll(pot, 1, ax) + torch.roll(pot, -1, ax)) / 3.0
    x = R.curl(pot) if nd == 2 else R.curl3(pot)
    x = x / x.abs().max()
    return x.to(dtype), y.to(dtype)


def teacher_forced_backward(z, var, acts, dpot, num_conv=4, name="G", operand_round=None, mask_from_acts=False,
                            min_level=0, return_dz=False):
    """Backward pass of the generator evaluated layer by layer with torch autograd, where every layer's INPUT is the
    activation tensor the device path actually stored (`acts` = {"x0": [per level], "y": [[per conv] per level],
    "s": top-level residual sum}, fp32 CPU copies).  Because the leaky-ReLU masks are then computed from the same
    forward state as on the device, this isolates the accuracy of the backward kernels (dgrad / wgrad / bias-grad /
    pooling / FC-backward) from the sign flips that bf16 activation noise causes in a free-running comparison.
    `operand_round` (e.g. ref_model.bf16_round_ste) models the bf16 operand copy of the conv weights.
    `mask_from_acts`: take the leaky-ReLU derivative from the sign of the STORED layer output instead of recomputing
    the pre-activation here (a recomputation that differs by 1e-5 relative still flips ~1e-5 of the signs, each flip a
    5x change of that element: rel-L2 ~ sqrt(1e-5) = 3e-3, which would hide a 1e-4-grade kernel error).
    `min_level` > 0 stops after that level's convolutions (BASELINE-size checks of the finest levels only; `acts` then only
    needs those levels' tensors).  `return_dz`: also return dL/dz (the AE decoder's gradient into the latent code)."""
    rnd = operand_round if operand_round is not None else (lambda t: t)
    nd = acts["s"].dim() - 2
    rep = len(acts["x0"])
    grads = OrderedDict()
    n_last = rep * num_conv + 1

    def layer_grads(xin, wname, act, gout, yact=None):
        xin = xin.detach().clone().requires_grad_(True)
        w = var[wname + "/weights"].detach().clone().requires_grad_(True)
        b = var[wname + "/biases"].detach().clone().requires_grad_(True)
        if mask_from_acts and yact is not None:
            gout = gout * torch.where(yact > 0, torch.ones(()), torch.full((), 0.2))     # lrelu' (ops.py:13-14)
            act = None
        out = R.conv_nd(xin, rnd(w), b, 1, act)
        gx, gw, gb = torch.autograd.grad(out, [xin, w, b], gout)
        grads[wname + "/weights"], grads[wname + "/biases"] = gw, gb
        return gx

    g = layer_grads(acts["s"], "%s/%d_conv" % (name, n_last), None, dpot)      # ds
    gz = None
    for i in range(rep - 1, min_level - 1, -1):
        ds = g
        gy = ds
        for c in range(num_conv - 1, -1, -1):
            xin = acts["y"][i][c - 1] if c > 0 else acts["x0"][i]
            gy = layer_grads(xin, "%s/%d_conv" % (name, i * num_conv + c + 1), R.lrelu, gy, acts["y"][i][c])
        gx0 = gy + ds
        if i == min_level and i > 0:
            break
        if i > 0:      # adjoint of nearest x2 upsampling: sum over the children
            u = torch.zeros_like(acts["x0"][i - 1]).requires_grad_(True)
            up = R.upscale(u, 2) if nd == 2 else R.upscale3(u, 2)
            (g,) = torch.autograd.grad(up, u, gx0)
        else:
            flat = gx0.reshape(gx0.shape[0], -1)
            grads["%s/0_fc/weights" % name] = z.t() @ flat
            grads["%s/0_fc/biases" % name] = flat.sum(0)
            gz = flat @ var["%s/0_fc/weights" % name].t()
    return (grads, gz) if return_dz else grads


def teacher_forced_backward_encoder(x, var, acts, dz, num_conv=3, name="enc", operand_round=None):
    """Backward pass of EncoderBE / EncoderBE3 (model.py:118-188) layer by layer with torch autograd, every layer fed with
    the activation the device stored (same idea as `teacher_forced_backward`).  `acts` (fp32 CPU):
      "cat":  per level idx the concat tensor [B,(D,)H,W,128*(idx+2)] = concat([x, x0]) of model.py:144 / :180
              (channels 0..127 = output of the level's last conv, the rest = x0);
      "ylev": per level the outputs of its convs 0..num_conv-2.
    dz: upstream gradient w.r.t. the latent code [B, z_num].  Returns the gradients of every encoder variable."""
    rnd = operand_round if operand_round is not None else (lambda t: t)
    rep = len(acts["cat"])
    grads = OrderedDict()
    # variable numbering (model.py:129-149): 0 = first conv; per level num_conv convs (+ one stride-2 conv below the top)
    n_conv, n_s2, n = [], [], 1
    for idx in range(rep):
        n_conv.append(list(range(n, n + num_conv)))
        n += num_conv
        if idx < rep - 1:
            n_s2.append(n)
            n += 1
    n_fc = n

    def layer_grads(xin, num, stride, gout):
        wname = "%s/%d_conv" % (name, num)
        xin = xin.detach().clone().requires_grad_(True)
        w = var[wname + "/weights"].detach().clone().requires_grad_(True)
        b = var[wname + "/biases"].detach().clone().requires_grad_(True)
        out = R.conv_nd(xin, rnd(w), b, stride, R.lrelu)
        gx, gw, gb = torch.autograd.grad(out, [xin, w, b], gout)
        grads[wname + "/weights"], grads[wname + "/biases"] = gw, gb
        return gx

    top = acts["cat"][-1]
    flat = top.reshape(top.shape[0], -1)
    grads["%s/%d_fc/weights" % (name, n_fc)] = flat.t() @ dz
    grads["%s/%d_fc/biases" % (name, n_fc)] = dz.sum(0)
    gcat = (dz @ var["%s/%d_fc/weights" % (name, n_fc)].t()).reshape(top.shape)
    for idx in range(rep - 1, -1, -1):
        cat = acts["cat"][idx]
        g, gx0 = gcat[..., :128], gcat[..., 128:]
        for c in range(num_conv - 1, -1, -1):
            xin = acts["ylev"][idx][c - 1] if c > 0 else cat[..., 128:]
            g = layer_grads(xin, n_conv[idx][c], 1, g)
        gx0 = gx0 + g
        if idx > 0:
            gcat = layer_grads(acts["cat"][idx - 1], n_s2[idx - 1], 2, gx0)
        else:
            layer_grads(x, 0, 1, gx0)
    return grads
