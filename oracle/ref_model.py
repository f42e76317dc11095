"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the deep-fluids model zoo on the hot path.

Restates reference model.py:5-87 (GeneratorBE / GeneratorBE3), :118-188 (EncoderBE / EncoderBE3)
and :190-216 (AE / AE3) as pure functions over an ordered dict of TF-named, TF-laid-out
variables (`<scope>/<n>_fc/{weights,biases}`, `<scope>/<n>_conv/{weights,biases}`; FC `[in,out]`,
conv `[k,(k,)k,Cin,Cout]`).  Parity status: see oracle/ref_ops.py (conv/linear unpinned, structure
restated from the reference source).
"""
from collections import OrderedDict

import numpy as np
import torch

from . import ref_ops as R


def _repeat_num(spatial, repeat):
    # model.py:9-13 / :51-55
    rep = int(np.log2(np.max(spatial))) - 2 if repeat == 0 else repeat
    assert rep > 0 and sum(int(i) % (2 ** (rep - 1)) for i in spatial) == 0
    return rep


def generator_layout(output_shape, filters=128, num_conv=4, repeat=0, z_dim=3, name="G",
                     conv_k=3, last_k=3):
    """Ordered (name -> shape) table of the generator's variables (model.py:15-42 / :57-84)."""
    spatial = list(output_shape[:-1])
    nd = len(spatial)
    rep = _repeat_num(spatial, repeat)
    x0 = [int(i // 2 ** (rep - 1)) for i in spatial]
    tab = OrderedDict()
    tab["%s/0_fc/weights" % name] = (z_dim, int(np.prod(x0)) * filters)
    tab["%s/0_fc/biases" % name] = (int(np.prod(x0)) * filters,)
    n = 1
    for _ in range(rep):
        for _ in range(num_conv):
            tab["%s/%d_conv/weights" % (name, n)] = (conv_k,) * nd + (filters, filters)
            tab["%s/%d_conv/biases" % (name, n)] = (filters,)
            n += 1
    tab["%s/%d_conv/weights" % (name, n)] = (last_k,) * nd + (filters, output_shape[-1])
    tab["%s/%d_conv/biases" % (name, n)] = (output_shape[-1],)
    return tab, rep, x0


def encoder_layout(x_shape, filters=128, z_num=16, num_conv=3, repeat=0, name="enc", conv_k=3):
    """Variables of EncoderBE/EncoderBE3 (model.py:118-152 / :154-188). x_shape = [..spatial.., C]."""
    spatial = list(x_shape[:-1])
    nd = len(spatial)
    rep = _repeat_num(spatial, repeat)
    tab = OrderedDict()
    ch = filters
    n = 0
    tab["%s/%d_conv/weights" % (name, n)] = (conv_k,) * nd + (x_shape[-1], ch)
    tab["%s/%d_conv/biases" % (name, n)] = (ch,)
    n += 1
    cur = spatial
    for idx in range(rep):
        cin = ch
        for _ in range(num_conv):
            tab["%s/%d_conv/weights" % (name, n)] = (conv_k,) * nd + (cin, filters)
            tab["%s/%d_conv/biases" % (name, n)] = (filters,)
            cin = filters
            n += 1
        ch += filters
        if idx < rep - 1:
            tab["%s/%d_conv/weights" % (name, n)] = (conv_k,) * nd + (ch, ch)
            tab["%s/%d_conv/biases" % (name, n)] = (ch,)
            n += 1
            cur = [c // 2 for c in cur]
    flat = int(np.prod(cur)) * ch
    tab["%s/%d_fc/weights" % (name, n)] = (flat, z_num)
    tab["%s/%d_fc/biases" % (name, n)] = (z_num,)
    return tab, rep


def init_variables(table, seed=123, dtype=torch.float32):
    """slim defaults: xavier-uniform weights, zero biases (ops.py:12-24)."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, shp in table.items():
        if k.endswith("weights"):
            out[k] = R.xavier_uniform_(tuple(shp), g, dtype)
        else:
            out[k] = torch.zeros(shp, dtype=dtype)
    return out


def bf16_round_ste(x):
    """Round to bf16 with a straight-through gradient: models the B200 path's bf16 *storage* of activations (fp32
    accumulation inside each layer) so that leaky-ReLU masks agree with the device; used only by parity tests."""
    return x + (x.detach().bfloat16().to(x.dtype) - x.detach())


def generator_forward(z, var, output_shape, filters=128, num_conv=4, repeat=0, name="G",
                      act=R.lrelu, keep=None, store=None):
    """GeneratorBE / GeneratorBE3 forward (model.py:5-46 / :48-87), skip_concat=False.

    z [B,z_dim]; returns out [B,*spatial,C_out].  `keep`, if a list, receives every intermediate.  `store`, if given
    (e.g. bf16_round_ste), is applied to every tensor the device path materialises (FC output, each conv output,
    each residual sum) and to the conv weights' operand copy; default None = exact fp32 reference semantics."""
    st = store if store is not None else (lambda t: t)
    spatial = list(output_shape[:-1])
    nd = len(spatial)
    rep = _repeat_num(spatial, repeat)
    x0s = [int(i // 2 ** (rep - 1)) for i in spatial]
    x = R.linear(z, var["%s/0_fc/weights" % name], var["%s/0_fc/biases" % name])
    x = st(x.reshape([-1] + x0s + [filters]))
    x0 = x
    n = 1
    up = R.upscale if nd == 2 else R.upscale3
    for idx in range(rep):
        for _ in range(num_conv):
            x = st(R.conv_nd(x, st(var["%s/%d_conv/weights" % (name, n)]), var["%s/%d_conv/biases" % (name, n)],
                             1, act))
            if keep is not None:
                keep.append(x)
            n += 1
        x = st(x + x0)                  # model.py:34 / :76 and :39 / :82
        if idx < rep - 1:
            x = up(x, 2)                # model.py:35-36 / :77-78
            x0 = x
    out = R.conv_nd(x, st(var["%s/%d_conv/weights" % (name, n)]), var["%s/%d_conv/biases" % (name, n)], 1, None)
    return out


def encoder_forward(x, var, filters=128, num_conv=3, repeat=0, name="enc", act=R.lrelu):
    """EncoderBE / EncoderBE3 forward (model.py:118-152 / :154-188)."""
    spatial = list(x.shape[1:-1])
    rep = _repeat_num(spatial, repeat)
    n = 0
    x = R.conv_nd(x, var["%s/%d_conv/weights" % (name, n)], var["%s/%d_conv/biases" % (name, n)], 1, act)
    x0 = x
    n += 1
    for idx in range(rep):
        for _ in range(num_conv):
            x = R.conv_nd(x, var["%s/%d_conv/weights" % (name, n)], var["%s/%d_conv/biases" % (name, n)], 1, act)
            n += 1
        x = torch.cat([x, x0], dim=-1)
        if idx < rep - 1:
            x = R.conv_nd(x, var["%s/%d_conv/weights" % (name, n)], var["%s/%d_conv/biases" % (name, n)], 2, act)
            n += 1
            x0 = x
    flat = x.reshape(x.shape[0], -1)
    return R.linear(flat, var["%s/%d_fc/weights" % (name, n)], var["%s/%d_fc/biases" % (name, n)])


def ae_layout(x_shape, filters=128, z_num=16, num_conv=4, repeat=0, name="AE"):
    """AE / AE3 (model.py:190-216): enc uses num_conv-1, dec = generator with output_shape = x_shape."""
    tab = OrderedDict()
    e, _ = encoder_layout(x_shape, filters, z_num, num_conv - 1, repeat, name + "/enc")
    d, _, _ = generator_layout(x_shape, filters, num_conv, repeat, z_num, name + "/dec")
    tab.update(e)
    tab.update(d)
    return tab


def ae_forward(x, var, filters=128, z_num=16, num_conv=4, repeat=0, name="AE", use_sparse=False):
    z = encoder_forward(x, var, filters, num_conv - 1, repeat, name + "/enc")
    if use_sparse:
        z = torch.sigmoid(z)
    out = generator_forward(z, var, list(x.shape[1:]), filters, num_conv, repeat, name + "/dec")
    return out, z


def discriminator_layout(in_channels, filters=128, nd=2, name="D"):
    """DiscriminatorPatch / DiscriminatorPatch3 (model.py:89-116): three stride-2 k3 convs with filters/2, filters,
    2*filters channels, one stride-1 conv with 4*filters, one stride-1 conv to 1 channel; slim's default layer names
    (`conv2d(x, d, k=3, act=lrelu)` passes no scope): D/Conv, D/Conv_1, ..., D/Conv_4."""
    tab = OrderedDict()
    cin, d = in_channels, int(filters / 2)
    widths = [d, 2 * d, 4 * d, 8 * d, 1]
    for i, co in enumerate(widths):
        ln = "%s/Conv" % name if i == 0 else "%s/Conv_%d" % (name, i)
        tab[ln + "/weights"] = (3,) * nd + (cin, co)
        tab[ln + "/biases"] = (co,)
        cin = co
    return tab


def discriminator_forward(x, var, name="D", store=None):
    """forward of DiscriminatorPatch(3) on a channels-last tensor; strides 2,2,2,1,1; lrelu on all but the last conv"""
    st = store if store is not None else (lambda t: t)
    for i, stride in enumerate((2, 2, 2, 1, 1)):
        ln = "%s/Conv" % name if i == 0 else "%s/Conv_%d" % (name, i)
        x = R.conv_nd(x, st(var[ln + "/weights"]), var[ln + "/biases"], stride, R.lrelu if i < 4 else None)
        if i < 4:
            x = st(x)
    return x


def count_params(table):
    return int(sum(int(np.prod(s)) for s in table.values()))
