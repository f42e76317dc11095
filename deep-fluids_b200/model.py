"""Mirror of the reference's model.py generators (model.py:5-87), encoders (:118-188) and auto-encoders (:190-216) on the
B200 engines.

`GeneratorBE(z, filters, output_shape, name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse)` keeps the
reference's argument names and returns `(out, variables)`; `out` is the generator output on the GPU and `variables`
the ordered TF-named variable list.  `EncoderBE(3)(x, filters, z_num, ...) -> (z, variables)` and
`AE(3)(x, filters, z_num, ...) -> (out, z, variables)` likewise.  The engine object (buffers, weights) is cached per `name`
so `reuse=True` re-applies the same variables, as tf.variable_scope(reuse=True) does.
"""
import numpy as np
import torch

from .engine import GeneratorEngine
from .ops import batch_norm, conv2d, conv3d, elu, get_variables, linear, lrelu, upscale, upscale3, variable_scope
from .ops import add as ops_add
from .ops import dropout as ops_dropout

_ENGINES = {}


# ---------------------------------------------------------------------------------------------------------------------
# General form on the differentiable ops-level layers (ops.py / layers.py): any `filters`, skip_concat, any scope.  The
# fused engines below cover the configuration every BASELINE workload uses (filters=128, skip_concat=False); everything
# else a reference recipe can ask for (run.bat:56,73 train the AE with filters=64; skip_concat is a model.py option) runs
# here, layer by layer, on the same tensor-core kernels (channels zero-padded to 128-blocks), with torch autograd as the
# tape.  The model functions keep the statement order of reference model.py:5-87 / :118-216.
# ---------------------------------------------------------------------------------------------------------------------
def _repeat_num(spatial, repeat):
    rep = int(np.log2(np.max(spatial))) - 2 if repeat == 0 else repeat
    assert rep > 0 and sum(int(i) % (2 ** (rep - 1)) for i in spatial) == 0
    return rep


def generator_ops(z, filters, output_shape, name='G', num_conv=4, conv_k=3, last_k=3, repeat=0, skip_concat=False,
                  act=lrelu, reuse=False):
    nd = len(output_shape) - 1
    conv, up = (conv2d, lambda t: upscale(t, 2)) if nd == 2 else (conv3d, lambda t: upscale3(t, 2))
    with variable_scope(name, reuse=reuse) as vs:
        rep = _repeat_num(output_shape[:-1], repeat)
        x0_shape = [int(i // 2 ** (rep - 1)) for i in output_shape[:-1]] + [filters]
        n = 0
        x = linear(z, int(np.prod(x0_shape)), name='%d_fc' % n)
        n += 1
        x = x.reshape([z.shape[0]] + x0_shape)
        x0 = x
        for idx in range(rep):
            for _ in range(num_conv):
                x = conv(x, filters, k=conv_k, s=1, act=act, name='%d_conv' % n)
                n += 1
            if idx < rep - 1:
                if skip_concat:                          # model.py:30-33 / :72-75
                    x, x0 = up(x), up(x0)
                    x = torch.cat([x, x0.to(x.dtype)], dim=-1)
                else:                                    # model.py:35-37 / :77-79
                    x = up(ops_add(x, x0.to(x.dtype)))
                    x0 = x
            elif not skip_concat:
                x = ops_add(x, x0.to(x.dtype))
        out = conv(x, output_shape[-1], k=last_k, s=1, name='%d_conv' % n)
    return out, get_variables(vs)


def encoder_ops(x, filters, z_num, name='enc', num_conv=4, conv_k=3, repeat=0, act=lrelu, reuse=False):
    nd = x.dim() - 2
    conv = conv2d if nd == 2 else conv3d
    with variable_scope(name, reuse=reuse) as vs:
        rep = _repeat_num(list(x.shape[1:-1]), repeat)
        n = 0
        x = conv(x, filters, k=conv_k, s=1, act=act, name='%d_conv' % n)
        n += 1
        x0 = x
        ch = filters
        for idx in range(rep):
            for _ in range(num_conv):
                x = conv(x, filters, k=conv_k, s=1, act=act, name='%d_conv' % n)
                n += 1
            x = torch.cat([x, x0], dim=-1)               # model.py:144 / :180
            ch += filters
            if idx < rep - 1:
                x = conv(x, ch, k=conv_k, s=2, act=act, name='%d_conv' % n)
                n += 1
                x0 = x
        out = linear(x.reshape(x.shape[0], -1).float(), z_num, name='%d_fc' % n)
    return out, get_variables(vs)


def ae_ops(x, filters, z_num, name='AE', num_conv=4, conv_k=3, last_k=3, repeat=0, act=lrelu, skip_concat=False,
           use_sparse=False, reuse=False):
    with variable_scope(name, reuse=reuse) as vs:
        z, _ = encoder_ops(x, filters, z_num, 'enc', num_conv - 1, conv_k, repeat, act, reuse)
        if use_sparse:
            z = torch.sigmoid(z)
        out, _ = generator_ops(z, filters, list(x.shape[1:]), 'dec', num_conv, conv_k, last_k, repeat, skip_concat, act, reuse)
    return out, z, get_variables(vs)


def _needs_ops_path(filters, skip_concat):
    return filters != 128 or bool(skip_concat)


def _scope_engine(key, reuse, batch, build):
    """tf.variable_scope(name, reuse=reuse) for the fused engines.  reuse=False (re)creates the scope's variables;
    reuse=True must find them (TF: "Variable ... does not exist") and applies THE SAME variables: at the scope's own batch
    size the scope's engine itself, at another batch size a sibling engine built over the same flat parameter buffer (no
    copy -- an update of the variables, e.g. continued training or a checkpoint load, is seen by every sibling).  The bf16
    tensor-core operands are re-packed from the live variables on every reuse call."""
    eng = _ENGINES.get(key)
    if not reuse:
        eng = build(None)
        _ENGINES[key] = eng
        for k in [k for k in _ENGINES if len(k) == len(key) + 1 and k[:-1] == key and isinstance(k[-1], tuple)]:
            del _ENGINES[k]                  # siblings of the replaced variables
        return eng
    if eng is None:
        raise ValueError("Variable scope %r does not exist: reuse=True before the scope was built" % (key[0],))
    B = getattr(eng, "B", None) or eng.enc.B
    if B != batch:
        sib = key + (("batch", int(batch)),)
        if sib not in _ENGINES:
            _ENGINES[sib] = build(eng.params)
        eng = _ENGINES[sib]
    eng.repack()
    return eng


def _generator(z, filters, output_shape, name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse, nd):
    assert conv_k == 3 and last_k == 3, "k=3 only (the reference never uses another size on this path)"
    assert act is lrelu
    assert len(output_shape) == nd + 1
    if _needs_ops_path(filters, skip_concat):
        return generator_ops(z, filters, list(output_shape), name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse)
    eng = _scope_engine((name, nd), reuse, z.shape[0], lambda shared: GeneratorEngine(
        z.shape[0], list(output_shape), z_dim=z.shape[1], filters=filters, num_conv=num_conv, repeat=repeat, name=name,
        device=z.device, params=shared))
    out = eng.forward(z)
    return out, eng.variables


def GeneratorBE(z, filters, output_shape, name='G', num_conv=4, conv_k=3, last_k=3, repeat=0, skip_concat=False,
                act=lrelu, reuse=False):
    return _generator(z, filters, output_shape, name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse, 2)


def GeneratorBE3(z, filters, output_shape, name='G', num_conv=4, conv_k=3, last_k=3, repeat=0, skip_concat=False,
                 act=lrelu, reuse=False):
    return _generator(z, filters, output_shape, name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse, 3)


def _discriminator(x, filters, name, reuse, conv):
    """Patch discriminator of arch=dg (model.py:89-116): three stride-2 convs (filters/2, filters, 2*filters channels), one
    stride-1 conv (4*filters), one stride-1 conv to a single channel; k=3, lrelu on all but the last.  Built on the ops-level
    layers (ops.py), i.e. differentiable tcgen05 kernels with slim's default variable names D/Conv ... D/Conv_4."""
    with variable_scope(name, reuse=reuse) as vs:
        width = int(filters / 2)
        for _ in range(3):
            x = conv(x, width, k=3, act=lrelu)            # default stride 2 (ops.py:12,15)
            width *= 2
        x = conv(x, width, k=3, s=1, act=lrelu)
        out = conv(x, 1, k=3, s=1)
    return out, get_variables(vs)


def DiscriminatorPatch(x, filters, name='D', train=True, reuse=False):
    return _discriminator(x, filters, name, reuse, conv2d)


def DiscriminatorPatch3(x, filters, name='D', train=True, reuse=False):
    return _discriminator(x, filters, name, reuse, conv3d)


def NN(x, filters, onum, name='NN', act=elu, dropout=0.1, train=True, reuse=False):
    """latent-space integrator of arch=nn (model.py:218-224): linear(2*filters) -> batch_norm(act) -> dropout,
    linear(filters) -> batch_norm(act) -> dropout, linear(onum); slim's default variable names NN/fully_connected{,_1,_2}
    and NN/BatchNorm{,_1}.  (`dropout` reaches slim.dropout as its keep_prob argument, as in the reference.)"""
    with variable_scope(name, reuse=reuse) as vs:
        x = ops_dropout(batch_norm(linear(x, filters * 2), train, act=act), dropout, is_training=train)
        x = ops_dropout(batch_norm(linear(x, filters), train, act=act), dropout, is_training=train)
        out = linear(x, onum)
    return out, get_variables(vs)


def _encoder(x, filters, z_num, name, num_conv, conv_k, repeat, act, reuse, nd):
    from .encoder import EncoderEngine
    assert conv_k == 3, "k=3 only (the reference never uses another size on this path)"
    assert act is lrelu
    assert x.dim() == nd + 2, "x must be channels-last [B,(D,)H,W,C]"
    if _needs_ops_path(filters, False):
        return encoder_ops(x, filters, z_num, name, num_conv, conv_k, repeat, act, reuse)
    eng = _scope_engine((name, nd, "enc"), reuse, x.shape[0], lambda shared: EncoderEngine(
        x.shape[0], list(x.shape[1:]), filters, z_num, num_conv, repeat, name, x.device, params=shared))
    return eng.forward(x), eng.variables


def EncoderBE(x, filters, z_num, name='enc', num_conv=4, conv_k=3, repeat=0, act=lrelu, reuse=False):
    return _encoder(x, filters, z_num, name, num_conv, conv_k, repeat, act, reuse, 2)


def EncoderBE3(x, filters, z_num, name='enc', num_conv=3, conv_k=3, repeat=0, act=lrelu, reuse=False):
    return _encoder(x, filters, z_num, name, num_conv, conv_k, repeat, act, reuse, 3)


def _ae(x, filters, z_num, name, num_conv, conv_k, last_k, repeat, act, skip_concat, use_sparse, reuse, nd):
    from .encoder import AEEngine
    assert conv_k == 3 and last_k == 3, "k=3 only"
    assert act is lrelu
    assert x.dim() == nd + 2, "x must be channels-last [B,(D,)H,W,C]"
    if _needs_ops_path(filters, skip_concat):
        return ae_ops(x, filters, z_num, name, num_conv, conv_k, last_k, repeat, act, skip_concat, use_sparse, reuse)
    eng = _scope_engine((name, nd, "ae"), reuse, x.shape[0], lambda shared: AEEngine(
        x.shape[0], list(x.shape[1:]), filters, z_num, num_conv, repeat, name, x.device, use_sparse=use_sparse, params=shared))
    out, z = eng.forward(x)          # z = Enc(x, num_conv - 1) (sigmoid if use_sparse); out = Gen(z, x.shape[1:], num_conv)
    return out, z, eng.variables


def AE(x, filters, z_num, name='AE', num_conv=4, conv_k=3, last_k=3, repeat=0, act=lrelu, skip_concat=False,
       use_sparse=False, reuse=False):
    return _ae(x, filters, z_num, name, num_conv, conv_k, last_k, repeat, act, skip_concat, use_sparse, reuse, 2)


def AE3(x, filters, z_num, name='AE', num_conv=4, conv_k=3, last_k=3, repeat=0, act=lrelu, skip_concat=False,
        use_sparse=False, reuse=False):
    return _ae(x, filters, z_num, name, num_conv, conv_k, last_k, repeat, act, skip_concat, use_sparse, reuse, 3)


def get_engine(name='G', nd=2, kind=None):
    return _ENGINES[(name, nd) if kind is None else (name, nd, kind)]


def reset():
    _ENGINES.clear()
