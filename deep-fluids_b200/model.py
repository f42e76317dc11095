"""Mirror of the reference's model.py generators (model.py:5-87), encoders (:118-188) and auto-encoders (:190-216) on the
B200 engines.

`GeneratorBE(z, filters, output_shape, name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse)` keeps the
reference's argument names and returns `(out, variables)`; `out` is the generator output on the GPU and `variables`
the ordered TF-named variable list.  `EncoderBE(3)(x, filters, z_num, ...) -> (z, variables)` and
`AE(3)(x, filters, z_num, ...) -> (out, z, variables)` likewise.  The engine object (buffers, weights) is cached per `name`
so `reuse=True` re-applies the same variables, as tf.variable_scope(reuse=True) does.
"""
from .engine import GeneratorEngine
from .ops import conv2d, conv3d, get_variables, lrelu, variable_scope

_ENGINES = {}


def _generator(z, filters, output_shape, name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse, nd):
    assert conv_k == 3 and last_k == 3, "k=3 only (the reference never uses another size on this path)"
    assert not skip_concat, "skip_concat=True is never enabled by the reference's trainers"
    assert act is lrelu
    assert len(output_shape) == nd + 1
    key = (name, nd)
    eng = _ENGINES.get(key)
    if eng is None or not reuse or eng.B != z.shape[0]:
        init = eng.params.state_dict() if (eng is not None and reuse) else None   # reuse=True shares the variables
        eng = GeneratorEngine(z.shape[0], list(output_shape), z_dim=z.shape[1], filters=filters, num_conv=num_conv,
                              repeat=repeat, name=name, device=z.device, init=init)
        _ENGINES[key] = eng
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


def elu(x):
    """tf.nn.elu: the default activation of NN (model.py:218)"""
    raise NotImplementedError("arch='nn' (latent-space integrator, model.py:218-224) is outside the B200 hot path")


def NN(x, filters, onum, name='NN', act=elu, dropout=0.1, train=True, reuse=False):
    """latent-space MLP of arch=nn (model.py:218-224): linear(2*filters) + batch_norm + dropout, linear(filters) +
    batch_norm + dropout, linear(onum).  Not built: it is not on the conv / stencil hot path."""
    raise NotImplementedError("arch='nn' (latent-space integrator, model.py:218-224) is outside the B200 hot path")


def _encoder(x, filters, z_num, name, num_conv, conv_k, repeat, act, reuse, nd):
    from .encoder import EncoderEngine
    assert conv_k == 3, "k=3 only (the reference never uses another size on this path)"
    assert act is lrelu
    assert x.dim() == nd + 2, "x must be channels-last [B,(D,)H,W,C]"
    key = (name, nd, "enc")
    eng = _ENGINES.get(key)
    if eng is None or not reuse or eng.B != x.shape[0]:
        params = eng.params if (eng is not None and reuse and eng.B == x.shape[0]) else None
        eng = EncoderEngine(x.shape[0], list(x.shape[1:]), filters, z_num, num_conv, repeat, name, x.device, params=params)
        _ENGINES[key] = eng
    return eng.forward(x), eng.variables


def EncoderBE(x, filters, z_num, name='enc', num_conv=4, conv_k=3, repeat=0, act=lrelu, reuse=False):
    return _encoder(x, filters, z_num, name, num_conv, conv_k, repeat, act, reuse, 2)


def EncoderBE3(x, filters, z_num, name='enc', num_conv=3, conv_k=3, repeat=0, act=lrelu, reuse=False):
    return _encoder(x, filters, z_num, name, num_conv, conv_k, repeat, act, reuse, 3)


def _ae(x, filters, z_num, name, num_conv, conv_k, last_k, repeat, act, skip_concat, use_sparse, reuse, nd):
    from .encoder import AEEngine
    assert conv_k == 3 and last_k == 3, "k=3 only"
    assert not skip_concat, "skip_concat=True is never enabled by the reference's trainers"
    assert act is lrelu
    assert x.dim() == nd + 2, "x must be channels-last [B,(D,)H,W,C]"
    key = (name, nd, "ae")
    eng = _ENGINES.get(key)
    if eng is None or not reuse or eng.enc.B != x.shape[0]:
        eng = AEEngine(x.shape[0], list(x.shape[1:]), filters, z_num, num_conv, repeat, name, x.device, use_sparse=use_sparse)
        _ENGINES[key] = eng
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
