"""Mirror of the reference's model.py generators (model.py:5-87) on the B200 engine.

`GeneratorBE(z, filters, output_shape, name, num_conv, conv_k, last_k, repeat, skip_concat, act, reuse)` keeps the
reference's argument names and returns `(out, variables)`; `out` is the generator output on the GPU and `variables`
the ordered TF-named variable list.  The engine object (buffers, weights) is cached per `name` so `reuse=True`
re-applies the same variables, as tf.variable_scope(reuse=True) does.
"""
from .engine import GeneratorEngine
from .ops import lrelu

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


def get_engine(name='G', nd=2):
    return _ENGINES[(name, nd)]


def reset():
    _ENGINES.clear()
