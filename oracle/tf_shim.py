"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A minimal stand-in for the `tensorflow` module (TF 1.15 is the reference's only
compute dependency, README.md:18,23, and cannot be installed here: no network,
no cp312 wheel).  Installing it in `sys.modules` lets the reference's own
`ops.py` be imported *unchanged* from /root/reference and its pure
slicing/concat/stack functions (`curl` ops.py:264-274, `jacobian` :205-225,
`jacobian3` :227-262, `divergence` :276-284, `divergence3` :286-290, `lrelu`
:9-10, `upscale`/`upscale3` :66-91) be executed verbatim on torch CPU tensors,
forward and (through torch autograd) backward.

Used only by `oracle/make_golden.py` (in the build container, where
/root/reference exists) to pin `oracle/ref_ops.py` and to write the committed
fixtures under `tests/golden/`.  Everything the reference delegates to TF
*library kernels* (slim.conv2d/conv3d/fully_connected, AdamOptimizer) is NOT
covered by this shim -- see `oracle/ref_ops.py` for the restatement and the
"parity unpinned" note there.
"""
import sys
import types

import numpy as np
import torch


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def install():
    """Install the fake `tensorflow` (+ `tensorflow.contrib.slim`) modules."""
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "_dfl_shim", False):
        return sys.modules["tensorflow"]
    tf = types.ModuleType("tensorflow")
    tf._dfl_shim = True
    tf.float32 = torch.float32
    tf.concat = lambda values, axis=0, name=None: torch.cat([_t(v) for v in values], dim=axis)
    tf.expand_dims = lambda x, axis=None, name=None: torch.unsqueeze(_t(x), axis)
    tf.stack = lambda values, axis=0, name=None: torch.stack([_t(v) for v in values], dim=axis)
    tf.transpose = lambda x, perm=None: _t(x).permute(*perm)
    tf.maximum = lambda a, b: torch.maximum(_t(a), _t(b))
    tf.reshape = lambda x, shape: _t(x).reshape(*shape)
    tf.abs = torch.abs
    tf.reduce_mean = lambda x, axis=None: torch.mean(x) if axis is None else torch.mean(x, dim=axis)

    image = types.ModuleType("tensorflow.image")

    def resize_nearest_neighbor(x, new_size):
        # tf.image.resize_nearest_neighbor, align_corners=False: out[i] = in[floor(i*in/out)]
        x = _t(x)
        b, h, w, c = x.shape
        nh, nw = int(new_size[0]), int(new_size[1])
        iy = torch.div(torch.arange(nh) * h, nh, rounding_mode="floor")
        ix = torch.div(torch.arange(nw) * w, nw, rounding_mode="floor")
        return x[:, iy][:, :, ix]

    image.resize_nearest_neighbor = resize_nearest_neighbor
    tf.image = image

    class _Shape(object):
        def __init__(self, s):
            self._s = list(s)

        def as_list(self):
            return list(self._s)

        ndims = property(lambda self: len(self._s))

    # ops.int_shape (ops.py:96-98) calls tensor.get_shape().as_list()
    if not hasattr(torch.Tensor, "get_shape"):
        torch.Tensor.get_shape = lambda self: _Shape(self.shape)

    contrib = types.ModuleType("tensorflow.contrib")
    slim = types.ModuleType("tensorflow.contrib.slim")
    contrib.slim = slim
    tf.contrib = contrib
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.image"] = image
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.slim"] = slim
    return tf


def import_reference_ops(reference_root="/root/reference"):
    """Import the reference's ops.py verbatim (read-only) under the shim."""
    import importlib.util

    install()
    spec = importlib.util.spec_from_file_location("_dfl_reference_ops", reference_root + "/ops.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
