"""Mirror of the reference's ops.py surface on channels-last torch CUDA tensors, backed by the C-ABI kernels.

Same names / argument meaning as reference ops.py: lrelu (:9-10), conv2d/conv3d (:12-16), linear (:23-24), upscale /
upscale3 (:75-91), jacobian / jacobian3 (:205-262), curl (:264-274), divergence / divergence3 (:276-290).  The train
step itself does not go through these one-op-at-a-time wrappers (it uses engine.GeneratorEngine, which fuses them);
they exist so reference call sites keep working and so each kernel is individually testable.
"""
import torch

from . import kernels as K


def lrelu(x, leak=0.2):
    assert leak == 0.2, "the fused kernels implement the reference's leak=0.2"
    return torch.maximum(x, leak * x)        # standalone use only; in the train step lrelu is a conv epilogue


def curl(x, data_format='NHWC'):
    """2D: [B,H,W,>=1] -> [B,H,W,2] (ops.py:264-274).  3D input -> second return of jacobian3 (trainer3.py:18)."""
    assert data_format == 'NHWC'
    return K.curl_fwd(x.contiguous())


def jacobian(x, data_format='NHWC'):
    assert data_format in ('NHWC', 'NHCW')   # the reference's default literal is the typo 'NHCW' (ops.py:205)
    return K.jacobian_fwd(x.contiguous())


def jacobian3(x):
    return K.jacobian_fwd(x.contiguous())


def divergence(x, data_format='NHWC'):
    return K.divergence(x.contiguous())


def divergence3(x):
    return K.divergence(x.contiguous())


def linear(x, weights, biases, out_dtype=torch.bfloat16):
    """slim.fully_connected(activation_fn=None) with explicit variables (TF layout [in,out])."""
    return K.fc_fwd(x.contiguous().float(), weights, biases, out_dtype=out_dtype)


def _conv(x, weights, biases, act):
    w_fwd, _ = K.pack_conv_weights(weights.contiguous())
    out = torch.empty(x.shape[:-1] + (weights.shape[-1],), dtype=torch.bfloat16, device=x.device)
    K.conv3x3(x.contiguous(), w_fwd, biases, out=out, flags=K.CONV_LRELU if act is lrelu else 0)
    assert act in (None, lrelu)
    return out


def conv2d(x, weights, biases, k=3, s=1, act=None):
    """slim.conv2d, SAME, k=3, s=1, 128->128 (ops.py:12-13) with explicit TF-layout variables."""
    assert k == 3 and s == 1
    return _conv(x, weights, biases, act)


def conv3d(x, weights, biases, k=3, s=1, act=None):
    assert k == 3 and s == 1
    return _conv(x, weights, biases, act)


def upscale(x, scale, data_format='NHWC'):
    assert scale == 2
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)   # standalone only; fused in the conv epilogue


def upscale3(x, scale):
    assert scale == 2
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
