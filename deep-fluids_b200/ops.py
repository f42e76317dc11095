"""Drop-in mirror of the reference's ops.py call surface on channels-last torch CUDA tensors (SURVEY.md 8b "Ops").

Same names, argument order, defaults and returns as reference ops.py:
    lrelu(x, leak)                                              ops.py:9-10
    conv2d / conv3d(x, o_dim, data_format, name, k, s, act)     ops.py:12-16   (slim.conv2d / conv3d, SAME padding)
    linear(x, o_dim, name, act)                                 ops.py:23-24   (slim.fully_connected)
    batch_norm(x, train, data_format, name, act, epsilon, momentum)  ops.py:26-36 (slim.batch_norm, [batch, features] only)
    upscale(x, scale, data_format) / upscale3(x, scale)         ops.py:75-91
    jacobian(x, data_format) -> (j, w) / jacobian3(x) -> (j, c) ops.py:205-262
    curl(x, data_format)                                        ops.py:264-274
    divergence(x, data_format) / divergence3(x)                 ops.py:276-290
Like their TF originals the layer ops CREATE their variables on first use -- `<scope>/<name>/weights|biases`, xavier-
uniform / zeros, TF layouts -- in a scope-keyed variable store (`variable_scope`, `get_variables`: the stand-ins for
tf.variable_scope / tf.contrib.framework.get_variables, incl. slim's default layer names `Conv`, `Conv_1`, ... and
`fully_connected`), and every op is DIFFERENTIABLE (torch.autograd.Function over the C-ABI's backward kernels, layers.py),
so a reference-style model function written against this module trains.  All arithmetic runs in the sm_100a library; the
train step of arch=de/ae itself does not go op by op through here (engine.py fuses it).
Restrictions (raised loudly, no fallback): k must be 3 and s in {1, 2} (every call site of the reference on this path
passes k=3; the reference's default k=4 exists for call sites it no longer has), act in {None, lrelu}, channels-last only.
"""
import contextlib
from collections import OrderedDict

import numpy as np
import torch

from . import kernels as K
from . import layers as L
from .engine import xavier_uniform


# ------------------------------------------------------------------ variable store (tf.variable_scope / get_variables)
class VariableStore(object):
    def __init__(self, seed=123, device=None):
        self.vars = OrderedDict()
        self.scopes = []                 # [(name, reuse)]
        self.default_counts = {}         # parent scope path -> {default layer name: uses}  (reset when the parent closes)
        self.generator = torch.Generator().manual_seed(seed)
        self.device = device

    def path(self):
        return "/".join(n for n, _ in self.scopes)

    def reuse(self):
        return any(r for _, r in self.scopes)

    def unique_default(self, default_name):
        cnt = self.default_counts.setdefault(self.path(), {})
        i = cnt.get(default_name, 0)
        cnt[default_name] = i + 1
        return default_name if i == 0 else "%s_%d" % (default_name, i)

    def get(self, name, shape, device, zeros=False, fill=None, trainable=True):
        full = self.path() + "/" + name if self.scopes else name
        v = self.vars.get(full)
        if v is not None:
            if not self.reuse():
                raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True in variable_scope?" % full)
            if tuple(v.shape) != tuple(shape):
                raise ValueError("Trying to share variable %s, but specified shape %s and found shape %s." % (full, tuple(shape), tuple(v.shape)))
            return v
        if self.reuse():
            raise ValueError("Variable %s does not exist, or was not created in this store (reuse=True)" % full)
        data = torch.zeros(shape, dtype=torch.float32, device=device) if zeros else xavier_uniform(tuple(shape), self.generator, device)
        if fill is not None:
            data.fill_(fill)
        v = torch.nn.Parameter(data, requires_grad=trainable)
        self.vars[full] = v
        return v


_STORE = VariableStore()


def reset_variables(seed=123):
    """start from an empty variable store (a fresh tf.Graph)"""
    global _STORE
    _STORE = VariableStore(seed)
    return _STORE


class _Scope(object):
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(name, reuse=False):
    _STORE.scopes.append((name, bool(reuse)))
    vs = _Scope(_STORE.path())
    try:
        yield vs
    finally:
        # tf closes the sub-scope counters when a scope exits, so re-entering it (reuse=True) yields Conv, Conv_1, ... again
        for k in [k for k in _STORE.default_counts if k == vs.name or k.startswith(vs.name + "/")]:
            del _STORE.default_counts[k]
        _STORE.scopes.pop()


def get_variables(scope=None):
    """names of the variables under `scope` (a variable_scope object or a name prefix), in creation order"""
    pre = scope.name if isinstance(scope, _Scope) else (scope or "")
    return [n for n in _STORE.vars if not pre or n == pre or n.startswith(pre + "/")]


def get_variable(name):
    return _STORE.vars[name]


# ------------------------------------------------------------------ activations / resampling
def lrelu(x, leak=0.2):
    if leak != 0.2:
        raise NotImplementedError("lrelu: the kernels implement the reference's only value leak=0.2 (ops.py:9)")
    return _LreluFn.apply(x)


class _LreluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        # standalone use only (in every layer of this path lrelu is the conv epilogue): y = x * lrelu'(x) via dfl_add_mask
        xb = x.to(torch.bfloat16).contiguous()
        assert xb.numel() % 8 == 0, "lrelu: element count must be a multiple of 8"
        y = torch.empty_like(xb)
        K.add_mask(xb, None, xb, y)
        ctx.save_for_backward(y)
        ctx.dtype = x.dtype
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        g = torch.empty_like(y)
        K.add_mask(gy.to(torch.bfloat16).contiguous(), None, y, g)
        return g.to(ctx.dtype)


def elu(x):
    """tf.nn.elu -- the activation of model.NN (model.py:218).  As a standalone op it runs through the batch-norm kernel's
    activation stage with an identity normalisation (gamma = 1, beta = 0, moving statistics 0 / 1 - eps)."""
    flat = x.reshape(-1, x.shape[-1]).float().contiguous()
    n = flat.shape[1]
    one, zero = torch.ones(n, device=x.device), torch.zeros(n, device=x.device)
    return _BnActFn.apply(flat, one, zero, zero.clone(), one.clone(), 0.0, 1.0, False, K.ACT_ELU).reshape(x.shape)


class _BnActFn(torch.autograd.Function):
    """slim.batch_norm (+ activation) on [M, N] fp32: dfl_bn_act_fwd / dfl_bn_act_bwd"""

    @staticmethod
    def forward(ctx, x, gamma, beta, mmean, mvar, eps, decay, training, act):
        x = x.float().contiguous()
        y, sm, sr = K.bn_act_fwd(x, gamma.detach(), beta.detach(), mmean, mvar, eps, decay, training, act)
        ctx.training, ctx.act = training, act
        if training:
            ctx.save_for_backward(x, y, gamma.detach(), sm, sr)
        else:   # inference statistics are constants: the layer is an affine map followed by the activation
            ctx.save_for_backward(x, y, gamma.detach(), mmean.clone(), torch.rsqrt(mvar + eps))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, sm, sr = ctx.saved_tensors
        dy = dy.float().contiguous()
        if ctx.training:
            dx, dg, db = K.bn_act_bwd(x, y, dy, gamma, sm, sr, ctx.act, want_dx=ctx.needs_input_grad[0])
            return dx, dg, db, None, None, None, None, None, None
        raise NotImplementedError("batch_norm(train=False) is not differentiated on this path (the reference only "
                                  "differentiates the training-mode graph)")


_DROPOUT = {"seed": 0x5EED, "offset": 0}


def dropout_seed(seed):
    _DROPOUT["seed"], _DROPOUT["offset"] = int(seed), 0


class _DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keep_prob):
        x = x.float().contiguous()
        ctx.keep, ctx.seed, ctx.offset = keep_prob, _DROPOUT["seed"], _DROPOUT["offset"]
        _DROPOUT["offset"] += x.numel()
        return K.dropout(x, keep_prob, ctx.seed, ctx.offset)

    @staticmethod
    def backward(ctx, dy):
        return K.dropout(dy.float().contiguous(), ctx.keep, ctx.seed, ctx.offset), None


def dropout(x, keep_prob, is_training=True):
    """slim.dropout(x, keep_prob, is_training): identity at inference; y = x * mask / keep_prob in training (counter-based
    mask, re-created in the backward pass).  NOTE model.NN passes its `dropout=0.1` argument as the KEEP probability."""
    if not is_training or keep_prob >= 1.0:
        return x
    return _DropoutFn.apply(x, float(keep_prob))


def batch_norm(x, train, data_format='NHWC', name=None, act=lrelu, epsilon=1e-5, momentum=0.9):
    """slim.batch_norm(decay=momentum, epsilon, scale=True, fused=True, updates_collections=None, is_training=train,
    activation_fn=act) (ops.py:26-36) for the [batch, features] tensors of model.NN.  Variables beta, gamma, moving_mean,
    moving_variance under `name` (default scope BatchNorm, BatchNorm_1, ...); the moving statistics are updated in place by
    every training-mode call, as updates_collections=None does."""
    if x.dim() != 2:
        raise NotImplementedError("batch_norm: [batch, features] tensors only (the conv stacks of this path have no batch norm)")
    code = {None: K.ACT_NONE, lrelu: K.ACT_LRELU, elu: K.ACT_ELU}.get(act)
    if code is None:
        raise NotImplementedError("batch_norm activation %r: None, ops.lrelu or ops.elu" % (act,))
    n = int(x.shape[1])
    scope = name if name is not None else _STORE.unique_default("BatchNorm")
    with variable_scope(scope):
        beta = _STORE.get("beta", (n,), x.device, zeros=True)
        gamma = _STORE.get("gamma", (n,), x.device, zeros=True, fill=1.0)
        mmean = _STORE.get("moving_mean", (n,), x.device, zeros=True, trainable=False)
        mvar = _STORE.get("moving_variance", (n,), x.device, zeros=True, fill=1.0, trainable=False)
    return _BnActFn.apply(x, gamma, beta, mmean.data, mvar.data, float(epsilon), float(momentum), bool(train), code)


class _UpscaleFn(torch.autograd.Function):
    """nearest x2 (dfl_upscale2) and its adjoint (dfl_pool2)"""

    @staticmethod
    def forward(ctx, x):
        return K.upscale2(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return K.pool2(g.contiguous())


class _AddFn(torch.autograd.Function):
    """residual add of two bf16 activations (dfl_add_mask without a mask)"""

    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty_like(a)
        K.add_mask(a.contiguous(), b.contiguous(), None, out)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    """x + x0 of the residual blocks (model.py:35,77) on the library when both are bf16 activations"""
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.shape == b.shape and a.numel() % 8 == 0 and a.is_cuda:
        return _AddFn.apply(a, b)
    return a + b


def upscale(x, scale, data_format='NHWC'):
    """nearest x2 (ops.py:75-77); standalone only -- in the generator it is the conv epilogue's replicated store"""
    if data_format != 'NHWC' or scale != 2:
        raise NotImplementedError("upscale: NHWC, scale 2 only")
    if x.dim() != 4:
        raise ValueError("upscale expects [B,H,W,C]")
    return _UpscaleFn.apply(x)


def upscale3(x, scale):
    if scale != 2:
        raise NotImplementedError("upscale3: scale 2 only")
    if x.dim() != 5:
        raise ValueError("upscale3 expects [B,D,H,W,C]")
    return _UpscaleFn.apply(x)


# ------------------------------------------------------------------ layers
def _act_flag(act):
    if act is None:
        return False
    if act is lrelu:
        return True
    raise NotImplementedError("activation %r: the conv epilogue implements None and ops.lrelu" % (act,))


def _conv(x, o_dim, data_format, name, k, s, act, nd):
    if data_format not in ('NHWC', 'NDHWC'):
        raise NotImplementedError("data_format %s: channels-last only (the reference's 3D path is NDHWC, ops.py:228)" % data_format)
    if k != 3 or s not in (1, 2):
        raise NotImplementedError("conv k=%d s=%d: the tensor-core kernels implement k=3, s in {1,2} (all call sites of the "
                                  "reference's de/ae/dg path)" % (k, s))
    if x.dim() != nd + 2:
        raise ValueError("conv%dd expects a [B,%sH,W,C] tensor" % (nd, "D," if nd == 3 else ""))
    lre = _act_flag(act)
    scope = name if name is not None else _STORE.unique_default("Conv")      # slim.convolution: variable_scope(scope, 'Conv')
    with variable_scope(scope):
        w = _STORE.get("weights", (3,) * nd + (int(x.shape[-1]), int(o_dim)), x.device)
        b = _STORE.get("biases", (int(o_dim),), x.device, zeros=True)
    return L.conv(x, w, b, s, lre)


def conv2d(x, o_dim, data_format='NHWC', name=None, k=4, s=2, act=None):
    return _conv(x, o_dim, data_format, name, k, s, act, 2)


def conv3d(x, o_dim, data_format='NDHWC', name=None, k=4, s=2, act=None):
    return _conv(x, o_dim, data_format, name, k, s, act, 3)


def linear(x, o_dim, name=None, act=None):
    if x.dim() != 2:
        raise ValueError("linear expects [B, in]")
    scope = name if name is not None else _STORE.unique_default("fully_connected")
    with variable_scope(scope):
        w = _STORE.get("weights", (int(x.shape[1]), int(o_dim)), x.device)
        b = _STORE.get("biases", (int(o_dim),), x.device, zeros=True)
    y = L.linear(x, w, b)
    return y if act is None else act(y)


# ------------------------------------------------------------------ finite-difference stencils
def _cl(x):
    return x.contiguous()


def curl(x, data_format='NHWC'):
    """2D: psi [B,H,W,>=1] -> [B,H,W,2] (ops.py:264-274); a 3D [B,D,H,W,3] input gives the curl jacobian3 returns second"""
    if data_format != 'NHWC':
        raise NotImplementedError("curl: channels-last only")
    return L.curl(_cl(x))


def jacobian(x, data_format='NHCW'):
    """-> (j [B,H,W,4], w [B,H,W,1]).  (The default literal 'NHCW' is the reference's own, ops.py:205: a typo for NHWC.)"""
    if data_format not in ('NHWC', 'NHCW'):
        raise NotImplementedError("jacobian: channels-last only")
    return L.jacobian(_cl(x))


def jacobian3(x):
    """-> (j [B,D,H,W,9], c [B,D,H,W,3])  (ops.py:227-262)"""
    return L.jacobian(_cl(x))


def divergence(x, data_format='NHWC'):
    return K.divergence(_cl(x))


def divergence3(x):
    return K.divergence(_cl(x))


# ------------------------------------------------------------------ numpy twins (ops.py:305-374)
# Host-side helpers the reference's test / plotting code calls on de-normalised numpy fields.  Written over ONE replicate-
# last difference helper; bit-identical to the reference's slice/concatenate formulation (same subtractions, same order).
def _fdiff_np(f, axis):
    d = np.diff(f, axis=axis)
    return np.concatenate([d, np.take(d, [-1], axis=axis)], axis=axis)


def vort_np(x):
    """[B,H,W,2] -> vorticity dv/dx - du/dy [B,H,W,1]  (ops.py:305-310)"""
    return (_fdiff_np(x[..., 1], 2) - _fdiff_np(x[..., 0], 1))[..., None]


def curl_np(x):
    """psi [B,H,W,>=1] -> (d psi/dy, -d psi/dx) [B,H,W,2]  (ops.py:312-317)"""
    psi = x[..., 0]
    return np.stack([_fdiff_np(psi, 1), _fdiff_np(-psi, 2)], axis=-1)


def grad_np(x):
    """p [B,H,W,>=1] -> (dp/dx, dp/dy) [B,H,W,2]  (ops.py:319-324)"""
    p = x[..., 0]
    return np.stack([_fdiff_np(p, 2), _fdiff_np(p, 1)], axis=-1)


def jacobian_np3(x):
    """[B,D,H,W,3] ("bzyxd") -> (j [..,9] = d(u,v,w)/d(x,y,z), c [..,3] = curl)  (ops.py:344-374)"""
    d = [[_fdiff_np(x[..., c], ax) for ax in (3, 2, 1)] for c in range(3)]          # d[c][a]: component c along x, y, z
    j = np.stack([d[c][a] for c in range(3) for a in range(3)], axis=-1)
    c = np.stack([d[2][1] - d[1][2], d[0][2] - d[2][0], d[1][0] - d[0][1]], axis=-1)
    return j, c
