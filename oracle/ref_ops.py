"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the deep-fluids hot path (ops level).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this package.  The product path
(`deepfluids_b200`) never does; it fails loudly when its CUDA library is missing.

A torch-CPU restatement (fp32 or fp64) of what the reference computes on the
path BASELINE.json's north_star names.  All tensors are channels-last
(NHWC `[B,H,W,C]` / NDHWC `[B,D,H,W,C]`, axis W=x, H=y, D=z; channel 0/1/2 =
u/v/w), exactly as in the reference (ops.py:228 "x: bzyxd").

Pinning status
--------------
* stencils (curl / jacobian / jacobian3 / divergence / divergence3 / lrelu /
  upscale / upscale3): PINNED -- `oracle/make_golden.py` executes the
  reference's own ops.py source verbatim through `oracle/tf_shim.py` and the
  restatement below is checked bit-exactly against it (and against the
  reference's numpy twins ops.py:305-374); the outputs are committed under
  `tests/golden/`.
* conv2d/conv3d/linear (slim -> TF library kernels, ops.py:12-24) and the Adam
  update (tf.train.AdamOptimizer, trainer.py:160-162): PARITY UNPINNED -- the
  arithmetic lives in TensorFlow 1.15 (README.md:18,23), which is not vendored
  and cannot be installed here, and the reference holds no golden vectors or
  tests for them.  They are restated from TF's published
  semantics (SAME padding, cross-correlation, HWIO/DHWIO weights, Adam with
  un-corrected epsilon) and cross-checked against torch's own conv/linear.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# a1  lrelu  (reference ops.py:9-10: tf.maximum(x, leak*x))
# ----------------------------------------------------------------------------------------------
def lrelu(x, leak=0.2):
    return torch.maximum(x, leak * x)


# ----------------------------------------------------------------------------------------------
# a9/a8  finite-difference stencils (reference ops.py:205-290)
# ----------------------------------------------------------------------------------------------
def fdiff(f, axis):
    """Forward difference f[i+1]-f[i] along `axis`, last entry replicated.

    This is the single building block of every reference stencil: e.g.
    ops.py:208 `dudx = x[:,:,1:,0]-x[:,:,:-1,0]` followed by ops.py:213
    `concat([dudx, dudx[:,:,-1:]])`.  Needs size >= 2 along `axis`.
    """
    n = f.shape[axis]
    d = f.narrow(axis, 1, n - 1) - f.narrow(axis, 0, n - 1)
    return torch.cat([d, d.narrow(axis, n - 2, 1)], dim=axis)


def jacobian(x):
    """2D velocity gradient. x [B,H,W,2] -> (j [B,H,W,4]=[dudx,dudy,dvdx,dvdy], w [B,H,W,1]=dvdx-dudy).

    Reference ops.py:205-225 (NHWC branch)."""
    u, v = x[..., 0], x[..., 1]
    dudx, dudy = fdiff(u, 2), fdiff(u, 1)
    dvdx, dvdy = fdiff(v, 2), fdiff(v, 1)
    j = torch.stack([dudx, dudy, dvdx, dvdy], dim=-1)
    w = (dvdx - dudy).unsqueeze(-1)
    return j, w


def jacobian3(x):
    """3D velocity gradient + curl.  x [B,D,H,W,3] ->
    (j [B,D,H,W,9] = [dudx,dudy,dudz,dvdx,dvdy,dvdz,dwdx,dwdy,dwdz],
     c [B,D,H,W,3] = [dwdy-dvdz, dudz-dwdx, dvdx-dudy]).   Reference ops.py:227-262."""
    comps = []
    for ch in range(3):
        f = x[..., ch]
        comps += [fdiff(f, 3), fdiff(f, 2), fdiff(f, 1)]  # d/dx (W), d/dy (H), d/dz (D)
    dudx, dudy, dudz, dvdx, dvdy, dvdz, dwdx, dwdy, dwdz = comps
    j = torch.stack(comps, dim=-1)
    c = torch.stack([dwdy - dvdz, dudz - dwdx, dvdx - dudy], dim=-1)
    return j, c


def curl(x):
    """2D curl of a stream function. x [B,H,W,>=1] (channel 0 used) -> [B,H,W,2] =
    (d psi/dy, -d psi/dx).   Reference ops.py:264-274."""
    s = x[..., 0]
    return torch.stack([fdiff(s, 1), -fdiff(s, 2)], dim=-1)


def curl3(x):
    """3D curl of a vector potential = second return of jacobian3 (trainer3.py:18)."""
    return jacobian3(x)[1]


def divergence(x):
    """[B,H,W,2] -> [B,H-1,W-1,1].   Reference ops.py:276-284."""
    dudx = x[:, :-1, 1:, 0] - x[:, :-1, :-1, 0]
    dvdy = x[:, 1:, :-1, 1] - x[:, :-1, :-1, 1]
    return (dudx + dvdy).unsqueeze(-1)


def divergence3(x):
    """[B,D,H,W,3] -> [B,D-1,H-1,W-1,1].   Reference ops.py:286-290."""
    c = x[:, :-1, :-1, :-1]
    dudx = x[:, :-1, :-1, 1:, 0] - c[..., 0]
    dvdy = x[:, :-1, 1:, :-1, 1] - c[..., 1]
    dwdz = x[:, 1:, :-1, :-1, 2] - c[..., 2]
    return (dudx + dvdy + dwdz).unsqueeze(-1)


# ----------------------------------------------------------------------------------------------
# a4  nearest-neighbour x2 upsampling (reference ops.py:66-91; align_corners=False => out[i]=in[i>>1])
# ----------------------------------------------------------------------------------------------
def upscale(x, scale=2):
    """[B,H,W,C] -> [B,H*s,W*s,C]."""
    return x.repeat_interleave(scale, dim=1).repeat_interleave(scale, dim=2)


def upscale3(x, scale=2):
    """[B,D,H,W,C] -> [B,D*s,H*s,W*s,C] (the reference's two-pass transpose/resize, ops.py:79-91,
    is value-identical to replicating every axis)."""
    return (x.repeat_interleave(scale, dim=1).repeat_interleave(scale, dim=2)
             .repeat_interleave(scale, dim=3))


# ----------------------------------------------------------------------------------------------
# a2/a3  linear / conv (reference ops.py:12-24 -> slim.fully_connected / slim.conv2d / slim.conv3d)
# ----------------------------------------------------------------------------------------------
def linear(x, weights, biases):
    """slim.fully_connected with activation_fn=None: y = x @ W + b, W [in,out] (TF layout)."""
    return x @ weights + biases


def _same_pads(size, k, s):
    """TF 'SAME' padding along one axis: returns (pad_before, pad_after)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv_nd(x, weights, biases, stride=1, act=None):
    """slim.conv2d / slim.conv3d: cross-correlation, padding='SAME', bias add, optional activation.

    x [B,(D,)H,W,Cin] channels-last; weights TF layout [k,(k,)k,Cin,Cout]; stride-2 k=3 on an even
    size pads 0 before / 1 after (TF SAME), unlike torch's symmetric padding."""
    nd = x.dim() - 2
    k = weights.shape[0]
    if nd == 2:
        xt = x.permute(0, 3, 1, 2)
        wt = weights.permute(3, 2, 0, 1)
        pads = []
        for ax in (2, 1):  # F.pad wants last axis first: W then H
            pb, pa = _same_pads(x.shape[ax], k, stride)
            pads += [pb, pa]
        y = F.conv2d(F.pad(xt, pads), wt, biases, stride=stride)
        y = y.permute(0, 2, 3, 1)
    else:
        xt = x.permute(0, 4, 1, 2, 3)
        wt = weights.permute(4, 3, 0, 1, 2)
        pads = []
        for ax in (3, 2, 1):
            pb, pa = _same_pads(x.shape[ax], k, stride)
            pads += [pb, pa]
        y = F.conv3d(F.pad(xt, pads), wt, biases, stride=stride)
        y = y.permute(0, 2, 3, 4, 1)
    return act(y) if act is not None else y


def conv_nd_direct(x, weights, biases):
    """Independent second witness for conv_nd (stride 1, k odd): explicit shift-and-matmul loops, no
    library convolution.  Small cases only."""
    nd = x.dim() - 2
    k = weights.shape[0]
    r = k // 2
    cout = weights.shape[-1]
    y = torch.zeros(x.shape[:-1] + (cout,), dtype=x.dtype) + biases
    sp = x.shape[1:-1]
    xp = F.pad(x, [0, 0] + [r, r] * nd)
    if nd == 2:
        for a in range(k):
            for b in range(k):
                y = y + xp[:, a:a + sp[0], b:b + sp[1], :] @ weights[a, b]
    else:
        for a in range(k):
            for b in range(k):
                for c in range(k):
                    y = y + xp[:, a:a + sp[0], b:b + sp[1], c:c + sp[2], :] @ weights[a, b, c]
    return y


def xavier_uniform_(shape, generator, dtype=torch.float32):
    """slim's default weights_initializer = xavier_initializer(uniform=True):
    U(-l, l), l = sqrt(6/(fan_in+fan_out)); conv fan = receptive field * channels."""
    rf = 1
    for s in shape[:-2]:
        rf *= s
    fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=generator, dtype=torch.float64) * 2 - 1).mul_(lim).to(dtype)
