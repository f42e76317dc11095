"""Differentiable layer ops assembled from the block kernels (C-ABI), for any channel count.

The engines (engine.py / encoder.py) plan the generator / AE train step as a fixed kernel sequence.  This module is the
general form behind the reference's ops-level surface (`ops.conv2d / conv3d / linear / curl / jacobian`, ops.py:12-24,
205-274), where a layer is a differentiable graph node with ANY width -- e.g. the patch discriminator of arch=dg
(model.py:89-116: 3|6 -> 64 -> 128 -> 256 (stride 2) -> 512 -> 1).  Channels are zero-padded to multiples of 128 and stored
as channel blocks [nblk*B,(D,)H,W,128] (the encoder's layout), so every layer runs on the same tcgen05 kernels:
    stride 1:  dfl_conv3x3_fwd per 128-channel output block (tap-window kernel);
    stride 2:  dfl_conv_taps (TMA element stride 2; TF SAME padding through the tap offsets);
    Cout<=3 :  dfl_lastconv_fwd / dfl_lastconv_bwd per 128-channel input block (fp32 output);
    backward:  dfl_add_mask (lrelu'), dfl_conv_wgrad_ex per (input block, output block), dgrad = dfl_conv3x3_fwd with the
               flipped operand (stride 1) or one dfl_conv_taps launch per output parity class (stride 2).
torch is used for memory plumbing only (zero padding, block (un)packing, concatenation): no arithmetic runs through it.
"""
import torch

from . import kernels as K

BF = torch.bfloat16


def _cdiv(a, b):
    return -(-a // b)


def _same_pad_before(n, k=3, s=2):
    out = _cdiv(n, s)
    return max((out - 1) * s + k - n, 0) // 2


# ------------------------------------------------------------------ channel blocks
def to_blocks(x):
    """[B,(D,)H,W,C] fp32 / bf16 -> (bf16 channel blocks [nb*B,(D,)H,W,128] zero padded, nb)"""
    B, C = x.shape[0], x.shape[-1]
    nb = _cdiv(C, 128)
    if C <= 8 and x.dtype == torch.float32:              # dfl_pad_cast: the 2..6-channel network inputs
        out = torch.empty(x.shape[:-1] + (128,), dtype=BF, device=x.device)
        K.pad_cast(x.contiguous(), out)
        return out, 1
    xb = x.to(BF)
    if C != nb * 128:
        xb = torch.nn.functional.pad(xb, (0, nb * 128 - C))
    if nb == 1:
        return xb.contiguous(), 1
    xb = xb.reshape(x.shape[:-1] + (nb, 128)).movedim(-2, 0)          # [nb, B, ..., 128]
    return xb.reshape((nb * B,) + tuple(x.shape[1:-1]) + (128,)).contiguous(), nb


def from_blocks(xb, B, C, dtype=None):
    """inverse of to_blocks: [nb*B,...,128] -> [B,...,C]"""
    nb = xb.shape[0] // B
    if nb == 1:
        y = xb[..., :C]
    else:
        y = xb.reshape((nb, B) + tuple(xb.shape[1:])).movedim(0, -2).reshape((B,) + tuple(xb.shape[1:-1]) + (nb * 128,))[..., :C]
    y = y.contiguous()
    return y if dtype is None or y.dtype == dtype else y.to(dtype)


# ------------------------------------------------------------------ convolution layer (k = 3, SAME, stride 1 | 2)
class ConvPack(object):
    """bf16 tensor-core operands of one conv layer, built from the fp32 TF-layout variable [3,(3,)3,Cin,Cout]."""

    def __init__(self, w, b):
        self.nd = w.dim() - 2
        assert all(int(k) == 3 for k in w.shape[:-2]), "the tensor-core conv kernels are 3x3(x3) (got kernel %s)" % (tuple(w.shape[:-2]),)
        self.taps = 3 ** self.nd
        self.cin, self.cout = int(w.shape[-2]), int(w.shape[-1])
        self.nb_in, self.nb_out = _cdiv(self.cin, 128), _cdiv(self.cout, 128)
        self.small = self.cout <= 3                      # 128 -> 1..3 output-conv kernels
        self.w = w
        w3 = w.detach().reshape(self.taps, self.cin, self.cout).float()
        if self.small:
            wp = torch.zeros(self.taps, self.nb_in * 128, self.cout, dtype=torch.float32, device=w.device)
            wp[:, :self.cin] = w3
            self.w_blk = [wp[:, i * 128:(i + 1) * 128].contiguous() for i in range(self.nb_in)]
            self.bias = None if b is None else b.detach().float().contiguous()
        else:
            wp = torch.zeros(self.taps, self.nb_in * 128, self.nb_out * 128, dtype=torch.float32, device=w.device)
            wp[:, :self.cin, :self.cout] = w3
            self.wf, self.wd = K.pack_conv_weights(wp)   # [Cout_p, taps*Cin_p], [Cin_p, taps*Cout_p] (taps flipped)
            self.bias = torch.zeros(self.nb_out * 128, dtype=torch.float32, device=w.device)
            if b is not None:
                self.bias[:self.cout] = b.detach().float()


def _tap_digits(t, nd):
    return [(t // 3 ** (nd - 1 - a)) % 3 for a in range(nd)]


def conv_fwd(xb, B, pk, stride, lrelu):
    """xb: channel blocks of the input.  -> channel blocks of lrelu?(conv + bias) (bf16), or fp32 [B,...,Cout] if Cout <= 3"""
    nd = pk.nd
    fine = list(xb.shape[1:-1])
    assert xb.shape[0] == pk.nb_in * B and stride in (1, 2)
    if pk.small:
        assert stride == 1 and not lrelu, "the 128 -> 1..3 output conv is stride 1 without activation (model.py:42,84,98,113)"
        out = None
        for ib in range(pk.nb_in):
            o = K.lastconv_fwd(xb[ib * B:(ib + 1) * B], pk.w_blk[ib], pk.bias if ib == 0 else None)
            out = o if out is None else out.add_(o)
        return out
    flags = K.CONV_LRELU if lrelu else 0
    if stride == 1:
        yb = torch.empty([pk.nb_out * B] + fine + [128], dtype=BF, device=xb.device)
        for ob in range(pk.nb_out):
            K.conv3x3(xb, pk.wf[ob * 128:(ob + 1) * 128], pk.bias[ob * 128:(ob + 1) * 128], out=yb[ob * B:(ob + 1) * B],
                      flags=flags, nblk=pk.nb_in)
        return yb
    assert all(v % 2 == 0 for v in fine), "stride-2 layers need even extents (got %s)" % fine
    coarse = [v // 2 for v in fine]
    pb = [_same_pad_before(v) for v in fine]
    cin_p = pk.nb_in * 128
    taps = []
    for t in range(pk.taps):
        tt = _tap_digits(t, nd)
        taps.append([0] * (3 - nd) + [tt[a] - pb[a] for a in range(nd)] + [t * cin_p])
    yb = torch.empty([pk.nb_out * B] + coarse + [128], dtype=BF, device=xb.device)
    for ob in range(pk.nb_out):
        K.conv_taps(xb, pk.wf[ob * 128:(ob + 1) * 128], pk.bias[ob * 128:(ob + 1) * 128], yb[ob * B:(ob + 1) * B], None, None,
                    None, [B] + coarse, coarse, cin_p, 2, taps, 1, [0] * nd, flags=flags)
    return yb


def conv_bwd(xb, yb, gy, B, pk, stride, lrelu, want_w=True, want_x=True):
    """gy: upstream gradient, channel blocks (bf16) -- or fp32 [B,...,Cout] for the Cout <= 3 layer.
    -> (gw [3,(3,)3,Cin,Cout] fp32 | None, gb [Cout] | None, gxb channel blocks | None)"""
    nd = pk.nd
    dev = xb.device
    fine = list(xb.shape[1:-1])
    if pk.small:
        gy = gy.float().contiguous()
        gw = torch.zeros(pk.taps, pk.nb_in * 128, pk.cout, dtype=torch.float32, device=dev)
        gb = torch.zeros(pk.cout, dtype=torch.float32, device=dev)
        gxb = torch.empty_like(xb) if want_x else None
        for ib in range(pk.nb_in):
            dw = torch.zeros(pk.taps, 128, pk.cout, dtype=torch.float32, device=dev)
            db = gb if ib == 0 else torch.zeros_like(gb)
            K.lastconv_bwd(xb[ib * B:(ib + 1) * B], gy, pk.w_blk[ib], None, gxb[ib * B:(ib + 1) * B] if want_x else None, None, dw, db)
            gw[:, ib * 128:(ib + 1) * 128] = dw
        return (gw[:, :pk.cin].reshape(pk.w.shape) if want_w else None), (gb if want_w else None), gxb
    if lrelu:
        dpre = torch.empty_like(gy)
        K.add_mask(gy, None, yb, dpre)                   # dL/d(pre-activation) = gy * lrelu'(y)
    else:
        dpre = gy
    cin_p, cout_p = pk.nb_in * 128, pk.nb_out * 128
    pad = 1 if stride == 1 else _same_pad_before(fine[0])
    if stride == 2:
        assert len(set(_same_pad_before(v) for v in fine)) == 1
    gw = gb = None
    if want_w:
        gwp = torch.zeros(pk.taps, cin_p, cout_p, dtype=torch.float32, device=dev)
        gbp = torch.zeros(cout_p, dtype=torch.float32, device=dev)
        for ib in range(pk.nb_in):
            for ob in range(pk.nb_out):
                K.conv_wgrad_ex(xb[ib * B:(ib + 1) * B], dpre[ob * B:(ob + 1) * B], gwp[0, ib * 128:, ob * 128:],
                                gbp[ob * 128:(ob + 1) * 128] if ib == 0 else None, stride, pad, cin_p * cout_p, cout_p)
        gw = gwp[:, :pk.cin, :pk.cout].reshape(pk.w.shape)
        gb = gbp[:pk.cout]
    gxb = None
    if want_x:
        if stride == 1:
            gxb = torch.empty_like(xb)
            for ib in range(pk.nb_in):
                K.conv3x3(dpre, pk.wd[ib * 128:(ib + 1) * 128], None, out=gxb[ib * B:(ib + 1) * B], nblk=pk.nb_out)
        else:
            # dX[2p + r] = sum_{t == r (mod 2)} dY[p + (r - t + pad)/2] W[t]^T: one launch per parity class r of the fine grid
            gxb = torch.zeros_like(xb)
            coarse = [v // 2 for v in fine]
            pb = [_same_pad_before(v) for v in fine]
            for r in range(2 ** nd):
                rr = [(r >> (nd - 1 - a)) & 1 for a in range(nd)]
                taps = []
                for t in range(pk.taps):
                    tt = _tap_digits(t, nd)
                    if any((rr[a] - tt[a] + pb[a]) % 2 for a in range(nd)):
                        continue
                    taps.append([0] * (3 - nd) + [(rr[a] - tt[a] + pb[a]) // 2 for a in range(nd)] + [(pk.taps - 1 - t) * cout_p])
                if not taps:
                    continue
                for ib in range(pk.nb_in):
                    K.conv_taps(dpre, pk.wd[ib * 128:(ib + 1) * 128], None, gxb[ib * B:(ib + 1) * B], None, None, None,
                                [B] + coarse, fine, cout_p, 1, taps, 2, rr)
    return gw, gb, gxb


class _ConvFn(torch.autograd.Function):
    """y = act(conv(x, w) + b) on logical channels-last tensors (slim.conv2d / conv3d, ops.py:12-16)"""

    @staticmethod
    def forward(ctx, x, w, b, stride, lrelu):
        B = x.shape[0]
        pk = ConvPack(w, b)
        xb, _ = to_blocks(x)
        y = conv_fwd(xb, B, pk, stride, lrelu)
        ctx.pk, ctx.stride, ctx.lrelu, ctx.B, ctx.xdtype = pk, stride, lrelu, B, x.dtype
        if pk.small:
            ctx.save_for_backward(xb)
            return y
        ctx.save_for_backward(xb, y)
        return from_blocks(y, B, pk.cout)

    @staticmethod
    def backward(ctx, gy):
        pk, B = ctx.pk, ctx.B
        if pk.small:
            (xb,) = ctx.saved_tensors
            yb, g = None, gy
        else:
            xb, yb = ctx.saved_tensors
            g, _ = to_blocks(gy)
        want_x, want_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        gw, gb, gxb = conv_bwd(xb, yb, g, B, pk, ctx.stride, ctx.lrelu, want_w, want_x)
        gx = from_blocks(gxb, B, pk.cin, ctx.xdtype) if want_x else None
        return gx, gw, gb, None, None


def conv(x, w, b, stride=1, lrelu=False):
    return _ConvFn.apply(x, w, b, int(stride), bool(lrelu))


# ------------------------------------------------------------------ stencils as graph nodes (ops.py:205-274)
class _CurlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pot):
        ctx.cs = pot.shape[-1]
        return K.curl_fwd(pot.contiguous())

    @staticmethod
    def backward(ctx, gvel):
        return K.curl_bwd(gvel.float().contiguous(), ctx.cs)


class _JacobianFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vel):
        ctx.set_materialize_grads(False)          # an unused output (e.g. only the vorticity is consumed) arrives as None
        jac, aux = K.jacobian_fwd(vel.contiguous())
        return jac, aux

    @staticmethod
    def backward(ctx, gjac, gaux):
        if gjac is None and gaux is None:
            return None
        gj = None if gjac is None else gjac.float().contiguous()
        ga = None if gaux is None else gaux.float().contiguous()
        return K.jacobian_bwd(gj, ga)


def curl(pot):
    return _CurlFn.apply(pot)


def jacobian(vel):
    return _JacobianFn.apply(vel)


# ------------------------------------------------------------------ fully connected (slim.fully_connected, ops.py:23-24)
class _LinearFn(torch.autograd.Function):
    """y = x W + b, fp32 (dfl_gemm_f32; backward: dx = dy W^T, dW = x^T dy, db = column sums of dy)"""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.float().contiguous()
        ctx.save_for_backward(x, w)
        return K.gemm(x, w.detach().contiguous(), None if b is None else b.detach().contiguous())

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.float().contiguous()
        gw = K.gemm(x, gy, trans_a=True) if ctx.needs_input_grad[1] else None
        gb = K.colsum(gy) if ctx.needs_input_grad[2] else None
        gx = K.gemm(gy, w.detach().contiguous(), trans_b=True) if ctx.needs_input_grad[0] else None
        return gx, gw, gb


def linear(x, w, b):
    return _LinearFn.apply(x, w, b)
