"""EncoderBE / EncoderBE3 (reference model.py:118-188) and AE / AE3 (model.py:190-216) on the B200 kernels.

Encoder structure (model.py:129-150 / :165-186), filters F = 128:
    x  = conv(x_in, F, s1) + lrelu ;  x0 = x                                     "E0"  (Cin = 2/3)
    for idx in range(repeat_num):
        num_conv x [ x = conv(x, F, s1) + lrelu ]         first one sees ch_in = F*(idx+1) channels
        x = concat([x, x0])                               ch = F*(idx+2)
        if idx < repeat_num-1:  x = conv(x, ch, s2) + lrelu ;  x0 = x
    z = fc(flatten(x), z_num)

Layout: every tensor wider than 128 channels is a channel-blocked buffer [nblk*B,(D,)H,W,128]; the concat buffer of
level idx holds block 0 = x (written by the level's last conv) and blocks 1.. = x0 (written by E0 / the stride-2 conv
of the level below), so the concat costs nothing.  Stride-2 convs run on the generic per-tap tensor-core kernel (TMA
element stride 2 forward and wgrad; one launch per output parity class for the data gradient).
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import kernels as K
from .engine import FlatParams, GeneratorEngine, _repeat_num, xavier_uniform


def _same_pad_before(n, k=3, s=2):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2


class EncoderEngine(object):
    def __init__(self, batch, x_shape, filters=128, z_num=16, num_conv=3, repeat=0, name="enc", device=None, seed=123,
                 params=None):
        assert filters == 128 and num_conv >= 1, "encoder needs filters=128 and at least one conv per level"
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.B, self.name, self.F, self.nc, self.z_num = int(batch), name, filters, int(num_conv), int(z_num)
        self.spatial = [int(s) for s in x_shape[:-1]]
        self.cin = int(x_shape[-1])
        self.nd = len(self.spatial)
        self.rep = _repeat_num(self.spatial, repeat)
        self.taps = 3 ** self.nd
        self.res = [[s // 2 ** i for s in self.spatial] for i in range(self.rep)]
        for r in self.res[:-1]:
            assert all(v % 2 == 0 for v in r), "stride-2 levels need even extents"
        self.nblk = [i + 2 for i in range(self.rep)]            # channel blocks of the concat buffer of level i
        # ---- variables (TF names, model.py:129,137,141,149)
        tab = OrderedDict()
        kshape = (3,) * self.nd
        n = 0
        self.n_e0 = "%s/%d_conv" % (name, n)
        tab[self.n_e0 + "/weights"] = kshape + (self.cin, filters)
        tab[self.n_e0 + "/biases"] = (filters,)
        n += 1
        self.n_conv, self.n_s2 = [], []
        for idx in range(self.rep):
            row = []
            cin = filters * (idx + 1)
            for _ in range(self.nc):
                nm = "%s/%d_conv" % (name, n)
                tab[nm + "/weights"] = kshape + (cin, filters)
                tab[nm + "/biases"] = (filters,)
                row.append(nm)
                cin = filters
                n += 1
            self.n_conv.append(row)
            if idx < self.rep - 1:
                ch = filters * (idx + 2)
                nm = "%s/%d_conv" % (name, n)
                tab[nm + "/weights"] = kshape + (ch, ch)
                tab[nm + "/biases"] = (ch,)
                self.n_s2.append(nm)
                n += 1
        self.n_fc = "%s/%d_fc" % (name, n)
        self.flat = int(np.prod(self.res[-1])) * filters * self.nblk[-1]
        tab[self.n_fc + "/weights"] = (self.flat, self.z_num)
        tab[self.n_fc + "/biases"] = (self.z_num,)
        self.table = tab
        if params is None:
            params = FlatParams(tab, self.device)
            g = torch.Generator().manual_seed(seed)
            for k, shp in tab.items():
                if k.endswith("weights"):
                    params.p(k).copy_(xavier_uniform(tuple(shp), g, self.device))
        self.params = params
        self.variables = list(tab.keys())
        # ---- bf16 operands
        bf = dict(dtype=torch.bfloat16, device=self.device)
        self.wf, self.wd = {}, {}
        self.wf[self.n_e0] = torch.zeros(filters, self.taps * 128, **bf)      # Cin zero-padded to 128
        for idx in range(self.rep):
            for c, nm in enumerate(self.n_conv[idx]):
                cin = filters * (idx + 1) if c == 0 else filters
                self.wf[nm] = torch.empty(filters, self.taps * cin, **bf)
                self.wd[nm] = torch.empty(cin, self.taps * filters, **bf)
            if idx < self.rep - 1:
                ch = filters * (idx + 2)
                self.wf[self.n_s2[idx]] = torch.empty(ch, self.taps * ch, **bf)
                self.wd[self.n_s2[idx]] = torch.empty(ch, self.taps * ch, **bf)
        self.repack()
        # ---- activations
        B = self.B
        self.xpad = torch.empty([B] + self.res[0] + [128], **bf)
        self.C = [torch.empty([self.nblk[i] * B] + self.res[i] + [128], **bf) for i in range(self.rep)]
        self.ylev = [[torch.empty([B] + self.res[i] + [128], **bf) for _ in range(self.nc - 1)] for i in range(self.rep)]
        self.z = torch.empty(B, self.z_num, dtype=torch.float32, device=self.device)
        # ---- gradients
        self.G = [torch.empty_like(c) for c in self.C]
        self.DS = [torch.empty([(self.nblk[i] - 1) * B] + self.res[i] + [128], **bf) for i in range(self.rep)]
        self._dp = [torch.empty([B] + self.res[0] + [128], **bf) for _ in range(2)]
        self._dw_e0 = torch.zeros(self.taps, 128, 128, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ helpers
    def _blk(self, t, i, n=1):
        return t[i * self.B:(i + n) * self.B]

    def _dpv(self, k, idx):
        shp = [self.B] + self.res[idx] + [128]
        n = int(np.prod(shp))
        return self._dp[k].view(-1)[:n].view(shp)

    def repack(self):
        P = self.params
        K.pack_conv_weights_ex(P.p(self.n_e0 + "/weights"), self.wf[self.n_e0], None, 128)
        for idx in range(self.rep):
            for nm in self.n_conv[idx]:
                K.pack_conv_weights(P.p(nm + "/weights"), self.wf[nm], self.wd[nm])
            if idx < self.rep - 1:
                nm = self.n_s2[idx]
                K.pack_conv_weights(P.p(nm + "/weights"), self.wf[nm], self.wd[nm])

    def _s2_geometry(self, idx):
        """taps of the stride-2 conv from level idx to idx+1 (TF SAME padding)"""
        fine = self.res[idx]
        pb = [_same_pad_before(n) for n in fine]
        return fine, self.res[idx + 1], pb

    # ------------------------------------------------------------------ forward (model.py:129-150 / :165-186)
    def forward(self, x):
        """x: fp32 [B,(D,)H,W,Cin] -> z fp32 [B, z_num]"""
        P, B, F = self.params, self.B, self.F
        L = K.CONV_LRELU
        K.pad_cast(x.contiguous(), self.xpad)
        K.conv3x3(self.xpad, self.wf[self.n_e0], P.p(self.n_e0 + "/biases"), out=self._blk(self.C[0], 1), flags=L)
        for idx in range(self.rep):
            nb_in = self.nblk[idx] - 1
            cur, cur_blk = self.C[idx][B:], nb_in
            for c, nm in enumerate(self.n_conv[idx]):
                dst = self.ylev[idx][c] if c < self.nc - 1 else self._blk(self.C[idx], 0)
                K.conv3x3(cur, self.wf[nm], P.p(nm + "/biases"), out=dst, flags=L, nblk=cur_blk)
                cur, cur_blk = dst, 1
            if idx < self.rep - 1:
                nm = self.n_s2[idx]
                ch = F * self.nblk[idx]
                fine, coarse, pb = self._s2_geometry(idx)
                taps = []
                for t in range(self.taps):
                    tt = [(t // 3 ** (self.nd - 1 - a)) % 3 for a in range(self.nd)]       # (tz,)ty,tx
                    off = [tt[a] - pb[a] for a in range(self.nd)]
                    taps.append(([0] * (3 - self.nd) + off) + [t * ch])
                bias = P.p(nm + "/biases")
                for ob in range(self.nblk[idx]):
                    K.conv_taps(self.C[idx], self.wf[nm][ob * 128:(ob + 1) * 128], bias[ob * 128:(ob + 1) * 128],
                                self._blk(self.C[idx + 1], 1 + ob), None, None, None, [B] + coarse, coarse, ch, 2, taps,
                                1, [0] * self.nd, flags=L)
        K.enc_fc_fwd(self.C[-1], P.p(self.n_fc + "/weights"), P.p(self.n_fc + "/biases"), self.z, self.nblk[-1])
        return self.z

    # ------------------------------------------------------------------ backward
    def backward(self, dz):
        """dz fp32 [B, z_num]; accumulates into params.grad (zero it first)"""
        P, B, F, nd = self.params, self.B, self.F, self.nd
        top = self.rep - 1
        K.enc_fc_bwd(self.C[top], P.p(self.n_fc + "/weights"), dz, P.g(self.n_fc + "/weights"), P.g(self.n_fc + "/biases"),
                     self.G[top], self.nblk[top])
        for idx in range(top, -1, -1):
            nb_in = self.nblk[idx] - 1
            G, Cc = self.G[idx], self.C[idx]
            dpre = self._dpv(0, idx)
            other = self._dpv(1, idx)
            K.add_mask(self._blk(G, 0), None, self._blk(Cc, 0), dpre)        # grad of the level's last conv (pre-act)
            for c in range(self.nc - 1, 0, -1):
                nm = self.n_conv[idx][c]
                xin = self.ylev[idx][c - 1]
                K.conv3x3_wgrad(xin, dpre, P.g(nm + "/weights"), P.g(nm + "/biases"))
                K.conv3x3(dpre, self.wd[nm], None, out=other, mask_src=xin)
                dpre, other = other, dpre
            # first conv of the level: channel-blocked input C[idx][1:]
            nm = self.n_conv[idx][0]
            cin = F * nb_in
            gw = P.g(nm + "/weights").view(self.taps, cin, F)
            for ib in range(nb_in):
                K.conv_wgrad_ex(self._blk(Cc, 1 + ib), dpre, gw[0, ib * 128:], P.g(nm + "/biases") if ib == 0 else None,
                                1, 1, cin * F, F)
                # dL/d(pre-activation of the producer of x0 block ib) = (dgrad + G[1+ib]) * lrelu'(x0 block)
                K.conv3x3(dpre, self.wd[nm][ib * 128:(ib + 1) * 128], None, out2=self._blk(self.DS[idx], ib),
                          residual=self._blk(G, 1 + ib), mask_src=self._blk(Cc, 1 + ib),
                          flags=K.CONV_MASK_AFTER_RESIDUAL)
            if idx > 0:
                nm = self.n_s2[idx - 1]
                ch = F * nb_in                                   # channels of C[idx-1] == outputs of the s2 conv
                fine, coarse, pb = self._s2_geometry(idx - 1)
                gw = P.g(nm + "/weights").view(self.taps, ch, ch)
                gb = P.g(nm + "/biases")
                Cf = self.C[idx - 1]
                for ib in range(nb_in):
                    for ob in range(nb_in):
                        K.conv_wgrad_ex(self._blk(Cf, ib), self._blk(self.DS[idx], ob), gw[0, ib * 128:, ob * 128:],
                                        gb[ob * 128:(ob + 1) * 128] if ib == 0 else None, 2, pb[0], ch * ch, ch)
                # data gradient: one launch per (input-channel block, parity class of the fine grid)
                for r in range(2 ** nd):
                    rr = [(r >> (nd - 1 - a)) & 1 for a in range(nd)]
                    taps = []
                    for t in range(self.taps):
                        tt = [(t // 3 ** (nd - 1 - a)) % 3 for a in range(nd)]
                        if any((rr[a] - tt[a] + pb[a]) % 2 for a in range(nd)):
                            continue
                        off = [(rr[a] - tt[a] + pb[a]) // 2 for a in range(nd)]
                        taps.append(([0] * (3 - nd) + off) + [(self.taps - 1 - t) * ch])
                    if not taps:
                        continue
                    for ib in range(nb_in):
                        K.conv_taps(self.DS[idx], self.wd[nm][ib * 128:(ib + 1) * 128], None, self._blk(self.G[idx - 1], ib),
                                    None, None, None, [B] + coarse, fine, ch, 1, taps, 2, rr)
            else:
                # E0: the padded-input weight gradient lands in a 128x128 scratch; rows >= Cin are discarded
                self._dw_e0.zero_()
                K.conv3x3_wgrad(self.xpad, self._blk(self.DS[0], 0), self._dw_e0, P.g(self.n_e0 + "/biases"))
                P.g(self.n_e0 + "/weights").view(self.taps, self.cin, F).add_(self._dw_e0[:, :self.cin, :])


class AEEngine(object):
    """AE / AE3 (model.py:190-216): z = Enc(x, num_conv-1);  out = Gen(z, x.shape, num_conv);  one flat parameter buffer
    with the reference's variable names (`AE/enc/...`, `AE/dec/...`)."""

    def __init__(self, batch, x_shape, filters=128, z_num=16, num_conv=4, repeat=0, name="AE", device=None, seed=123,
                 use_sparse=False, params=None):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.use_sparse = bool(use_sparse)
        # build both tables first so they can share ONE flat buffer (one Adam launch / one all-reduce)
        self.enc = EncoderEngine(batch, x_shape, filters, z_num, num_conv - 1, repeat, name + "/enc", self.device, seed)
        self.dec = GeneratorEngine(batch, list(x_shape), z_dim=z_num, filters=filters, num_conv=num_conv, repeat=repeat,
                                   name=name + "/dec", device=self.device, seed=seed + 1)
        tab = OrderedDict()
        tab.update(self.enc.table)
        tab.update(self.dec.params.table)
        if params is not None:        # reuse=True: the variables of an existing AE (same names and shapes)
            missing = [k for k in tab if k not in params.table or tuple(params.table[k]) != tuple(tab[k])]
            if missing:
                raise ValueError("reuse: variable %s does not exist with this shape" % missing[0])
            flat = params
        else:
            flat = FlatParams(tab, self.device)
            for k in self.enc.table:
                flat.p(k).copy_(self.enc.params.p(k))
            for k in self.dec.params.table:
                flat.p(k).copy_(self.dec.params.p(k))
        self.params = flat
        self.enc.params = flat
        self.dec.params = flat
        self.repack()
        self.variables = list(tab.keys())
        self.z_num = z_num
        self.dz = torch.zeros(batch, z_num, dtype=torch.float32, device=self.device)
        self.z_sig = torch.zeros_like(self.dz)
        self.dz_lin = torch.zeros_like(self.dz)
        self.loss_kl = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.adam_t = 0

    def repack(self):
        self.enc.repack()
        self.dec.repack()

    def forward(self, x):
        z = self.enc.forward(x)
        if self.use_sparse:                       # model.py:196 / :210
            K.ae_sigmoid(z, self.z_sig)
            z = self.z_sig
        pot = self.dec.forward(z)
        return pot, z

    def backward(self, dpot, p_num=0, sparsity=0.01, w5=1.0, fused=None):
        """dz must already hold d(loss_p)/dz (ae_loss_p); the decoder adds its FC input gradient, then the encoder runs.
        With use_sparse the Bernoulli-KL term (trainer.py:389-394) and the sigmoid derivative are applied in between.
        fused: see GeneratorEngine.backward (3D: loss + adjoint in the prologue of the output conv's backward)."""
        self.dec.backward(dpot, dz=self.dz, fused=fused)
        if self.use_sparse:
            K.ae_sparse_bwd(self.z_sig, self.dz, self.dz_lin, self.loss_kl, p_num, sparsity, w5)
            self.enc.backward(self.dz_lin)
        else:
            self.enc.backward(self.dz)

    def zero_grad(self):
        self.params.grad.zero_()

    def adam_lr_t(self, lr, beta1, beta2):
        self.adam_t += 1
        t = self.adam_t
        return lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)

    def optimizer_step_dev(self, lr_t_dev, adam, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0):
        """optimizer update with the step size read from device memory + operand repack (CUDA-graph capturable)"""
        P = self.params
        if adam:
            K.adam_step_dev(P.data, P.grad, P.m, P.v, lr_t_dev, beta1, beta2, eps, grad_scale)
        else:
            K.adam_step_dev(P.data, P.grad, None, None, lr_t_dev, 0.0, 0.0, 0.0, grad_scale)
        self.repack()

    def optimizer_step(self, lr, adam=True, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0):
        P = self.params
        if adam:
            K.adam_step(P.data, P.grad, P.m, P.v, self.adam_lr_t(lr, beta1, beta2), beta1, beta2, eps, grad_scale)
        else:
            K.adam_step(P.data, P.grad, None, None, lr, 0.0, 0.0, 0.0, grad_scale)
        self.repack()
