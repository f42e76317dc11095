"""fp32-grade generator engine (BASELINE config 2: the reference computes in fp32, placeholders/vars tf.float32,
data.py:81-83) on the bf16 tensor cores via split operands ("bf16x3").

Every activation / gradient tensor is a (hi, lo) pair of bf16 tensors (hi = bf16(v), lo = bf16(v - hi): 16 bits of
mantissa) stored as two channel blocks [2*B,(D,)H,W,128]; weights are split the same way when the GEMM operands are
packed.  x*w ~ x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (dropped term 2^-18 relative) is a single launch of the tap-window
kernel over the virtual input blocks [hi, lo, hi] (3x the K of the bf16 path) with fp32 accumulation in TMEM; weight
gradients are one launch of the wgrad kernel that walks the three operand combinations per brick.  Same kernel sequence, same
fusions and the same flat fp32 parameter / Adam buffers as engine.GeneratorEngine.
"""
import numpy as np
import torch

from . import kernels as K
from .engine import GeneratorEngine


class GeneratorEngineFP32(GeneratorEngine):
    precision = "fp32x3"

    def __init__(self, *a, **kw):
        super(GeneratorEngineFP32, self).__init__(*a, **kw)

    # ---- buffers: called from the base constructor through the hooks below
    def _alloc(self):
        bf = dict(dtype=torch.bfloat16, device=self.device)
        B, F = self.B, self.filters
        self.x0, self.y = [], []
        for i in range(self.rep):
            shp = [2 * B] + self.level_shape[i] + [F]
            self.x0.append(torch.empty(shp, **bf))
            self.y.append([torch.empty(shp, **bf) for _ in range(self.num_conv)])
        top = [2 * B] + self.level_shape[-1] + [F]
        self.s = torch.empty(top, **bf)
        self.pot = torch.empty([B] + self.spatial + [self.cout], dtype=torch.float32, device=self.device)
        self._gbuf = [] if self.inference else [torch.empty(top, **bf) for _ in range(4)]
        self.dpot_pad = None if self.inference else torch.empty(top, **bf)
        n_fc = int(np.prod(self.level_shape[0])) * F
        self.x0f = torch.empty(B, n_fc, dtype=torch.float32, device=self.device)
        self.gx0f = torch.empty(B, n_fc, dtype=torch.float32, device=self.device)
        self._dw_last = torch.zeros(self.taps, F, F, dtype=torch.float32, device=self.device)
        self._db_last = torch.zeros(F, dtype=torch.float32, device=self.device)

    def _alloc_operands(self):
        bf = dict(dtype=torch.bfloat16, device=self.device)
        F = self.filters
        self.wf, self.wd = {}, {}
        for row in self.conv_names:
            for cn in row:
                self.wf[cn] = torch.empty(F, self.taps * 3 * F, **bf)
                self.wd[cn] = torch.empty(F, self.taps * 3 * F, **bf)
        self.w_last16 = torch.zeros(16, self.taps * 3 * F, **bf)      # forward operand of the output conv (rows >= C zero)
        self.wd_last = torch.zeros(F, self.taps * 3 * F, **bf)        # dgrad operand (columns co >= C zero)

    def repack(self):
        for row in self.conv_names:
            for cn in row:
                K.pack_conv_weights_split(self.params.p(cn + "/weights"), self.wf[cn], self.wd[cn])
        K.pack_conv_weights_split(self.params.p(self.last_name + "/weights"), self.w_last16, self.wd_last)

    def _gview(self, k, level):
        shp = [2 * self.B] + self.level_shape[level] + [self.filters]
        n = int(np.prod(shp))
        return self._gbuf[k].view(-1)[:n].view(shp)

    def _hi(self, t):
        return t[:self.B]

    def _lo(self, t):
        return t[self.B:]

    # ------------------------------------------------------------------ forward
    def forward(self, z):
        P, B = self.params, self.B
        self.z = z.contiguous().float()
        assert self.z.shape == (B, self.z_dim)
        for b0 in range(0, B, 64):
            K.fc_fwd(self.z[b0:b0 + 64], P.p(self.name + "/0_fc/weights"), P.p(self.name + "/0_fc/biases"),
                     out=self.x0f[b0:b0 + 64])
        K.split_f32(self.x0f.view(-1, self.filters), self.x0[0])
        L = K.CONV_LRELU
        for i in range(self.rep):
            cur = self.x0[i]
            for c in range(self.num_conv):
                cn = self.conv_names[i][c]
                bias = P.p(cn + "/biases")
                if c < self.num_conv - 1:
                    K.conv3x3_split(cur, self.wf[cn], bias, out=self.y[i][c], flags=L)
                elif i < self.rep - 1:
                    K.conv3x3_split(cur, self.wf[cn], bias, out=self.y[i][c], out2=self.x0[i + 1], residual=self.x0[i],
                                    flags=L | K.CONV_OUT2_UPSAMPLE)
                else:
                    K.conv3x3_split(cur, self.wf[cn], bias, out=self.y[i][c], out2=self.s, residual=self.x0[i], flags=L)
                cur = self.y[i][c]
        K.conv3x3_split(self.s, self.w_last16, P.p(self.last_name + "/biases"), out=self.pot, cout=self.cout)
        return self.pot

    # ------------------------------------------------------------------ backward
    def _wgrad3(self, x2, dp2, dw, db):
        # x_hi^T dP_hi + x_lo^T dP_hi + x_hi^T dP_lo accumulated in one TMEM tile, one reduction into dw
        K.conv3x3_wgrad_split(x2, dp2, dw, db)

    def backward(self, dpot, dz=None):
        assert not self.inference and self.B <= 64
        P, nc, top, C = self.params, self.num_conv, self.rep - 1, self.cout
        # output conv: its C-channel gradient is zero-padded to 128 channels so the 128-wide kernels apply
        K.split_f32(dpot.contiguous(), self.dpot_pad, cpad=128)
        ds = self._gview(0, top)
        dpre = self._gview(1, top)
        K.conv3x3_split(self.dpot_pad, self.wd_last, None, out=dpre, out2=ds, mask_src=self._hi(self.y[top][nc - 1]))
        self._dw_last.zero_()
        self._db_last.zero_()
        self._wgrad3(self.s, self.dpot_pad, self._dw_last, self._db_last)
        P.g(self.last_name + "/weights").view(self.taps, self.filters, C).add_(self._dw_last[:, :, :C])
        P.g(self.last_name + "/biases").add_(self._db_last[:C])
        for i in range(top, -1, -1):
            other = self._gview(2, i)
            gx0 = self._gview(3, i)
            for c in range(nc - 1, -1, -1):
                cn = self.conv_names[i][c]
                xin = self.y[i][c - 1] if c > 0 else self.x0[i]
                self._wgrad3(xin, dpre, P.g(cn + "/weights"), P.g(cn + "/biases"))
                if c > 0:
                    K.conv3x3_split(dpre, self.wd[cn], None, out=other, mask_src=self._hi(self.y[i][c - 1]))
                    dpre, other = other, dpre
                else:
                    K.conv3x3_split(dpre, self.wd[cn], None, out2=gx0, residual=ds)
            if i > 0:
                ds = self._gview(0, i - 1)
                dpre = self._gview(1, i - 1)
                K.pool_mask_split(gx0, self._hi(self.y[i - 1][nc - 1]), ds, dpre)
            else:
                K.merge_split(gx0, self.gx0f)
                K.fc_bwd(self.z, self.gx0f, P.g(self.name + "/0_fc/weights"), P.g(self.name + "/0_fc/biases"))
                if dz is not None:
                    K.fc_dz(self.gx0f, P.p(self.name + "/0_fc/weights"), dz, accumulate=True)
