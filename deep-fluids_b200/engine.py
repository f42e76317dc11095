"""Generator train-step engine: plans buffers once, then drives the C-ABI kernels in a fixed order.

This is the B200 replacement for what `sess.run(g_optim)` executes in the reference (trainer.py:269 /
trainer3.py:156): GeneratorBE/GeneratorBE3 forward (model.py:5-87), curl + Jacobian-L1 loss
(trainer.py:138-172 / trainer3.py:16-51), TF autodiff backward and AdamOptimizer.minimize (trainer.py:160-184).

Data layout in HBM (all channels-last, ops.py:228):
  * activations  bf16 [B,(D,)H,W,128]; every conv output y (post leaky-ReLU) is kept for the backward pass
    (it is the next layer's wgrad operand and its sign gives the leaky-ReLU derivative);
  * `x0` of block i+1 is written directly by the last conv of block i (fused residual add + nearest x2 upsample);
  * parameters   one flat fp32 buffer with TF-named, TF-laid-out views (`G/<n>_conv/weights` = [k,(k,)k,Cin,Cout]),
    matching flat fp32 gradient / Adam-m / Adam-v buffers (one fused Adam launch, one all-reduce);
  * GEMM operands bf16 re-packs of the conv weights (forward: [Cout, taps*Cin]; dgrad: [Cin, taps'*Cout]).
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import kernels as K


def _repeat_num(spatial, repeat):
    rep = int(np.log2(np.max(spatial))) - 2 if repeat == 0 else repeat   # model.py:9-12 / :51-54
    assert rep > 0 and sum(int(i) % (2 ** (rep - 1)) for i in spatial) == 0
    return rep


def xavier_uniform(shape, generator, device):
    """slim default initializer (xavier_initializer, uniform): U(-l,l), l = sqrt(6/(fan_in+fan_out))."""
    rf = 1
    for s in shape[:-2]:
        rf *= s
    lim = math.sqrt(6.0 / (rf * shape[-2] + rf * shape[-1]))
    return (torch.rand(shape, generator=generator, dtype=torch.float64) * 2 - 1).mul_(lim).float().to(device)


class FlatParams(object):
    """One flat fp32 buffer + named TF-layout views; companion grad / m / v buffers."""

    def __init__(self, table, device):
        self.table = OrderedDict(table)
        self.offsets = OrderedDict()
        off = 0
        for k, shp in self.table.items():
            self.offsets[k] = off
            off += (int(np.prod(shp)) + 63) // 64 * 64      # keep every view 256-byte aligned
        self.total = off
        self.data = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device)
        self.m = torch.zeros(off, dtype=torch.float32, device=device)
        self.v = torch.zeros(off, dtype=torch.float32, device=device)

    def _view(self, buf, k):
        n = int(np.prod(self.table[k]))
        return buf[self.offsets[k]:self.offsets[k] + n].view(*self.table[k])

    def p(self, k):
        return self._view(self.data, k)

    def g(self, k):
        return self._view(self.grad, k)

    def num_params(self):
        return int(sum(int(np.prod(s)) for s in self.table.values()))

    def state_dict(self):
        return OrderedDict((k, self.p(k).detach().cpu().clone()) for k in self.table)

    def load_state_dict(self, sd):
        for k in self.table:
            self.p(k).copy_(torch.as_tensor(sd[k]).to(self.data.device).view(*self.table[k]))


class GeneratorEngine(object):
    """GeneratorBE / GeneratorBE3 (skip_concat=False) forward + backward on hand-written sm_100a kernels."""

    def __init__(self, batch, output_shape, z_dim=3, filters=128, num_conv=4, repeat=0, name="G", device=None,
                 seed=123, init=None, inference=False, params=None):
        assert filters == 128, "the fused engine is specialised for filters=128 (config.py:21 default); other widths run ops_engine.OpsGeneratorEngine"
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.B, self.name, self.filters, self.num_conv = int(batch), name, filters, int(num_conv)
        self.spatial = [int(s) for s in output_shape[:-1]]
        self.cout = int(output_shape[-1])
        self.nd = len(self.spatial)
        self.z_dim = int(z_dim)
        self.rep = _repeat_num(self.spatial, repeat)
        self.level_shape = [[int(s // 2 ** (self.rep - 1 - i)) for s in self.spatial] for i in range(self.rep)]
        self.taps = 3 ** self.nd
        # ---- variables, TF names & layouts (model.py:19,26,42 / :61,68,84)
        tab = OrderedDict()
        n_fc = int(np.prod(self.level_shape[0])) * filters
        tab["%s/0_fc/weights" % name] = (self.z_dim, n_fc)
        tab["%s/0_fc/biases" % name] = (n_fc,)
        self.conv_names = []
        n = 1
        for i in range(self.rep):
            row = []
            for _ in range(self.num_conv):
                tab["%s/%d_conv/weights" % (name, n)] = (3,) * self.nd + (filters, filters)
                tab["%s/%d_conv/biases" % (name, n)] = (filters,)
                row.append("%s/%d_conv" % (name, n))
                n += 1
            self.conv_names.append(row)
        self.last_name = "%s/%d_conv" % (name, n)
        tab[self.last_name + "/weights"] = (3,) * self.nd + (filters, self.cout)
        tab[self.last_name + "/biases"] = (self.cout,)
        if params is not None:      # tf.variable_scope(reuse=True): the SAME variables (a FlatParams holding every name of `tab`)
            missing = [k for k in tab if k not in params.table or tuple(params.table[k]) != tuple(tab[k])]
            if missing:
                raise ValueError("reuse: variable %s does not exist with this shape" % missing[0])
            self.params = params
        else:
            self.params = FlatParams(tab, self.device)
        if params is not None:
            pass
        elif init is not None:
            self.params.load_state_dict(init)
        else:
            g = torch.Generator().manual_seed(seed)
            for k, shp in tab.items():
                if k.endswith("weights"):
                    self.params.p(k).copy_(xavier_uniform(tuple(shp), g, self.device))
        self.variables = list(tab.keys())
        self.inference = bool(inference)
        # Phase-decomposed upsample-conv (default; DFL_PHASE_UPCONV=0 selects the dense layer): the first conv of blocks 1..
        # reads upscale(s), i.e. each fine
        # voxel 2p + r only sees two coarse voxels per axis -> 2^nd convolutions with 2^nd taps and pre-summed weights on
        # the coarse tensor: 8/27 (3D) / 4/9 (2D) of the dense layer's forward and data-gradient FLOPs, and the data
        # gradient lands on the coarse grid (no full-resolution gradient of the up-sampled tensor is written or pooled).
        self.phase = (os.environ.get("DFL_PHASE_UPCONV", "1") == "1" and type(self).precision == "bf16" and self.rep > 1)
        self.phase_wgrad = self.phase and os.environ.get("DFL_PHASE_WGRAD", "1") == "1"
        self._alloc_operands()
        self.repack()
        self._alloc()
        # conv tiles per level (3D 2x16x8, 2D 16x16 voxels): on levels with fewer tiles than SMs the weight gradient runs
        # beside the data gradient on a second stream (backward()); DFL_FORK_BELOW_TILES overrides the threshold (0 = off)
        edges = (2, 16, 8) if self.nd == 3 else (16, 16)
        self._level_tiles = [self.B * int(np.prod([-(-int(ext) // e) for ext, e in zip(shp, edges)])) for shp in self.level_shape]
        on_gpu = torch.device(self.device).type == "cuda"
        n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count if on_gpu else 0
        self._fork_below = int(os.environ.get("DFL_FORK_BELOW_TILES", n_sm))
        # DFL_DETERMINISTIC=1: split-K gradient reductions in a fixed order (dfl_set_deterministic); their shared workspace
        # rules out the second stream
        if on_gpu and not inference and os.environ.get("DFL_DETERMINISTIC", "0") == "1" and not K.deterministic():
            K.set_deterministic(True, self.device)
        if on_gpu and K.deterministic():
            self._fork_below = 0
        self._side = torch.cuda.Stream(device=self.device) if (on_gpu and not inference and self._fork_below > 0) else None
        self.z = None
        self.adam_t = 0
        self.debug = None

    # ------------------------------------------------------------------ buffers (overridden by the fp32-grade engine)
    precision = "bf16"

    def _alloc_operands(self):
        """bf16 GEMM operands: forward [Cout, taps*Cin], dgrad [Cin, taps'*Cout] (the output conv reads the fp32 variable)"""
        filters = self.filters
        self.wf, self.wd = {}, {}
        for row in self.conv_names:
            for cn in row:
                self.wf[cn] = torch.empty(filters, self.taps * filters, dtype=torch.bfloat16, device=self.device)
                self.wd[cn] = torch.empty(filters, self.taps * filters, dtype=torch.bfloat16, device=self.device)
        self.wf_phase, self.wd_phase = {}, {}
        if getattr(self, "phase", False):
            P = 2 ** self.nd
            for row in self.conv_names[1:]:
                cn = row[0]
                self.wf_phase[cn] = torch.empty(P, filters, P * filters, dtype=torch.bfloat16, device=self.device)
                self.wd_phase[cn] = torch.empty(filters, P * P * filters, dtype=torch.bfloat16, device=self.device)
            self._t_phase = torch.empty(4 ** self.nd, filters, filters, dtype=torch.float32, device=self.device)

    def _alloc(self):
        """activations (bf16) and gradient scratch"""
        filters = self.filters
        bf = dict(dtype=torch.bfloat16, device=self.device)
        self.x0, self.y = [], []
        for i in range(self.rep):
            shp = [self.B] + self.level_shape[i] + [filters]
            self.x0.append(torch.empty(shp, **bf))
            if self.inference:     # forward only: two ping-pong buffers per level instead of one tensor per layer
                pp = [torch.empty(shp, **bf) for _ in range(min(2, self.num_conv))]
                self.y.append([pp[c % len(pp)] for c in range(self.num_conv)])
            else:
                self.y.append([torch.empty(shp, **bf) for _ in range(self.num_conv)])
        top = [self.B] + self.level_shape[-1] + [filters]
        self.s = torch.empty(top, **bf)
        self.pot = torch.empty([self.B] + self.spatial + [self.cout], dtype=torch.float32, device=self.device)
        # gradient scratch is sized for the finest level and re-viewed per level
        self._gbuf = [] if self.inference else [torch.empty(top, **bf) for _ in range(4)]

    # ------------------------------------------------------------------ helpers
    def _gview(self, k, level):
        shp = [self.B] + self.level_shape[level] + [self.filters]
        n = int(np.prod(shp))
        return self._gbuf[k].view(-1)[:n].view(shp)

    def repack(self):
        """fp32 master weights -> bf16 tensor-core operands (after every optimizer step): one launch for all layers."""
        # the table holds raw pointers into the flat parameter buffer: rebuild it whenever that buffer was replaced
        # (AEEngine swaps in its shared FlatParams after construction)
        if getattr(self, "_pack_table", None) is None or self._pack_src != self.params.data.data_ptr():
            self._pack_src = self.params.data.data_ptr()
            names = [cn for row in self.conv_names for cn in row]
            tab = [[self.params.p(cn + "/weights").data_ptr() for cn in names], [self.wf[cn].data_ptr() for cn in names],
                   [self.wd[cn].data_ptr() for cn in names]]
            self._pack_table = torch.tensor(tab, dtype=torch.int64, device=self.device)
            self._pack_n = len(names)
        K.pack_conv_weights_multi(self._pack_table, self._pack_n, self.taps, self.filters, self.filters)
        for cn in self.wf_phase:
            K.pack_phase_weights(self.params.p(cn + "/weights"), self.wf_phase[cn], self.wd_phase[cn])

    # ------------------------------------------------------------------ phase-decomposed upsample-conv
    def _phase_bits(self, r):
        return [(r >> (self.nd - 1 - a)) & 1 for a in range(self.nd)]

    def _phase_fwd(self, i, cn, bias):
        """first conv of block i >= 1 on x0[i] = upscale(s): one launch per output phase r; coarse voxel p reads
        x0[2 (p + off)] = s[p + off] (TMA element stride 2), off in {-1, 0} (r_a = 0) / {0, +1} (r_a = 1) per axis, and the
        result goes to the fine voxel 2p + r."""
        nd, F, P = self.nd, self.filters, 2 ** self.nd
        coarse, fine = self.level_shape[i - 1], self.level_shape[i]
        dense = 2.0 * self.B * float(np.prod(fine)) * F * F * self.taps
        for r in range(P):
            rr = self._phase_bits(r)
            taps = []
            for o in range(P):
                ob = self._phase_bits(o)
                off = [(ob[a] - 1) if rr[a] == 0 else ob[a] for a in range(nd)]
                taps.append([0] * (3 - nd) + [2 * v for v in off] + [o * F])
            K.conv_taps(self.x0[i], self.wf_phase[cn][r], bias, self.y[i][0], None, None, None, [self.B] + coarse, fine, F, 2,
                        taps, 2, rr, flags=K.CONV_LRELU, alg_flops=dense / P)

    def _phase_dgrad(self, i, cn, dpre, out):
        """data gradient of that layer w.r.t. the COARSE tensor s: dS[q] = sum_r sum_off dY[2 (q - off) + r] Wp[r][off]^T,
        ONE launch with all 2^nd x 2^nd (phase, tap) pairs accumulated in TMEM."""
        nd, F, P = self.nd, self.filters, 2 ** self.nd
        coarse, fine = self.level_shape[i - 1], self.level_shape[i]
        dense = 2.0 * self.B * float(np.prod(fine)) * F * F * self.taps
        taps = []
        for r in range(P):
            rr = self._phase_bits(r)
            for o in range(P):
                ob = self._phase_bits(o)
                off = [(ob[a] - 1) if rr[a] == 0 else ob[a] for a in range(nd)]
                taps.append([0] * (3 - nd) + [rr[a] - 2 * off[a] for a in range(nd)] + [(r * P + o) * F])
        K.conv_taps(dpre, self.wd_phase[cn], None, out, None, None, None, [self.B] + coarse, coarse, F, 2, taps, 1, [0] * nd,
                    alg_flops=dense)

    # ------------------------------------------------------------------ forward (model.py:5-46 / :48-87)
    def forward(self, z):
        P = self.params
        self.z = z.contiguous().float()
        assert self.z.shape == (self.B, self.z_dim)
        x0v = self.x0[0].view(self.B, -1)
        for b0 in range(0, self.B, 64):      # the FC kernel keeps <= 64 parameter rows in smem
            K.fc_fwd(self.z[b0:b0 + 64], P.p(self.name + "/0_fc/weights"), P.p(self.name + "/0_fc/biases"),
                     out=x0v[b0:b0 + 64])
        for i in range(self.rep):
            cur = self.x0[i]
            for c in range(self.num_conv):
                cn = self.conv_names[i][c]
                bias = P.p(cn + "/biases")
                if c == 0 and i > 0 and self.phase and self.num_conv > 1:
                    self._phase_fwd(i, cn, bias)
                elif c < self.num_conv - 1:
                    K.conv3x3(cur, self.wf[cn], bias, out=self.y[i][c], flags=K.CONV_LRELU)
                elif i < self.rep - 1:   # x += x0; x = upscale(x, 2); x0 = x   (model.py:34-37 / :76-79)
                    K.conv3x3(cur, self.wf[cn], bias, out=self.y[i][c], out2=self.x0[i + 1], residual=self.x0[i],
                              flags=K.CONV_LRELU | K.CONV_OUT2_UPSAMPLE)
                else:                    # x += x0                                 (model.py:39 / :82)
                    K.conv3x3(cur, self.wf[cn], bias, out=self.y[i][c], out2=self.s, residual=self.x0[i],
                              flags=K.CONV_LRELU)
                cur = self.y[i][c]
        K.lastconv_fwd(self.s, P.p(self.last_name + "/weights"), P.p(self.last_name + "/biases"), out=self.pot)
        return self.pot

    # ------------------------------------------------------------------ backward (TF autodiff of the above)
    def backward(self, dpot, dz=None, fused=None):
        """dpot: fp32 gradient w.r.t. the generator output.  Accumulates into params.grad (call zero_grad first).
        dz (fp32 [B, z_dim], optional): the gradient w.r.t. the generator input is ADDED to it (AE decoder).
        fused (3D, optional): dict(x=target velocity, w1=, w2=, loss3=, workspace=[, dpot=, vel=]) -- the curl / Jacobian-L1
        loss and its adjoint run in the PROLOGUE of the output conv's backward kernel (dfl_lastconv_curl_loss_bwd): `dpot` is
        then not read (pass None) and nothing is launched between the output conv's forward and backward kernels."""
        assert not self.inference, "inference engine has no backward pass"
        assert self.B <= 64, "fc_bwd keeps <= 64 parameter rows in smem"
        P = self.params
        nc = self.num_conv
        top = self.rep - 1
        ds = self._gview(0, top)
        dpre = self._gview(1, top)
        if fused is not None:
            K.lastconv_curl_loss_bwd(self.s, self.pot, fused["x"], P.p(self.last_name + "/weights"), self.y[top][nc - 1], ds,
                                     dpre, P.g(self.last_name + "/weights"), P.g(self.last_name + "/biases"), fused["loss3"],
                                     fused["workspace"], fused.get("w1", 1.0), fused.get("w2", 1.0), 1.0,
                                     dpot=fused.get("dpot"), vel=fused.get("vel"))
        else:
            K.lastconv_bwd(self.s, dpot, P.p(self.last_name + "/weights"), self.y[top][nc - 1], ds, dpre,
                           P.g(self.last_name + "/weights"), P.g(self.last_name + "/biases"))
        # gradient scratch: four full-resolution buffers re-viewed per level; roles rotate so that no kernel reads and
        # writes different voxels of the same buffer (b_ds holds ds, b_dp the running dL/d(pre-activation))
        b_ds, b_dp, b_a, b_b = 0, 1, 2, 3
        for i in range(top, -1, -1):
            ds = self._gview(b_ds, i)
            dpre = self._gview(b_dp, i)
            other = self._gview(b_a, i)
            cur_b, oth_b = b_dp, b_a
            phase0 = self.phase and i > 0 and nc > 1
            for c in range(nc - 1, -1, -1):
                cn = self.conv_names[i][c]
                xin = self.y[i][c - 1] if c > 0 else self.x0[i]
                if self.debug is not None:        # parity debugging: dL/d(pre-activation) of every layer
                    self.debug[cn] = dpre.clone()
                # weight and data gradient of a layer both only READ dpre: on the coarse levels, where the data-gradient
                # grid leaves SMs idle, the weight gradient runs beside it on a second stream (a fork / join per layer,
                # captured into the step's CUDA graph as parallel branches)
                fork = self._side is not None and self._level_tiles[i] < self._fork_below
                if c == 0 and phase0 and self.phase_wgrad:
                    # weight gradient on the phase weights (8/27 of the dense FLOPs in 3D): a 4^nd-tap stride-2 correlation of
                    # dpre (fine) with the layer's coarse input s = x0[i][::2, ::2(, ::2)] (exactly the values the forward
                    # read), folded back onto the 3^nd taps; the bias gradient is a column sum of dpre
                    s_c = self._gview(oth_b, i - 1)
                    K.gather_stride2(self.x0[i], s_c)
                    dense = 2.0 * self.B * float(np.prod(self.level_shape[i])) * self.filters * self.filters * self.taps
                    K.phase_wgrad(dpre, s_c, self._t_phase, P.g(cn + "/weights"), alg_flops=dense)
                    K.bias_grad(dpre, P.g(cn + "/biases"))
                    fork = False
                elif fork:
                    cur = torch.cuda.current_stream()
                    self._side.wait_stream(cur)
                    with torch.cuda.stream(self._side):
                        K.conv3x3_wgrad(xin, dpre, P.g(cn + "/weights"), P.g(cn + "/biases"))
                else:
                    K.conv3x3_wgrad(xin, dpre, P.g(cn + "/weights"), P.g(cn + "/biases"))
                if c > 0:     # dL/d(pre-activation of layer c-1) = dgrad * lrelu'(y[c-1])
                    K.conv3x3(dpre, self.wd[cn], None, out=other, mask_src=self.y[i][c - 1])
                    dpre, other = other, dpre
                    cur_b, oth_b = oth_b, cur_b
                elif phase0:  # data gradient straight onto the coarse grid (the residual branch's ds is pooled below)
                    self._phase_dgrad(i, cn, dpre, self._gview(b_b, i - 1))
                else:         # dL/dx0 = dgrad + residual-branch gradient ds
                    gx0 = self._gview(b_b, i)
                    K.conv3x3(dpre, self.wd[cn], None, out2=gx0, residual=ds)
                if fork:
                    cur.wait_stream(self._side)
            if i > 0:         # x0[i] = upscale(y4[i-1] + x0[i-1]): pool the children, then the lrelu derivative
                ysrc = self.y[i - 1][nc - 1]
                if phase0:
                    # ds_low = pool(ds) + (phase data gradient, already coarse, in b_b): written in place over the addend;
                    # dpre_low overwrites the fine dpre (its readers are queued ahead on the stream)
                    t = self._gview(b_b, i - 1)
                    K.pool_mask(ds, ysrc, t, self._gview(cur_b, i - 1), addend=t)
                    b_ds, b_b = b_b, b_ds
                    b_dp, b_a = cur_b, oth_b
                else:
                    K.pool_mask(gx0, ysrc, self._gview(b_ds, i - 1), self._gview(cur_b, i - 1))
                    b_dp, b_a = cur_b, oth_b
            else:
                gx0 = self._gview(b_b, 0)
                gw, gb = P.g(self.name + "/0_fc/weights"), P.g(self.name + "/0_fc/biases")
                if getattr(self, "accumulate_fc", False):
                    # dfl_fc_bwd OVERWRITES its outputs; under gradient accumulation (several backward passes per
                    # optimizer step) the FC gradient goes through a scratch pair and is added
                    if getattr(self, "_fc_scratch", None) is None:
                        self._fc_scratch = (torch.empty_like(gw), torch.empty_like(gb))
                    K.fc_bwd(self.z, gx0.view(self.B, -1), *self._fc_scratch)
                    gw.add_(self._fc_scratch[0])
                    gb.add_(self._fc_scratch[1])
                else:
                    K.fc_bwd(self.z, gx0.view(self.B, -1), gw, gb)
                if dz is not None:
                    K.fc_dz(gx0.view(self.B, -1), P.p(self.name + "/0_fc/weights"), dz, accumulate=True)

    def zero_grad(self):
        self.params.grad.zero_()

    # ------------------------------------------------------------------ optimizer (trainer.py:160-165,184)
    def adam_step(self, lr, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0):
        self.adam_t += 1
        t = self.adam_t
        lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
        P = self.params
        K.adam_step(P.data, P.grad, P.m, P.v, lr_t, beta1, beta2, eps, grad_scale)
        self.repack()

    def adam_lr_t(self, lr, beta1, beta2):
        """advance Adam's step counter and return lr_t = lr*sqrt(1-b2^t)/(1-b1^t) (host side of the graph path)"""
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

    def sgd_step(self, lr, grad_scale=1.0):
        P = self.params
        K.adam_step(P.data, P.grad, None, None, lr, 0.0, 0.0, 0.0, grad_scale)
        self.repack()

    # kernel launches of one forward+backward+update (for bench.py's gpu_launches claim)
    def launches_per_step(self):
        n_conv = self.rep * self.num_conv
        fwd = 1 + n_conv + 1
        bwd = 2 + n_conv * 3 + (self.rep - 1) + 1
        return fwd + 2 + bwd + 1 + 1 + n_conv   # + stencil(2) + memset + adam + repack
