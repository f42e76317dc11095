"""Train-step engines for the configurations the fused engines do not cover: `filters != 128` (run.bat:56,73 train the AE
with --filters=64) and `skip_concat=True` (model.py:29-33,71-75).

Same interface as engine.GeneratorEngine / encoder.AEEngine (forward / backward / zero_grad / optimizer steps / one flat
parameter buffer with TF-named views), so trainer.py, the checkpoint code and the data-parallel exchange do not care which
one they drive.  The network itself is model.generator_ops / encoder_ops: the reference's layer sequence on the
differentiable ops-level layers (ops.py, layers.py) -- the same tcgen05 conv kernels with channels zero-padded to 128-
blocks, torch autograd as the tape, un-fused residual add / upsample / concat.  Slower than the fused engines (no CUDA
graph, 4x padded FLOPs at filters=64) but every FLOP still runs in the sm_100a library: there is no CPU path.
"""
import math
from collections import OrderedDict

import torch

from . import kernels as K
from . import model as Mo
from . import ops
from .engine import FlatParams


class _Steps(object):
    """optimizer plumbing shared with the fused engines (TF-Adam on the flat buffers)"""

    def zero_grad(self):
        self.params.grad.zero_()

    def repack(self):
        pass            # operands are re-packed from the live variables inside every layer call

    def adam_lr_t(self, lr, beta1, beta2):
        self.adam_t += 1
        t = self.adam_t
        return lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)

    def adam_step(self, lr, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0):
        P = self.params
        K.adam_step(P.data, P.grad, P.m, P.v, self.adam_lr_t(lr, beta1, beta2), beta1, beta2, eps, grad_scale)

    def sgd_step(self, lr, grad_scale=1.0):
        P = self.params
        K.adam_step(P.data, P.grad, None, None, lr, 0.0, 0.0, 0.0, grad_scale)

    def optimizer_step(self, lr, adam=True, beta1=0.5, beta2=0.999, eps=1e-8, grad_scale=1.0):
        if adam:
            self.adam_step(lr, beta1, beta2, eps, grad_scale)
        else:
            self.sgd_step(lr, grad_scale)

    def _adopt(self, names):
        """move the variables the first forward created into ONE flat buffer (one Adam launch, one all-reduce, the
        checkpoint code's FlatParams interface); the ops-level Parameters become views of it"""
        tab = OrderedDict((n, tuple(ops.get_variable(n).shape)) for n in names)
        flat = FlatParams(tab, self.device)
        for n in names:
            v = ops.get_variable(n)
            flat.p(n).copy_(v.data)
            v.data = flat.p(n)
        self.params, self.variables = flat, list(names)
        self._plist = [ops.get_variable(n) for n in names]

    def _accumulate(self, names, grads):
        for n, g in zip(names, grads):
            if g is not None:
                self.params.g(n).add_(g.view_as(self.params.g(n)))


class OpsGeneratorEngine(_Steps):
    def __init__(self, batch, output_shape, z_dim=3, filters=128, num_conv=4, repeat=0, name="G", device=None, seed=123,
                 skip_concat=False, fresh_store=True):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.B, self.name, self.filters, self.num_conv, self.repeat = int(batch), name, int(filters), int(num_conv), int(repeat)
        self.output_shape, self.z_dim, self.skip_concat = list(output_shape), int(z_dim), bool(skip_concat)
        self.nd = len(output_shape) - 1
        self.adam_t = 0
        self.precision = "bf16"
        if fresh_store:
            ops.reset_variables(seed)
        with torch.no_grad():
            pot, names = self._net(torch.zeros(self.B, self.z_dim, device=self.device), reuse=False)
        self._adopt(names)
        self.pot = pot
        self._z_leaf = None

    def _net(self, z, reuse=True):
        return Mo.generator_ops(z, self.filters, self.output_shape, self.name, self.num_conv, 3, 3, self.repeat,
                                self.skip_concat, Mo.lrelu, reuse)

    def forward(self, z, need_dz=False):
        self.z = z.contiguous().float()
        self._z_leaf = self.z.detach().requires_grad_(True) if need_dz else self.z
        self._out, _ = self._net(self._z_leaf)
        self.pot = self._out.detach()
        return self.pot

    def backward(self, dpot, dz=None):
        """accumulates into params.grad; dz (optional, fp32 [B, z_dim]): the input gradient is ADDED to it"""
        ins = self._plist + ([self._z_leaf] if dz is not None else [])
        grads = torch.autograd.grad(self._out, ins, dpot.to(self._out.dtype))
        self._accumulate(self.variables, grads[:len(self._plist)])
        if dz is not None:
            dz.add_(grads[-1])
        self._out = None


class _Holder(object):
    pass


class OpsAEEngine(_Steps):
    """AE / AE3 (model.py:190-216) on the ops-level layers: encoder and decoder are two autograd graphs joined by hand at the
    latent code, exactly where encoder.AEEngine joins its two kernel sequences (loss_p gradient, sigmoid + KL for use_sparse)."""

    def __init__(self, batch, x_shape, filters=128, z_num=16, num_conv=4, repeat=0, name="AE", device=None, seed=123,
                 use_sparse=False, skip_concat=False):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.B, self.name, self.filters, self.z_num, self.num_conv, self.repeat = int(batch), name, int(filters), int(z_num), int(num_conv), int(repeat)
        self.x_shape, self.use_sparse = list(x_shape), bool(use_sparse)
        self.adam_t = 0
        ops.reset_variables(seed)
        self.dec = OpsGeneratorEngine(batch, x_shape, z_num, filters, num_conv, repeat, name + "/dec", self.device, seed,
                                      skip_concat, fresh_store=False)
        with torch.no_grad():
            _, enc_names = self._enc(torch.zeros([self.B] + self.x_shape, device=self.device), reuse=False)
        # reference variable order: encoder first, then decoder (model.py:192-199)
        self._enc_names, self._dec_names = list(enc_names), list(self.dec.variables)
        self._adopt(self._enc_names + self._dec_names)
        self.dec.params, self.dec._plist = self.params, [ops.get_variable(n) for n in self._dec_names]
        self._enc_plist = [ops.get_variable(n) for n in self._enc_names]
        self.enc = _Holder()
        self.enc.z = None
        self.dz = torch.zeros(self.B, z_num, dtype=torch.float32, device=self.device)
        self.z_sig = torch.zeros_like(self.dz)
        self.dz_lin = torch.zeros_like(self.dz)
        self.loss_kl = torch.zeros(1, dtype=torch.float32, device=self.device)

    def _enc(self, x, reuse=True):
        return Mo.encoder_ops(x, self.filters, self.z_num, self.name + "/enc", self.num_conv - 1, 3, self.repeat, Mo.lrelu, reuse)

    def forward(self, x):
        self._z_lin, _ = self._enc(x.contiguous().float())
        z = self._z_lin.detach().contiguous()
        if self.use_sparse:
            K.ae_sigmoid(z, self.z_sig)
            z = self.z_sig
        self.enc.z = z
        pot = self.dec.forward(z, need_dz=True)
        return pot, z

    def backward(self, dpot, p_num=0, sparsity=0.01, w5=1.0):
        """self.dz must already hold d(loss_p)/dz (dfl_ae_loss_p)"""
        self.dec.backward(dpot, dz=self.dz)
        dz = self.dz
        if self.use_sparse:
            K.ae_sparse_bwd(self.z_sig, self.dz, self.dz_lin, self.loss_kl, p_num, sparsity, w5)
            dz = self.dz_lin
        grads = torch.autograd.grad(self._z_lin, self._enc_plist, dz)
        self._accumulate(self._enc_names, grads)
        self._z_lin = None


class OpsNNEngine(_Steps):
    """arch=nn: the latent-space integrator model.NN (model.py:218-224) and the windowed roll-out of Trainer.build_model_nn
    (trainer.py:586-640) on the ops-level layers (dfl_gemm_f32, dfl_bn_act_*, dfl_dropout), torch autograd as the tape.
    All variables -- incl. the batch-norm moving statistics, which carry no gradient -- live in one flat buffer."""

    def __init__(self, batch, in_dim, filters, z_num, p_num, w_num, out_std, code_std, device=None, seed=123, dropout=0.1):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.B, self.in_dim, self.filters, self.z_num, self.p_num, self.w_num = int(batch), int(in_dim), int(filters), int(z_num), int(p_num), int(w_num)
        self.rescale = float(out_std) / float(code_std)
        self.dropout = dropout
        self.adam_t = 0
        ops.reset_variables(seed)
        ops.dropout_seed(seed)
        with torch.no_grad():      # declares the variables (inference mode: the moving statistics stay at their initial values)
            _, names = Mo.NN(torch.zeros(self.B, self.in_dim, device=self.device), self.filters, self.z_num, dropout=dropout,
                             train=False, reuse=False)
        self._adopt(names)
        self._trainable = [n for n in names if ops.get_variable(n).requires_grad]
        self._tparams = [ops.get_variable(n) for n in self._trainable]

    def net(self, x, train):
        return Mo.NN(x, self.filters, self.z_num, dropout=self.dropout, train=train, reuse=True)[0]

    def rollout(self, xw, train):
        """w_num chained predictions (trainer.py:590-612): the predicted code increment, re-normalised to the scale of x, is
        added to the code part of the input; the parameter part comes from the next frame of the window"""
        x_ = xw[:, 0, :]
        outs = []
        for i in range(self.w_num):
            y_ = self.net(x_, train)
            outs.append(y_.unsqueeze(1))
            if i < self.w_num - 1:
                x_ = torch.cat([x_[:, :-self.p_num] + y_ * self.rescale, xw[:, i + 1, -self.p_num:]], dim=-1)
        return torch.cat(outs, dim=1)

    def loss_and_grads(self, xw, yw):
        """loss = tf.losses.mean_squared_error(yw, yw_) (trainer.py:627,629); gradients accumulate into params.grad"""
        yw_ = self.rollout(xw, True)
        diff = (yw_ - yw).contiguous()
        loss, g = K.mse_loss(diff.detach(), 0.0)
        grads = torch.autograd.grad(yw_, self._tparams, g.view_as(yw_))
        self._accumulate(self._trainable, grads)
        return loss
