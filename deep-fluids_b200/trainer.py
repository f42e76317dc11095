"""Trainer with the reference's constructor / method / attribute names (trainer.py:12-293), eager and B200-native.

Where the reference builds a TF graph and runs `sess.run(self.g_optim)` (trainer.py:269), `train_step()` here issues
the same computation as a fixed sequence of sm_100a kernels through the C-ABI:
    x, y = batch()                                    (trainer.py:17; dequeue)
    G_s  = GeneratorBE(y)                             (trainer.py:138)   engine.forward
    G_   = curl(G_s); jacobians; g_loss               (trainer.py:140-172) one fused stencil kernel (loss + dL/dG_s)
    grads, Adam                                       (trainer.py:160-184) engine.backward + fused Adam
    g_lr update                                       (trainer.py:284-288)
Data parallelism (new; the reference is single-GPU): when torch.distributed is initialised the flat fp32 gradient
buffer is summed with ONE NCCL all-reduce per step and the 1/world factor is folded into the Adam kernel.
arch=dg (generator + patch discriminator, LSGAN terms) runs the same fused step plus an adversarial branch over the
differentiable layer ops (ops.py / layers.py).  Out of scope here: arch 'nn', summaries / PNG images.
"""
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import dp
from . import kernels as K
from .engine import GeneratorEngine
from .engine_fp32 import GeneratorEngineFP32


class Trainer(object):
    def __init__(self, config, batch_manager):
        self.config = config
        self.batch_manager = batch_manager
        self.x, self.y = batch_manager.batch()     # normalized input (trainer.py:17)
        self.device = self.x.device

        self.is_3d = config.is_3d
        self.dataset = config.dataset
        self.data_type = config.data_type
        self.arch = config.arch
        if 'nn' in self.arch:                       # trainer.py:24-27
            self.xt, self.yt = batch_manager.test_batch()
            self.xtw, self.ytw = batch_manager.test_batch(is_window=True)
            self.xw, self.yw = batch_manager.batch(is_window=True)

        self.res_x, self.res_y, self.res_z = config.res_x, config.res_y, config.res_z
        self.c_num = batch_manager.c_num
        self.b_num = config.batch_size
        self.test_b_num = config.test_batch_size
        self.repeat = config.repeat
        self.filters = config.filters
        self.num_conv = config.num_conv
        self.w1, self.w2 = config.w1, config.w2
        self.w3 = getattr(config, "w3", 0.005)      # weight of the adversarial term (arch=dg, trainer.py:178)

        self.use_c = config.use_curl
        if 'nn' in self.arch:
            self._init_nn(config, batch_manager)
            return
        spatial = list(self.x.shape[1:-1])
        if self.use_c:                              # trainer.py:48-55
            self.output_shape = spatial + [3 if self.is_3d else 1]
        else:
            self.output_shape = spatial + [self.x.shape[-1]]

        self.optimizer = config.optimizer
        self.beta1, self.beta2 = config.beta1, config.beta2
        self.model_dir = getattr(config, "model_dir", "")
        self.load_path = config.load_path

        self.start_step = config.start_step
        self.step = self.start_step                 # tf.Variable 'step' (trainer.py:65)
        self.max_step = int(config.max_epoch // batch_manager.epochs_per_step)   # trainer.py:67
        if getattr(config, "max_step", 0):
            self.max_step = int(config.max_step)

        self.lr_update = config.lr_update
        self.lr_min, self.lr_max = config.lr_min, config.lr_max
        if self.lr_update not in ('decay', 'step'):
            raise Exception("[!] Invalid lr update method")     # trainer.py:80
        self.g_lr = config.lr_max                   # trainer.py:72 / :77

        self.lr_update_step = config.lr_update_step
        self.log_step = config.log_step
        self.test_step = config.test_step
        self.save_sec = config.save_sec
        self.is_train = config.is_train

        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0

        if 'ae' in self.arch:                       # trainer.py:88-96
            self.z_num = config.z_num
            self.p_num = self.batch_manager.dof
            self.use_sparse = config.use_sparse
            self.sparsity = config.sparsity
            self.w4 = config.w4
            self.w5 = config.w5
            self.build_model_ae()
        else:
            self.build_model()

        # tf.train.Supervisor restores the latest checkpoint of the log dir (trainer.py:110-123): a TensorFlow bundle
        # (`checkpoint` + `model.ckpt-<step>.index/.data-*`, e.g. written by the reference itself) wins over `model.pt`
        self._restore()

    def _restore(self):
        if self.load_path:
            from . import tf_checkpoint as tfc
            prefix = tfc.latest_checkpoint(self.load_path)
            if prefix is not None:
                self.load_tf(prefix)
            elif os.path.exists(os.path.join(self.load_path, "model.pt")):
                self.load(os.path.join(self.load_path, "model.pt"))

    def _slot_variables(self):
        """variables that own Adam slots in a TensorFlow checkpoint: the trainable ones (arch=nn keeps the batch-norm moving
        statistics, which have none, in the same flat buffer)"""
        return list(getattr(self.engine, "_trainable", None) or self.engine.params.table)

    # ------------------------------------------------------------------ build (trainer.py:136-184)
    def build_model(self):
        if self.optimizer not in ('adam', 'gd'):
            raise Exception("[!] Invalid opimizer")              # sic, trainer.py:167
        self.precision = getattr(self.config, "precision", "bf16")
        self.accum = max(1, int(getattr(self.config, "grad_accum", 1)))     # micro-batches per optimizer step (strong scaling)
        if self.filters != 128:
            # run.bat:56,73-style widths: the general ops-level engine (same kernels, channels padded to 128-blocks)
            from .ops_engine import OpsGeneratorEngine
            if self.precision != "bf16":
                raise NotImplementedError("filters=%d runs the bf16 ops-level engine (precision=%s requested)" % (self.filters, self.precision))
            self._engine_cls = None
            self.engine = OpsGeneratorEngine(self.b_num, self.output_shape, z_dim=self.c_num, filters=self.filters,
                                             num_conv=self.num_conv, repeat=self.repeat, name="G", device=self.device,
                                             seed=self.config.random_seed)
        else:
            self._engine_cls = GeneratorEngineFP32 if self.precision == "fp32x3" else GeneratorEngine
            self.engine = self._engine_cls(self.b_num, self.output_shape, z_dim=self.c_num, filters=self.filters,
                                           num_conv=self.num_conv, repeat=self.repeat, name="G", device=self.device,
                                           seed=self.config.random_seed)
        self.G_var = self.engine.variables
        self.g_optim = self.train_step          # `sess.run(self.g_optim)` == `self.g_optim()` (trainer.py:184,269)
        self.G_s = self.engine.pot
        self._loss3 = torch.zeros(3, dtype=torch.float32, device=self.device)
        self._dpot = torch.empty_like(self.engine.pot)
        nb = K.cabi.lib().dfl_stencil_loss_workspace_bytes(K.dims_array(self.x.shape[:-1]), self.x.dim() - 2)
        self._ws = torch.empty(nb, dtype=torch.uint8, device=self.device)
        self.G_ = None
        self.g_loss = self.g_loss_l1 = self.g_loss_j_l1 = None
        self.use_graph = bool(int(os.environ.get("DFL_CUDA_GRAPH", "1"))) and self._engine_cls is not None
        self._captured = False
        self._xs = self._ys = None
        if self.accum > 1:
            if self._engine_cls is not GeneratorEngine or 'dg' in self.arch:
                raise NotImplementedError("gradient accumulation is built for the bf16 fused generator engine (arch=de)")
            self.engine.accumulate_fc = True
            self._loss3_acc = torch.zeros_like(self._loss3)
        if 'dg' in self.arch:
            self._build_discriminator()

    # ------------------------------------------------------------------ arch=dg (trainer.py:149-156,174-184; trainer3.py:27-34,53-63)
    def _build_discriminator(self):
        """D_x = DiscriminatorPatch(concat(x, x_vort)); D_G = DiscriminatorPatch(concat(G_, G_vort_), reuse=True): the
        variables D/Conv ... D/Conv_4 are created by the first call, like building the TF graph does."""
        from . import model as Mo, ops
        if self.precision != "bf16":
            raise NotImplementedError("arch=dg runs the bf16 tensor-core path (precision=%s requested)" % self.precision)
        self._ops = ops
        self._disc = Mo.DiscriminatorPatch3 if self.is_3d else Mo.DiscriminatorPatch
        ops.reset_variables(self.config.random_seed + 1)
        with torch.no_grad():
            d_in = torch.cat([self.x, K.jacobian_fwd(self.x.contiguous(), want_jac=False)[1]], dim=-1)
            _, self.D_var = self._disc(d_in, self.filters)
        self._d_params = [ops.get_variable(n) for n in self.D_var]
        self._d_m = [torch.zeros_like(p) for p in self._d_params]
        self._d_v = [torch.zeros_like(p) for p in self._d_params]
        self.use_graph = False                    # the adversarial branch is an autograd graph over the layer ops
        self.d_optim = self.train_step            # `sess.run([self.g_optim, self.d_optim])` is ONE train_step() here
        self.g_loss_real = self.d_loss_fake = self.d_loss_real = self.d_loss = None
        self._dg4 = None

    def _train_step_dg(self, x, y):
        """one `sess.run([self.g_optim, self.d_optim])` (trainer.py:266): both updates are computed from the same forward.
          g_loss = w1*L1 + w2*L1(J) + w3*mean((D_G - 1)^2)   -> generator variables
          d_loss = mean((D_x - 1)^2) + mean(D_G^2)           -> discriminator variables
        The supervised part and its gradient come from the fused stencil kernel as in arch=de; the adversarial part runs
        through the differentiable layer ops (curl / jacobian adjoints, tensor-core convs) and is added to dL/dpot.
        Adam: the reference minimises both losses with ONE AdamOptimizer object, whose beta-power accumulators therefore
        advance twice per step; within a step TF does not order the two updates, here both read the powers at the start of
        the step, i.e. bias correction with t = 2n - 1 at step n."""
        ops, eng = self._ops, self.engine
        eng.zero_grad()
        pot = eng.forward(y)
        K.stencil_loss_fwdbwd(pot, x, self.w1, self.w2, 1.0, dpot=self._dpot, loss3=self._loss3, workspace=self._ws)
        leaf = pot.detach().requires_grad_(True)
        jac = ops.jacobian3 if self.is_3d else ops.jacobian
        G_ = jac(leaf)[1] if self.is_3d else ops.curl(leaf)                       # trainer3.py:18 / trainer.py:140
        G_in = torch.cat([G_, jac(G_)[1]], dim=-1)                                # trainer.py:155
        with torch.no_grad():
            D_in = torch.cat([x, K.jacobian_fwd(x.contiguous(), want_jac=False)[1]], dim=-1)
        D_x, _ = self._disc(D_in, self.filters, reuse=True)
        D_G, _ = self._disc(G_in, self.filters, reuse=True)
        l_adv, seed_adv = K.mse_loss(D_G.detach(), 1.0, scale=self.w3)            # g_loss_real, d(w3 * it)/dD_G
        l_fake, seed_fake = K.mse_loss(D_G.detach(), 0.0)
        l_real, seed_real = K.mse_loss(D_x.detach(), 1.0)
        d_grads = torch.autograd.grad([D_x, D_G], self._d_params, [seed_real, seed_fake], retain_graph=True)
        (dpot_adv,) = torch.autograd.grad(D_G, leaf, seed_adv)
        self._dpot.add_(dpot_adv)
        eng.backward(self._dpot)
        self.G_, self.D_x, self.D_G = G_.detach(), D_x.detach(), D_G.detach()
        self._dg4 = torch.cat([l_adv, l_fake, l_real])
        # ---- gradient exchange + the two updates
        scale = dp.allreduce_grads_(eng.params.grad)
        if self.world > 1:
            flat = torch.cat([g.reshape(-1) for g in d_grads])
            dist.all_reduce(flat)
            d_grads = [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in d_grads]), d_grads)]
        if self.optimizer == 'adam':
            t = 2 * (eng.adam_t // 2) + 1
            eng.adam_t += 2
            lr_t = self.g_lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
            P = eng.params
            K.adam_step(P.data, P.grad, P.m, P.v, lr_t, self.beta1, self.beta2, 1e-8, scale)
            for p_, g_, m_, v_ in zip(self._d_params, d_grads, self._d_m, self._d_v):
                K.adam_step(p_.data, g_.contiguous(), m_, v_, lr_t, self.beta1, self.beta2, 1e-8, scale)
        else:
            P = eng.params
            K.adam_step(P.data, P.grad, None, None, self.g_lr, 0.0, 0.0, 0.0, scale)
            for p_, g_ in zip(self._d_params, d_grads):
                K.adam_step(p_.data, g_.contiguous(), None, None, self.g_lr, 0.0, 0.0, 0.0, scale)
        eng.repack()
        self._d_grads = d_grads
        self.step += 1
        return self._loss3

    def losses_dg(self):
        """(g_loss, g_loss_l1, g_loss_j_l1, g_loss_real, d_loss_fake, d_loss_real, d_loss) of the last step"""
        l = self._loss3.tolist()
        adv, fake, real = self._dg4.tolist()
        self.g_loss_l1, self.g_loss_j_l1 = l[1], l[2]
        self.g_loss_real, self.d_loss_fake, self.d_loss_real = adv, fake, real
        self.g_loss = l[0] + self.w3 * adv
        self.d_loss = real + fake
        return self.g_loss, l[1], l[2], adv, fake, real, self.d_loss

    # ------------------------------------------------------------------ one `sess.run(self.g_optim)`
    def _loss_and_grad(self, pot, x, want_vel=False):
        """loss terms + d loss / d (network output).  use_curl (trainer.py:138-140 / trainer3.py:16-18): ONE fused kernel
        (curl, both Jacobians, both L1 means and their adjoints).  use_curl=False (trainer.py:141-144): the output is the
        velocity itself -- nothing to fuse the curl with; un-fused standalone kernels."""
        if self.use_c:
            _, _, vel = K.stencil_loss_fwdbwd(pot, x, self.w1, self.w2, 1.0, want_vel=want_vel, dpot=self._dpot,
                                              loss3=self._loss3, workspace=self._ws)
            return vel
        K.velocity_loss_fwdbwd(pot, x, self.w1, self.w2, dvel=self._dpot, loss3=self._loss3)
        return pot

    def _fused_args(self, x, want_vel=False):
        """use_curl on the fused bf16 engines, 2D and 3D: the loss stencil runs in the prologue of the output conv's backward
        kernel (DFL_FUSED_LOSS=0 restores the separate stencil launch for A/B runs)."""
        from .encoder import AEEngine
        # 2D arch=ae: the decoder emits x's two channels and curl reads channel 0 only (ops.py:267-268) -- the fused kernel
        # takes a one-channel stream function, so that case keeps the standalone stencil (full-shape gradient, channel 1 = 0)
        pot_c = getattr(self.engine.dec if isinstance(self.engine, AEEngine) else self.engine, "cout", None)
        ok = (self.use_c and x.dtype == torch.float32 and (not self.is_3d or x.shape[-2] % 2 == 0)
              and (self.is_3d or pot_c == 1)
              and (getattr(self, "_engine_cls", None) is GeneratorEngine or isinstance(self.engine, AEEngine))
              and os.environ.get("DFL_FUSED_LOSS", "1") != "0")
        if not ok:
            return None
        if getattr(self, "_fws", None) is None:
            self._fws = K.lastconv_curl_loss_workspace(self.device)
        f = dict(x=x, w1=self.w1, w2=self.w2, loss3=self._loss3, workspace=self._fws)
        if want_vel:
            self._vel = torch.empty_like(x)
            f["vel"] = self._vel
        return f

    def _step_body_ae(self, x, y):
        """AE: s, z = AE(x); x_ = curl(s); loss = w1*L1 + w2*L1(J) + w4*mean((y[:,:,-1] - z[:,-p_num:])^2) (trainer.py:359-387)
        and its backward (everything before the gradient exchange)"""
        ae = self.ae
        y_last = y[:, :, -1].contiguous() if y.dim() == 3 else y[:, -self.p_num:].contiguous()
        ae.zero_grad()
        pot, z = ae.forward(x)
        fused = self._fused_args(x)
        if fused is None:
            self._loss_and_grad(pot, x)   # use_curl: x_ = curl(s) (trainer.py:359-361); else x_ = the decoder output (:363)
        K.ae_loss_p(z, y_last, ae.dz, self._loss_p, self.w4)
        if fused is None:
            ae.backward(self._dpot, self.p_num, self.sparsity, self.w5)
        else:
            ae.backward(None, self.p_num, self.sparsity, self.w5, fused=fused)

    def _step_body_a(self, x, y, want_vel=False, zero=True):
        """zero grads, forward, fused loss + dL/dpot, backward (everything before the gradient exchange)"""
        if 'ae' in self.arch:
            return self._step_body_ae(x, y)
        eng = self.engine
        if zero:
            eng.zero_grad()
        with K.nvtx_range("forward"):
            pot = eng.forward(y)
        fused = self._fused_args(x, want_vel)
        if fused is not None:
            with K.nvtx_range("loss+backward"):
                eng.backward(None, fused=fused)
            vel = fused.get("vel")
        else:
            with K.nvtx_range("loss"):
                vel = self._loss_and_grad(pot, x, want_vel)
            with K.nvtx_range("backward"):
                eng.backward(self._dpot)
        if self.accum > 1:
            self._loss3_acc.add_(self._loss3)
        return vel

    def _step_body_b(self, scale):
        with K.nvtx_range("optimizer"):
            self._step_body_b_(scale)

    def _step_body_b_(self, scale):
        self.engine.optimizer_step_dev(self._lr_dev, self.optimizer == 'adam', self.beta1, self.beta2, 1e-8, scale)

    def _capture(self):
        """Capture the step as CUDA graph(s): the reference issues ONE sess.run per step; here ~115 kernel launches
        are replayed with one cudaGraphLaunch (two when a gradient all-reduce sits in between)."""
        self._xs, self._ys = torch.empty_like(self.x), torch.empty_like(self.y)
        self._xs.copy_(self.x)
        self._ys.copy_(self.y)
        self._lr_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        scale = 1.0 / (self.world * self.accum)
        # one eager pass on a side stream: sets every kernel's attributes, warms allocator (grad buffers stay zeroed
        # at the end because lr = 0 leaves the weights untouched)
        m0, v0 = self.engine.params.m.clone(), self.engine.params.v.clone()   # (a resumed run has non-zero moments)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step_body_a(self._xs, self._ys)
            self._step_body_b(scale)          # lr_dev == 0 -> parameters unchanged; Adam moments are restored below
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.engine.params.m.copy_(m0)
        self.engine.params.v.copy_(v0)
        del m0, v0
        if self.accum > 1:
            self.engine.zero_grad()
            self._loss3_acc.zero_()
        # DFL_NATIVE_NCCL=1: the gradient all-reduce goes through the library's own communicator (dfl_allreduce) on the
        # capture stream, so forward, backward, exchange and update are ONE graph launch per step at any world size
        native = dp.use_native()
        if native:
            dp.native_comm()
        l0 = K.PROF.launches
        self._graph_a = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_a):
            # with gradient accumulation graph_a is ONE micro-batch (no zeroing, no update): it is replayed accum times
            self._step_body_a(self._xs, self._ys, zero=(self.accum == 1))
            if self.accum == 1 and (self.world == 1 or native):
                if native:
                    dp.allreduce_grads_(self.engine.params.grad)
                self._step_body_b(scale)
        self._graph_b = None
        if (self.world > 1 and not native) or self.accum > 1:
            self._graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_b):
                self._step_body_b(scale)
        self.launches_per_step = K.PROF.launches - l0
        self._captured = True

    def train_step(self, x=None, y=None, want_vel=False):
        if 'ae' in self.arch:
            return self.train_step_ae(x, y)[0]
        if x is None:
            x, y = self.batch_manager.batch()
        self.x, self.y = x, y
        eng = self.engine
        if 'dg' in self.arch:
            return self._train_step_dg(x, y)
        if self.use_graph and not want_vel:
            return self._replay_step(x, y)
        # ---- eager path (debug / per-kernel timing / want_vel / ops-level engine) ----
        if self.accum > 1:
            raise NotImplementedError("gradient accumulation runs through the captured step (DFL_CUDA_GRAPH=1)")
        vel = self._step_body_a(x, y, want_vel)
        if want_vel:
            self.G_ = vel
        scale = dp.allreduce_grads_(eng.params.grad)
        if self.optimizer == 'adam':
            eng.adam_step(self.g_lr, self.beta1, self.beta2, 1e-8, scale)
        else:
            eng.sgd_step(self.g_lr, scale)
        self.step += 1
        return self._loss3

    def _replay_step(self, x, y):
        """one optimizer step through the captured CUDA graph(s) (generator and AE)"""
        eng = self.engine
        if True:
            if not self._captured:
                self._capture()
            if x.data_ptr() != self._xs.data_ptr():
                self._xs.copy_(x, non_blocking=True)
            if y.data_ptr() != self._ys.data_ptr():
                self._ys.copy_(y, non_blocking=True)
            lr_t = eng.adam_lr_t(self.g_lr, self.beta1, self.beta2) if self.optimizer == 'adam' else self.g_lr
            # by value: the scalar travels as a launch argument of a fill kernel.  (A pinned host slot + async H2D copy is
            # read when the STREAM reaches the copy, so a host running ahead would overwrite it: step k would see the
            # lr_t of step k+n.)
            self._lr_dev.fill_(float(lr_t))
            if self.accum > 1:
                # strong scaling: one optimizer step over `accum` micro-batches of b_num samples each (gradients summed in
                # params.grad, the mean's 1/accum folded into the Adam kernel's grad_scale), ONE all-reduce per optimizer step
                eng.params.grad.zero_()
                self._loss3_acc.zero_()
                self._graph_a.replay()
                for _ in range(self.accum - 1):
                    xm, ym = self.batch_manager.batch()
                    self._xs.copy_(xm, non_blocking=True)
                    self._ys.copy_(ym, non_blocking=True)
                    self._graph_a.replay()
                self._loss3.copy_(self._loss3_acc / self.accum)
                K.PROF.launches += self.launches_per_step * (self.accum - 1)
            else:
                self._graph_a.replay()
            if self._graph_b is not None:
                if self.world > 1:
                    dp.allreduce_grads_(eng.params.grad)     # ONE NCCL all-reduce over the flat gradient buffer
                self._graph_b.replay()
            K.PROF.launches += self.launches_per_step
            self.step += 1
            return self._loss3

    def update_lr(self, step_in_loop):
        """`sess.run(self.g_lr_update)` (trainer.py:284-288): evaluated with the already-incremented global step."""
        if self.lr_update == 'step':
            if step_in_loop % self.lr_update_step == self.lr_update_step - 1:
                self.g_lr = max(self.g_lr * 0.5, self.lr_min)
        else:
            self.g_lr = self.lr_min + 0.5 * (self.lr_max - self.lr_min) * (math.cos(self.step * math.pi / self.max_step) + 1)

    def losses(self):
        """(g_loss, g_loss_l1, g_loss_j_l1) of the last step as Python floats (one D2H read)."""
        if 'dg' in self.arch and self._dg4 is not None:
            return list(self.losses_dg()[:3])
        l = self._loss3.tolist()
        self.g_loss, self.g_loss_l1, self.g_loss_j_l1 = l
        return l

    # ------------------------------------------------------------------ loops (trainer.py:224-293)
    def train(self):
        if 'ae' in self.arch:
            self.train_ae()
        elif 'nn' in self.arch:
            self.train_nn()
        else:
            self.train_()

    def _checkpoint(self):
        if self.model_dir and self.rank == 0:
            self.save(os.path.join(self.model_dir, 'model.pt'))
            self.save_tf(self.model_dir)          # saver.save(sess, model_dir/model.ckpt, global_step=step), trainer.py:291-292

    def _periodic_checkpoint(self):
        """tf.train.Supervisor(save_model_secs=self.save_sec) (trainer.py:110-118): a checkpoint every save_sec seconds"""
        now = time.time()
        if not hasattr(self, "_last_save"):
            self._last_save = now
        elif self.save_sec and now - self._last_save >= self.save_sec:
            self._checkpoint()
            self._last_save = now

    def train_(self):
        for step in range(self.start_step, self.max_step):
            self.train_step()
            if step % self.log_step == 0 or step == self.max_step - 1:
                ep = step * self.batch_manager.epochs_per_step
                loss = self.losses()[0]
                assert not np.isnan(loss), 'Model diverged with loss = NaN'      # trainer.py:275
                if self.rank == 0:
                    print("\n[{}/{}/ep{:.2f}] Loss: {:.6f}".format(step, self.max_step, ep, loss))
            self.update_lr(step)
            self._periodic_checkpoint()
        self._checkpoint()
        self.batch_manager.stop_thread()

    # ------------------------------------------------------------------ auto-encoder (trainer.py:357-462, trainer3.py:240-345)
    def build_model_ae(self):
        from .encoder import AEEngine
        if self.optimizer not in ('adam', 'gd'):
            raise Exception("[!] Invalid opimizer")
        if getattr(self.config, "grad_accum", 1) > 1:
            raise NotImplementedError("gradient accumulation is built for arch=de")
        self.accum = 1
        x_shape = list(self.x.shape[1:])
        if self.filters != 128:      # run.bat:56,73 (--filters=64): the general ops-level engine
            from .ops_engine import OpsAEEngine
            self.ae = OpsAEEngine(self.b_num, x_shape, self.filters, self.z_num, self.num_conv, self.repeat, "AE", self.device,
                                  self.config.random_seed, use_sparse=self.use_sparse)
        else:
            self.ae = AEEngine(self.b_num, x_shape, self.filters, self.z_num, self.num_conv, self.repeat, "AE", self.device,
                               self.config.random_seed, use_sparse=self.use_sparse)
        self.engine = self.ae                        # checkpoint / DP code paths use `.engine.params`
        self.var = self.ae.variables
        self.optim = self.train_step_ae         # `sess.run(self.optim)` == `self.optim()` (trainer.py:396,437)
        self._loss3 = torch.zeros(3, dtype=torch.float32, device=self.device)
        self._loss_p = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._dpot = torch.empty_like(self.ae.dec.pot)
        nb = K.cabi.lib().dfl_stencil_loss_workspace_bytes(K.dims_array(self.x.shape[:-1]), self.x.dim() - 2)
        self._ws = torch.empty(nb, dtype=torch.uint8, device=self.device)
        # the whole AE step (~450 launches at 128^3) replays as one CUDA graph, like the generator's
        self.use_graph = bool(int(os.environ.get("DFL_CUDA_GRAPH", "1"))) and isinstance(self.ae, AEEngine)
        self._engine_cls = None
        self._captured = False
        self._xs = self._ys = None
        self.loss = self.loss_l1 = self.loss_j_l1 = self.loss_p = None

    def train_step_ae(self, x=None, y=None):
        """one `sess.run(self.optim)` of the AE graph (trainer.py:437): s,z = AE(x); x_ = curl(s);
        loss = w1*L1 + w2*L1(J) + w4*mean((y[:,:,-1] - z[:,-p_num:])^2)   (trainer.py:382-387)"""
        if x is None:
            x, y = self.batch_manager.batch()
        self.x, self.y = x, y
        ae = self.ae
        if self.use_graph:
            self._replay_step(x, y)
            return self._loss3, self._loss_p
        self._step_body_ae(x, y)
        scale = dp.allreduce_grads_(ae.params.grad)
        ae.optimizer_step(self.g_lr, self.optimizer == 'adam', self.beta1, self.beta2, 1e-8, scale)
        self.step += 1
        return self._loss3, self._loss_p

    # ---- AE inference pieces of test_ae / autoencode (trainer.py:464-583): encode x -> z, decode z -> velocity
    def _chunks(self, t):
        n = t.shape[0]
        for b0 in range(0, n, self.b_num):
            c = t[b0:b0 + self.b_num]
            k = c.shape[0]
            if k < self.b_num:
                c = torch.cat([c, c.new_zeros((self.b_num - k,) + tuple(c.shape[1:]))])
            yield c.contiguous(), k

    def encode(self, x):
        """latent codes z [n, z_num] of velocity fields x [n,(D,)H,W,C]  (`sess.run(self.z, {self.x: x})`, trainer.py:504)"""
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            return torch.cat([self._encode_chunk(c)[:k].clone() for c, k in self._chunks(x)])

    def _encode_chunk(self, c):
        if hasattr(self.ae.enc, "forward"):
            z = self.ae.enc.forward(c)
            if self.use_sparse:                       # model.py:196 / :210: the code IS the sigmoid output
                zs = torch.empty_like(z)
                K.ae_sigmoid(z, zs)
                z = zs
            return z
        return self.ae.forward(c)[1]

    def decode(self, z):
        """velocity fields from latent codes (`sess.run(self.x_, {self.z: z})`, trainer.py:551-552), normalised units"""
        z = torch.as_tensor(z, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            return torch.cat([self._velocity(self.ae.dec.forward(c))[:k].clone() for c, k in self._chunks(z)])

    def autoencode(self, x):
        return self.decode(self.encode(x))

    def losses_ae(self):
        l = self._loss3.tolist()
        lp = float(self._loss_p.item())
        self.loss_l1, self.loss_j_l1, self.loss_p = l[1], l[2], lp
        self.loss = l[0] + self.w4 * lp
        if self.use_sparse:
            self.loss_kl = float(self.ae.loss_kl.item())
            self.loss += self.w5 * self.loss_kl
        return self.loss, l[1], l[2], lp

    def train_ae(self):
        for step in range(self.start_step, self.max_step):
            self.train_step_ae()
            if step % self.log_step == 0 or step == self.max_step - 1:
                ep = step * self.batch_manager.epochs_per_step
                loss = self.losses_ae()[0]
                assert not np.isnan(loss), 'Model diverged with loss = NaN'      # trainer.py:444
                if self.rank == 0:
                    print("\n[{}/{}/ep{:.2f}] Loss: {:.6f}".format(step, self.max_step, ep, loss))
            self.update_lr(step)
            self._periodic_checkpoint()
        self._checkpoint()
        self.batch_manager.stop_thread()

    # ------------------------------------------------------------------ inference (trainer.py:295-354, 750-771)
    def build_test_model(self):
        """`reuse=True` generator on z:[test_b_num, c_num] (trainer.py:295-304): a forward-only engine sharing the
        trained variables.  The 3D trainer uses the 3D curl here (the reference's trainer3.py:188 applies the 2D curl
        to the 3-channel potential -- a bug that is not reproduced)."""
        cls = getattr(self, "_engine_cls", GeneratorEngine)
        if cls is None:                   # ops-level engine: the layer functions take any batch size
            self.test_engine = self.engine
            return
        # the SAME variables (the engine's flat parameter buffer, not a snapshot): a sweep run after further training steps or
        # after load() / load_tf() sees the current weights; its bf16 operands are re-packed before every sweep
        self.test_engine = cls(self.test_b_num, self.output_shape, z_dim=self.c_num, filters=self.filters,
                               num_conv=self.num_conv, repeat=self.repeat, name="G", device=self.device,
                               params=self.engine.params, inference=True)

    def generate_velocity(self, z):
        """G_ = curl(G_s(z)) for parameters z [n, c_num], n a multiple of test_b_num or <= it (sess.run(self.G_, {z}))."""
        if not hasattr(self, "test_engine"):
            self.build_test_model()
        z = torch.as_tensor(z, dtype=torch.float32, device=self.device)
        if self.test_engine is not self.engine:
            self.test_engine.repack()
        outs = []
        for b0 in range(0, z.shape[0], self.test_b_num):
            zb = z[b0:b0 + self.test_b_num]
            n = zb.shape[0]
            if n < self.test_b_num:
                zb = torch.cat([zb, zb.new_zeros(self.test_b_num - n, zb.shape[1])])
            with torch.no_grad():
                pot = self.test_engine.forward(zb)
            outs.append(self._velocity(pot)[:n].clone())
        return torch.cat(outs)

    def _velocity(self, out):
        """G_ of the network output: curl(G_s) with use_curl (trainer.py:140 / trainer3.py:18), else the output itself"""
        return K.curl_fwd(out) if self.use_c else out

    def test(self):
        if 'ae' in self.arch:                       # trainer.py:306-312
            self.test_ae()
        elif 'nn' in self.arch:
            self.test_nn()
        else:
            self.test_()

    def test_(self):
        """Sweep the last parameter at fixed p1,p2 = 10,2, de-normalise and dump `<model_dir>/10_2/%d.npz`
        (trainer.py:314-354)."""
        self.build_test_model()
        p1, p2 = 10, 2
        y1, y2, y3 = (int(v) for v in self.batch_manager.y_num[:3])
        assert y3 % self.test_b_num == 0
        c1 = p1 / float(y1 - 1) * 2 - 1
        c2 = p2 / float(y2 - 1) * 2 - 1
        z_c = np.zeros((y3, self.c_num), dtype=np.float32)
        z_c[:, 0], z_c[:, 1], z_c[:, -1] = c1, c2, np.linspace(-1, 1, num=y3)
        G = self.generate_velocity(z_c)
        G, _ = self.batch_manager.denorm(x=G)
        out_dir = os.path.join(self.model_dir, '%d_%d' % (p1, p2))
        os.makedirs(out_dir, exist_ok=True)
        G = G.cpu().numpy()
        for i, G_ in enumerate(G):
            np.savez_compressed(os.path.join(out_dir, '%d.npz' % i), x=G_)
        return out_dir

    # ------------------------------------------------------------------ graph tensors other code pokes (trainer.py:29-32,146,361)
    # The train step never materialises them (the fused stencil kernel works on residuals); they are evaluated on demand
    # with the stand-alone kernels from the tensors of the last step.
    @property
    def x_jaco(self):
        """jacobian(self.x)[0] (trainer.py:29-32 / trainer3 via jacobian3)"""
        return K.jacobian_fwd(self.x.contiguous())[0]

    @property
    def x_vort(self):
        return K.jacobian_fwd(self.x.contiguous())[1]

    @property
    def G_jaco_(self):
        """jacobian(self.G_)[0] of the last `train_step(want_vel=True)` (trainer.py:146)"""
        return None if self.G_ is None else K.jacobian_fwd(self.G_.contiguous())[0]

    @property
    def G_vort_(self):
        return None if self.G_ is None else K.jacobian_fwd(self.G_.contiguous())[1]

    @property
    def z(self):
        """AE: latent code of the last step (`self.s, self.z, self.var = AE(...)`, trainer.py:359)"""
        return self.ae.enc.z if hasattr(self, "ae") else None

    @property
    def x_(self):
        """AE: reconstructed velocity curl(s) of the last step (trainer.py:361)"""
        return self._velocity(self.ae.dec.pot) if hasattr(self, "ae") else None

    @property
    def s(self):
        """AE: the decoder's stream function / vector potential of the last step (trainer.py:359-361)"""
        return self.ae.dec.pot if hasattr(self, "ae") else None

    # ------------------------------------------------------------------ reference methods outside the hot path
    # (kept on the class so a call fails with the reason instead of an AttributeError)
    def _out_of_scope(self, what, where):
        raise NotImplementedError("%s (%s) is outside the B200 hot path this package replaces" % (what, where))

    def generate(self, inputs, root_path=None, idx=None):
        self._out_of_scope("generate(): PNG sample grids", "trainer.py:750-771; use generate_velocity(z) for the fields")

    def get_vort_image(self, x):
        self._out_of_scope("get_vort_image(): vorticity PNG rendering", "trainer.py:773-790")

    def build_test_model_ae(self):
        """trainer.py:464-473 re-declares the AE on a [test_b_num, ...] placeholder with reuse=True.  The engines here take
        the trained variables as they are: encode() / decode() chunk any number of fields through the training engines."""
        self.code_path = getattr(self.config, "code_path", "")

    def test_ae(self):
        """trainer.py:475-583.  Without --code_path: encode the WHOLE dataset in file order (`batch_manager.batch_`) and dump
        `<load_path>/code<z_num>.npz` = {x: codes of frames 0..f-2 of every simulation, y: codes of frames 1..f-1, p: the
        per-frame parameter increments from <dataset>/n.npz, s: #simulations, f: #frames} -- the training set of arch=nn.
        With --code_path: decode `code_out.npz` ({z_out, z_gt}: [sims][frames, z_num]) back to velocity fields, de-normalise
        and dump `<load_path>/v<s>.npz` = {v, v_gt} (the reference renders them to PNG frame pairs instead -- rendering is
        out of scope; the arrays are what its own commented-out `np.savez_compressed(v_path, v=v, v_gt=v_gt)` would write)."""
        self.build_test_model_ae()
        bm = self.batch_manager
        out_root = self.load_path or self.model_dir
        if not self.code_path:
            with np.load(os.path.join(bm.root, 'n.npz')) as data:
                nx = data['nx']
                nz = data['nz'] if self.is_3d else None
            num_sims, num_frames = int(nx.shape[0]), int(nx.shape[1])
            p_list = (nx[:, 1:] - nx[:, :-1]).reshape([-1, 1])
            if self.is_3d:
                p_list = np.concatenate((p_list, (nz[:, 1:] - nz[:, :-1]).reshape([-1, 1])), axis=-1)
            c_list = np.concatenate([self.encode(xb).cpu().numpy() for xb, _ in bm.batch_(self.test_b_num)])
            assert c_list.shape[0] == num_sims * num_frames, (c_list.shape, num_sims, num_frames)
            c = c_list.reshape(num_sims, num_frames, -1)
            x_list = c[:, :-1].reshape(-1, c.shape[-1])
            y_list = c[:, 1:].reshape(-1, c.shape[-1])
            code_path = os.path.join(out_root, 'code%d.npz' % self.z_num)
            np.savez_compressed(code_path, x=x_list, y=y_list, p=p_list, s=num_sims, f=num_frames)
            return code_path
        with np.load(os.path.join(self.code_path, 'code_out.npz'), allow_pickle=True) as data:
            z_, z_gt_ = data['z_out'], data['z_gt']
        paths = []
        for s_ in range(len(z_)):
            v, _ = bm.denorm(x=self.decode(np.asarray(z_[s_], dtype=np.float32)))
            v_gt, _ = bm.denorm(x=self.decode(np.asarray(z_gt_[s_], dtype=np.float32)))
            paths.append(os.path.join(out_root, 'v%d.npz' % s_))
            np.savez_compressed(paths[-1], v=v.cpu().numpy(), v_gt=v_gt.cpu().numpy())
        return paths

    # ------------------------------------------------------------------ arch=nn (trainer.py:586-747)
    def _init_nn(self, config, batch_manager):
        self.optimizer = config.optimizer
        self.beta1, self.beta2 = config.beta1, config.beta2
        self.model_dir = getattr(config, "model_dir", "")
        self.load_path = config.load_path
        self.start_step = config.start_step
        self.step = self.start_step
        self.max_step = int(config.max_epoch // batch_manager.epochs_per_step)
        if getattr(config, "max_step", 0):
            self.max_step = int(config.max_step)
        self.lr_update, self.lr_min, self.lr_max = config.lr_update, config.lr_min, config.lr_max
        if self.lr_update not in ('decay', 'step'):
            raise Exception("[!] Invalid lr update method")
        self.g_lr = config.lr_max
        self.lr_update_step, self.test_step, self.save_sec = config.lr_update_step, config.test_step, config.save_sec
        self.is_train = config.is_train
        self.log_step = batch_manager.train_steps                    # trainer.py:126-128
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.z_num, self.p_num, self.w_num = config.z_num, batch_manager.dof, config.w_size
        self.accum = 1
        self.build_model_nn()
        self._restore()

    def build_model_nn(self):
        """y_ = NN(x); roll-out over the window; loss = mean squared error of the w_num chained predictions
        (trainer.py:586-640).  Only Adam (the reference raises otherwise, trainer.py:622)."""
        from .ops_engine import OpsNNEngine
        if self.optimizer != 'adam':
            raise Exception("[!] Caution! Paper didn't use {} opimizer other than Adam".format(self.optimizer))
        bm = self.batch_manager
        self.engine = OpsNNEngine(self.b_num, self.z_num + self.p_num, self.filters, self.z_num, self.p_num, self.w_num,
                                  bm.out_std, bm.code_std, self.device, self.config.random_seed)
        self.var = self.engine.variables
        self.optim = self.train_step_nn
        self.use_graph = False
        self.loss = self.loss_train_w = self.l_test = self.l_test_w = None

    def train_step_nn(self, xw=None, yw=None):
        """one `sess.run(self.optim)` (trainer.py:645)"""
        if xw is None:
            xw, yw = self.batch_manager.batch(is_window=True)
        eng = self.engine
        eng.zero_grad()
        self.loss_train_w = eng.loss_and_grads(xw, yw)
        scale = dp.allreduce_grads_(eng.params.grad)
        eng.adam_step(self.g_lr, self.beta1, self.beta2, 1e-8, scale)
        self.step += 1
        return self.loss_train_w

    def _test_losses_nn(self):
        """mean test losses over one pass of the test iterators (trainer.py:650-662), inference mode"""
        bm, eng = self.batch_manager, self.engine
        bm.init_test_it()
        with torch.no_grad():
            tl = []
            for _ in range(bm.test_steps):
                xt, yt = bm.test_batch()
                tl.append(float(K.mse_loss((eng.net(xt, False) - yt).contiguous(), 0.0, want_grad=False)[0]))
            tw = []
            for _ in range(bm.test_w_steps):
                xtw, ytw = bm.test_batch(is_window=True)
                tw.append(float(K.mse_loss((eng.rollout(xtw, False) - ytw).contiguous(), 0.0, want_grad=False)[0]))
        self.l_test, self.l_test_w = sum(tl) / len(tl), sum(tw) / len(tw)
        return self.l_test, self.l_test_w

    def train_nn(self):
        for step in range(self.start_step, self.max_step):
            self.train_step_nn()
            if step % self.log_step == 0 or step == self.max_step - 1:
                test_loss, test_loss_w = self._test_losses_nn()
                self.loss = float(self.loss_train_w)
                assert not np.isnan(self.loss), 'Model diverged with loss = NaN'      # trainer.py:668
                if self.rank == 0:
                    print("\n[{}/{}] Loss: {:.6f}/{:.6f}/{:.6f}".format(step, self.max_step, self.loss, test_loss, test_loss_w))
            self.update_lr(step)
            self._periodic_checkpoint()
        self._checkpoint()

    def test_nn(self):
        """integrate every test simulation in latent space from its first frame and dump `<load_path>/code_out.npz` =
        {z_out, z_gt} ([simulations][frames, z_num], de-normalised codes) -- the input of test_ae's decode branch
        (trainer.py:698-747)"""
        bm, eng = self.batch_manager, self.engine
        z_out_list, z_gt_list = [], []
        nf, p = bm.num_frames, self.p_num
        with torch.no_grad():
            for i in range(bm.num_test_scenes):
                z0 = bm.x_test[i * (nf - 1)]
                z_in = z0.reshape(1, -1)
                z_out = [z0[:-p].reshape(1, -1) * bm.code_std]
                z_gt = [z0[:-p].reshape(1, -1) * bm.code_std]
                for t in range(nf - 1):
                    y_gt = bm.y_test[i * (nf - 1) + t] * bm.out_std + bm.x_test[i * (nf - 1) + t, :-p] * bm.code_std
                    z_gt.append(y_gt.reshape(1, -1))
                    pred = eng.net(torch.as_tensor(z_in, dtype=torch.float32, device=self.device), False).cpu().numpy()
                    y_ = pred * bm.out_std + z_in[:, :-p] * bm.code_std
                    y_[0, -p:] = y_gt[-p:]
                    z_out.append(y_)
                    if t < nf - 2:
                        zt = bm.x_test[i * (nf - 1) + t + 1]
                        z_in = np.append(y_.flatten() / bm.code_std, zt[-p:]).reshape(1, -1)
                z_out_list.append(np.concatenate(z_out))
                z_gt_list.append(np.concatenate(z_gt))
        code_path = os.path.join(self.load_path or self.model_dir, 'code_out.npz')
        np.savez_compressed(code_path, z_out=np.stack(z_out_list), z_gt=np.stack(z_gt_list))
        return code_path

    # ------------------------------------------------------------------ checkpoint (state_dict with TF variable names)
    def save(self, path):
        sd = {"variables": self.engine.params.state_dict(), "step": self.step, "g_lr": self.g_lr,
              "adam_t": self.engine.adam_t, "adam_m": self.engine.params.m.cpu(), "adam_v": self.engine.params.v.cpu()}
        torch.save(sd, path)

    def save_tf(self, model_dir):
        """`self.saver.save(self.sess, model_dir/model.ckpt, global_step=self.step)` (trainer.py:291-292, :460-461;
        trainer3.py:178-179, :343-344): a TensorFlow tensor bundle with the reference's variable names and layouts, Adam's
        slots, `step` and `g_lr` -- readable by tf.train.Saver / tf.train.load_checkpoint.  Returns the prefix."""
        from . import tf_checkpoint as tfc
        P = self.engine.params
        tensors = {k: P.p(k).detach().cpu().numpy() for k in P.table}
        if self.optimizer == 'adam':
            slot_tab = type(P.table)((k, P.table[k]) for k in self._slot_variables())
            tensors.update(tfc.adam_state_to_tf(slot_tab, {k: P._view(P.m, k).cpu().numpy() for k in slot_tab},
                                                {k: P._view(P.v, k).cpu().numpy() for k in slot_tab},
                                                self.engine.adam_t, self.beta1, self.beta2))
        tensors["step"] = np.int32(self.step)
        tensors["g_lr"] = np.float32(self.g_lr)
        base = "model.ckpt-%d" % self.step
        tfc.write_checkpoint(os.path.join(model_dir, base), tensors)
        tfc.update_checkpoint_state(model_dir, base)
        return os.path.join(model_dir, base)

    def load_tf(self, prefix):
        """restore from a TensorFlow checkpoint prefix (`.../model.ckpt-<step>`): variables by name (shapes must match),
        Adam slots / step / g_lr when present (a checkpoint of weights only leaves the optimizer state fresh)."""
        from . import tf_checkpoint as tfc
        P = self.engine.params
        have = {n: shp for n, shp, _ in tfc.list_variables(prefix)}
        missing = [k for k in P.table if k not in have]
        if missing:
            raise KeyError("checkpoint %s lacks variables %s" % (prefix, missing[:4]))
        for k in P.table:
            if tuple(have[k]) != tuple(P.table[k]):
                raise ValueError("variable %s: checkpoint shape %s, model shape %s" % (k, tuple(have[k]), tuple(P.table[k])))
        slot_vars = self._slot_variables()
        slots = [k + sfx for k in slot_vars for sfx in ("/Adam", "/Adam_1")]
        with_adam = all(n in have for n in slots) and "beta1_power" in have
        extra = [n for n in ("step", "g_lr") if n in have]
        powers = [n for n in ("beta1_power", "beta2_power") if n in have]
        t = tfc.read_checkpoint(prefix, list(P.table) + (slots + powers if with_adam else []) + extra)
        P.load_state_dict({k: torch.from_numpy(t[k]) for k in P.table})
        if with_adam:
            for k in slot_vars:
                P._view(P.m, k).copy_(torch.from_numpy(t[k + "/Adam"]).view(*P.table[k]))
                P._view(P.v, k).copy_(torch.from_numpy(t[k + "/Adam_1"]).view(*P.table[k]))
            self.engine.adam_t = tfc.adam_t_from_tf(float(t["beta1_power"]), self.beta1,
                                                    float(t["beta2_power"]) if "beta2_power" in t else None, self.beta2)
        if "step" in t:
            self.step = int(t["step"])
        if "g_lr" in t:
            self.g_lr = float(t["g_lr"])
        self.engine.repack()

    def load(self, path):
        sd = torch.load(path, map_location="cpu")
        self.engine.params.load_state_dict(sd["variables"])
        self.engine.params.m.copy_(sd["adam_m"])
        self.engine.params.v.copy_(sd["adam_v"])
        self.engine.adam_t = sd["adam_t"]
        self.step, self.g_lr = sd["step"], sd["g_lr"]
        self.engine.repack()
