"""GPU parity of the ops-level boundary (reference signatures, differentiable) and of arch=dg against the CPU oracle.

  * stencil adjoints dfl_curl_bwd / dfl_jacobian_bwd vs torch autograd of the oracle's (reference-pinned) curl / jacobian;
  * ops.conv2d / conv3d / linear with the reference's argument lists: forward and ALL gradients vs oracle autograd on the
    same bf16-rounded operands, for every layer shape of the patch discriminator (3|6 -> 64 s2, 64 -> 128 s2, 128 -> 256 s2,
    256 -> 512 s1, 512 -> 1 s1);
  * model.DiscriminatorPatch(3) vs oracle.ref_model.discriminator_forward (pinned against the reference's model.py by
    oracle/make_golden_model.py), variable names D/Conv ... D/Conv_4;
  * one arch=dg train step (trainer.py:149-156,174-184): every loss term, D outputs, discriminator and generator
    gradients vs oracle.ref_train.dg_losses_and_grads (pinned against the reference's build_model by
    oracle/make_golden_trainer.py).
Tolerances: stencil adjoints 1e-6 relative to max|g| (fp32, different summation order); single layers rel-L2 4e-3 forward /
dgrad, 1e-3 wgrad (fp32 accumulation of bf16 products; outputs stored in bf16); discriminator chain 2e-2; dg step: losses
2e-2 relative, gradients free-running rel-L2 <= 1.5e-1 (see test_gpu_trainstep.py on lrelu sign flips).
"""
from collections import OrderedDict

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_model as M
from oracle import ref_ops as R
from oracle import ref_train as T


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def dev():
    return torch.device("cuda:0")


def bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("shape", [[2, 9, 14], [1, 2, 2], [2, 6, 5, 8], [1, 2, 2, 2], [1, 16, 12, 10]])
def test_stencil_adjoints_vs_oracle_autograd(shape):
    from deepfluids_b200 import kernels as K
    nd = len(shape) - 1
    g = torch.Generator().manual_seed(sum(shape))
    if nd == 2:
        pot = torch.randn(shape + [2], generator=g)                 # curl reads channel 0 of a 2-channel tensor
        p = pot.clone().requires_grad_(True)
        vel = R.curl(p)
    else:
        pot = torch.randn(shape + [3], generator=g)
        p = pot.clone().requires_grad_(True)
        vel = R.curl3(p)
    gv = torch.randn(vel.shape, generator=g)
    (gp,) = torch.autograd.grad(vel, p, gv)
    mine = K.curl_bwd(gv.to(dev()), pot.shape[-1])
    assert mine.shape == pot.shape
    assert float((mine.cpu() - gp).abs().max()) <= 1e-6 * float(gp.abs().max())
    v = torch.randn(shape + [nd], generator=g)
    vl = v.clone().requires_grad_(True)
    j, a = (R.jacobian if nd == 2 else R.jacobian3)(vl)
    gj, ga = torch.randn(j.shape, generator=g), torch.randn(a.shape, generator=g)
    for use_j, use_a in ((True, True), (True, False), (False, True)):
        outs = [t for t, u in ((j, use_j), (a, use_a)) if u]
        seeds = [t for t, u in ((gj, use_j), (ga, use_a)) if u]
        (ref,) = torch.autograd.grad(outs, vl, seeds, retain_graph=True)
        got = K.jacobian_bwd(gj.to(dev()) if use_j else None, ga.to(dev()) if use_a else None)
        assert float((got.cpu() - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), (use_j, use_a)


def test_ops_stencils_are_differentiable_graph_nodes():
    """ops.curl / ops.jacobian / ops.jacobian3 inside an autograd graph == the oracle's graph (forward bit-exact)"""
    from deepfluids_b200 import ops as O
    g = torch.Generator().manual_seed(5)
    pot = torch.randn(2, 12, 10, 1, generator=g)
    x = torch.randn(2, 12, 10, 2, generator=g)
    pd = pot.to(dev()).requires_grad_(True)
    vel = O.curl(pd)
    j, w = O.jacobian(vel)
    loss = (vel - x.to(dev())).abs().mean() + (j * j).mean() + w.abs().mean()
    (gd,) = torch.autograd.grad(loss, pd)
    pc = pot.clone().requires_grad_(True)
    velc = R.curl(pc)
    jc, wc = R.jacobian(velc)
    lossc = (velc - x).abs().mean() + (jc * jc).mean() + wc.abs().mean()
    (gc,) = torch.autograd.grad(lossc, pc)
    assert torch.equal(vel.detach().cpu(), velc.detach()) and torch.equal(j.detach().cpu(), jc.detach())
    assert float((gd.cpu() - gc).abs().max()) <= 1e-6 * float(gc.abs().max())
    a3 = torch.randn(1, 6, 8, 10, 3, generator=g)
    ad = a3.to(dev()).requires_grad_(True)
    _, c = O.jacobian3(ad)                              # trainer3.py:18: `_, G_ = jacobian3(G_s)`
    (g3,) = torch.autograd.grad((c * c).sum(), ad)
    ac = a3.clone().requires_grad_(True)
    (g3c,) = torch.autograd.grad((R.curl3(ac) ** 2).sum(), ac)
    assert float((g3.cpu() - g3c).abs().max()) <= 1e-5 * float(g3c.abs().max())


LAYERS = [  # nd, spatial, Cin, Cout, stride, lrelu
    (2, [16, 12], 3, 64, 2, True), (2, [16, 12], 64, 128, 2, True), (2, [8, 12], 128, 256, 2, True),
    (2, [8, 6], 256, 512, 1, True), (2, [8, 6], 512, 1, 1, False), (2, [16, 16], 128, 128, 1, True),
    (3, [8, 8, 16], 6, 64, 2, True), (3, [8, 4, 8], 128, 256, 2, True), (3, [4, 4, 6], 256, 512, 1, True),
    (3, [4, 4, 6], 512, 1, 1, False), (3, [8, 8, 8], 128, 3, 1, False),
]


@pytest.mark.parametrize("nd,spatial,cin,cout,stride,lre", LAYERS)
def test_ops_conv_layers_forward_and_gradients(nd, spatial, cin, cout, stride, lre):
    """ops.conv2d / conv3d(x, o_dim, data_format, name, k, s, act): reference call signature, variables created in the
    scope store, y / dx / dW / db vs oracle autograd on the same bf16-rounded operands"""
    from deepfluids_b200 import ops as O
    O.reset_variables(11)
    B = 2
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = bf(torch.randn([B] + spatial + [cin], generator=g))
    conv = O.conv2d if nd == 2 else O.conv3d
    xd = x.to(dev()).requires_grad_(True)
    with O.variable_scope("T"):
        y = conv(xd, cout, name="c0", k=3, s=stride, act=O.lrelu if lre else None)
    assert O.get_variables("T") == ["T/c0/weights", "T/c0/biases"]
    w, b = O.get_variable("T/c0/weights"), O.get_variable("T/c0/biases")
    assert tuple(w.shape) == (3,) * nd + (cin, cout)
    with torch.no_grad():
        b.copy_(torch.randn(cout, generator=g).to(dev()) * 0.1)
    with O.variable_scope("T", reuse=True):
        y = conv(xd, cout, name="c0", k=3, s=stride, act=O.lrelu if lre else None)
    gy = bf(torch.randn(y.shape, generator=g))
    gx, gw, gb = torch.autograd.grad(y, [xd, w, b], gy.to(dev()).to(y.dtype))
    xc = x.clone().requires_grad_(True)
    wc = bf(w.detach().cpu()).requires_grad_(True)
    bc = b.detach().cpu().clone().requires_grad_(True)
    yc = R.conv_nd(xc, wc, bc, stride, R.lrelu if lre else None)
    gxc, gwc, gbc = torch.autograd.grad(yc, [xc, wc, bc], gy)
    assert y.shape == yc.shape
    e = dict(y=rel_l2(y, yc), dx=rel_l2(gx, gxc), dw=rel_l2(gw, gwc), db=rel_l2(gb, gbc))
    print(nd, spatial, cin, cout, stride, e)
    assert e["y"] <= 4e-3 and e["dx"] <= 6e-3 and e["dw"] <= 4e-3 and e["db"] <= 4e-3, e


def test_ops_conv_rejects_what_the_kernels_do_not_implement():
    from deepfluids_b200 import ops as O
    O.reset_variables()
    x = torch.zeros(1, 8, 8, 128, device=dev())
    with pytest.raises(NotImplementedError, match="k=4"):
        O.conv2d(x, 128)                                    # the reference's defaults k=4, s=2 have no call site on this path
    with pytest.raises(NotImplementedError, match="activation"):
        O.conv2d(x, 128, k=3, s=1, act=torch.tanh)
    with pytest.raises(NotImplementedError, match="channels-last"):
        O.conv2d(x, 128, data_format='NCHW', k=3, s=1)


def test_ops_linear_forward_and_gradients():
    from deepfluids_b200 import ops as O
    for (B, Kd, N) in ((4, 3, 6144), (7, 300, 130), (3, 20000, 16)):        # generator FC, an MLP layer, an encoder-style FC (split-K)
        O.reset_variables(3)
        g = torch.Generator().manual_seed(B)
        x = torch.randn(B, Kd, generator=g)
        xd = x.to(dev()).requires_grad_(True)
        y = O.linear(xd, N, name="fc")
        w, b = O.get_variable("fc/weights"), O.get_variable("fc/biases")
        gy = torch.randn(B, N, generator=g)
        gx, gw, gb = torch.autograd.grad(y, [xd, w, b], gy.to(dev()))
        xc, wc, bc = x.clone().requires_grad_(True), w.detach().cpu().clone().requires_grad_(True), b.detach().cpu().clone().requires_grad_(True)
        yc = R.linear(xc, wc, bc)
        gxc, gwc, gbc = torch.autograd.grad(yc, [xc, wc, bc], gy)
        e = (rel_l2(y, yc), rel_l2(gx, gxc), rel_l2(gw, gwc), rel_l2(gb, gbc))
        assert max(e) <= 2e-5, ((B, Kd, N), e)                      # fp32 arithmetic, different summation order


@pytest.mark.parametrize("spatial", [[32, 24], [16, 16, 16]])
def test_discriminator_patch_vs_oracle(spatial):
    from deepfluids_b200 import model as Mo, ops as O
    nd = len(spatial)
    cin = 3 if nd == 2 else 6
    O.reset_variables(21)
    g = torch.Generator().manual_seed(9)
    x = bf(torch.randn([2] + spatial + [cin], generator=g))
    fn = Mo.DiscriminatorPatch if nd == 2 else Mo.DiscriminatorPatch3
    out, names = fn(x.to(dev()), 128)
    tab = M.discriminator_layout(cin, 128, nd)
    assert names == list(tab.keys())
    var = OrderedDict((k, O.get_variable(k).detach().cpu()) for k in names)
    assert [tuple(v.shape) for v in var.values()] == [tuple(s) for s in tab.values()]
    ref = M.discriminator_forward(x, var, "D", store=M.bf16_round_ste)
    out2, names2 = fn(x.to(dev()), 128, reuse=True)
    assert names2 == names and torch.equal(out, out2)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel_l2(out, ref) <= 2e-2, rel_l2(out, ref)


@pytest.mark.parametrize("is3d", [False, True])
def test_dg_train_step_vs_oracle(is3d):
    """one `sess.run([g_optim, d_optim])` of arch=dg vs the oracle (itself pinned by the reference's build_model)"""
    from deepfluids_b200 import config as C, ops as O
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    from deepfluids_b200.trainer3 import Trainer3
    args = ["--synthetic=true", "--arch=dg", "--batch_size=2", "--num_conv=2", "--max_step=10", "--w3=0.5", "--lr_max=0.001"]
    args += ["--is_3d=true", "--res_x=16", "--res_y=16", "--res_z=16"] if is3d else ["--res_x=24", "--res_y=32"]
    cfg, _ = C.get_config(args)
    bm = BatchManager(cfg, pool=1)
    tr = (Trainer3 if is3d else Trainer)(cfg, bm)
    assert tr.D_var == list(M.discriminator_layout(6 if is3d else 3, 128, 3 if is3d else 2).keys())
    g_var = tr.engine.params.state_dict()
    d_var = OrderedDict((k, O.get_variable(k).detach().cpu().clone()) for k in tr.D_var)
    x, y = bm.batch()
    tr.train_step(x, y)
    got = tr.losses_dg()
    losses, gg, dg = T.dg_losses_and_grads(y.cpu(), x.cpu(), g_var, d_var, 128, 2, 0, cfg.w1, cfg.w2, cfg.w3)
    names = ("g_loss", "g_loss_l1", "g_loss_j_l1", "g_loss_real", "d_loss_fake", "d_loss_real", "d_loss")
    for k, v in zip(names, got):
        assert abs(v - float(losses[k])) <= 2e-2 * abs(float(losses[k])), (k, v, float(losses[k]))
    assert rel_l2(tr.D_x, losses["D_x"]) <= 2e-2 and rel_l2(tr.D_G, losses["D_G"]) <= 3e-2
    e_d = {k: rel_l2(g, dg[k]) for k, g in zip(tr.D_var, tr._d_grads) if k.endswith("weights")}
    print("dg %s: D grads %s" % ("3d" if is3d else "2d", {k: "%.1e" % v for k, v in e_d.items()}))
    assert max(e_d.values()) <= 1.5e-1, e_d
    # the optimizer moved both networks: Adam's first step is lr_t * sign-ish, so compare the direction of the update
    for k in ("D/Conv_1/weights", "D/Conv_3/weights"):
        upd = (O.get_variable(k).detach().cpu() - d_var[k])
        agree = float((torch.sign(upd) == -torch.sign(dg[k])).float().mean())
        assert agree >= 0.9, (k, agree)
    assert tr.engine.adam_t == 2                          # the shared optimizer's beta powers advance twice per step
    d0 = got[-1]
    for i in range(8):
        tr.train_step(x, y)
    after = tr.losses_dg()
    assert all(v == v and abs(v) < 1e6 for v in after) and after[-1] < d0, (d0, after)     # the discriminator learns to separate


# ---------------------------------------------------------------------------------------------------------------------
# reference flag values that used to hard-fail (round-1 review): --filters != 128, use_curl=False, skip_concat=True
# ---------------------------------------------------------------------------------------------------------------------
def _ops_vars(names):
    from deepfluids_b200 import ops as O
    return OrderedDict((k, O.get_variable(k).detach().cpu().clone()) for k in names)


@pytest.mark.parametrize("spatial,skip", [([32, 24], False), ([16, 16, 16], False), ([32, 24], True)])
def test_generator_filters64_and_skip_concat_vs_oracle(spatial, skip):
    """model.GeneratorBE(3) with filters=64 (run.bat:56,73 widths) and skip_concat=True (model.py:29-33): the ops-level
    network (zero-padded 128-channel blocks on the same kernels) vs the oracle's generator on the same variables."""
    from deepfluids_b200 import model as Mo, ops as O
    nd = len(spatial)
    cout = 3 if nd == 3 else 1
    O.reset_variables(5)
    g = torch.Generator().manual_seed(3)
    z = torch.rand(2, 3, generator=g) * 2 - 1
    fn = Mo.GeneratorBE3 if nd == 3 else Mo.GeneratorBE
    zd = z.to(dev())
    out, names = fn(zd, 64, spatial + [cout], num_conv=2, skip_concat=skip)
    var = _ops_vars(names)
    if skip:
        # oracle restatement of the skip_concat branch (model.py:29-33): upsample both, concatenate, no residual add
        x = R.linear(z, var["G/0_fc/weights"], var["G/0_fc/biases"]).reshape(2, 8, 6, 64)
        x0, n = x, 1
        for idx in range(3):
            for _ in range(2):
                x = R.conv_nd(x, var["G/%d_conv/weights" % n], var["G/%d_conv/biases" % n], 1, R.lrelu)
                n += 1
            if idx < 2:
                x, x0 = R.upscale(x, 2), R.upscale(x0, 2)
                x = torch.cat([x, x0], -1)
        ref = R.conv_nd(x, var["G/%d_conv/weights" % n], var["G/%d_conv/biases" % n], 1, None)
        # x0 is only up-sampled, never replaced (model.py:31-33): every later block's first conv sees filters + filters channels
        assert var["G/3_conv/weights"].shape[-2] == 128 and var["G/5_conv/weights"].shape[-2] == 128 and var["G/7_conv/weights"].shape[-2] == 64
    else:
        assert list(var.keys()) == list(M.generator_layout(spatial + [cout], 64, 2)[0].keys())
        ref = M.generator_forward(z, var, spatial + [cout], 64, 2)
    assert out.shape == ref.shape and rel_l2(out, ref) <= 2e-2, rel_l2(out, ref)
    # gradient w.r.t. every variable through the autograd tape of kernels
    gy = torch.randn(ref.shape, generator=g)
    O_vars = [O.get_variable(k) for k in names]
    out2, _ = fn(zd, 64, spatial + [cout], num_conv=2, skip_concat=skip, reuse=True)
    grads = torch.autograd.grad(out2, O_vars, gy.to(dev()).to(out2.dtype))
    assert all(torch.isfinite(t).all() for t in grads)
    if not skip:
        leaves = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in var.items())
        refg = torch.autograd.grad(M.generator_forward(z, leaves, spatial + [cout], 64, 2), list(leaves.values()), gy)
        errs = {k: rel_l2(a, b) for k, a, b in zip(names, grads, refg) if k.endswith("weights")}
        assert max(errs.values()) <= 1.5e-1, errs                           # free-running bf16 (see test_gpu_trainstep.py)


@pytest.mark.parametrize("is3d", [False, True])
def test_trainer_filters64_ae_recipe_vs_oracle(is3d):
    """run.bat:56,73: the AE recipes train with --filters=64.  One train step through the reference-style Trainer on the
    ops-level engine: latent code, losses and the Adam direction vs the oracle; then the loss must fall."""
    from deepfluids_b200 import config as C
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    from deepfluids_b200.trainer3 import Trainer3
    args = ["--synthetic=true", "--arch=ae", "--filter=64", "--batch_size=2", "--num_conv=2", "--max_step=20", "--lr_max=0.0005"]
    args += ["--is_3d=true", "--res_x=16", "--res_y=16", "--res_z=16"] if is3d else ["--res_x=24", "--res_y=32"]
    cfg, _ = C.get_config(args)
    assert cfg.filters == 64
    bm = BatchManager(cfg, pool=1)
    tr = (Trainer3 if is3d else Trainer)(cfg, bm)
    spatial = [16, 16, 16] if is3d else [32, 24]
    cin = 3 if is3d else 2
    assert tr.var == list(M.ae_layout(spatial + [cin], 64, 16, 2).keys())
    var = tr.engine.params.state_dict()
    x, y = bm.batch()
    ylast = y[:, :, -1] if y.dim() == 3 else y[:, -tr.p_num:]
    total, l1, jl1, lp, g, zref, grads = T.ae_loss_and_grads(x.cpu(), ylast.cpu(), var, tr.p_num, 64, 16, 2)
    tr.train_step_ae(x, y)
    got = tr.losses_ae()
    assert rel_l2(tr.z, zref) <= 2e-2
    assert abs(got[0] - float(total)) <= 2e-2 * abs(float(total)), (got, float(total))
    new = tr.engine.params.state_dict()
    for k in ("AE/enc/1_conv/weights", "AE/dec/2_conv/weights"):
        agree = float((torch.sign(new[k] - var[k]) == -torch.sign(grads[k])).float().mean())
        assert agree >= 0.85, (k, agree)
    first = got[0]
    for i in range(15):
        tr.train_step_ae(x, y)
    assert tr.losses_ae()[0] < first


@pytest.mark.parametrize("is3d", [False, True])
def test_use_curl_false_vs_oracle(is3d):
    """use_curl=False (trainer.py:141-144): the generator emits the velocity itself (2 / 3 channels); loss and its gradient
    w.r.t. the output from the un-fused kernels vs oracle autograd, then one trainer step end to end."""
    from deepfluids_b200 import config as C, kernels as K
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    from deepfluids_b200.trainer3 import Trainer3
    spatial = [12, 10, 14] if is3d else [18, 14]
    nd = len(spatial)
    g = torch.Generator().manual_seed(7)
    vel = torch.randn([2] + spatial + [nd], generator=g)
    x = torch.randn([2] + spatial + [nd], generator=g)
    loss3, dvel = K.velocity_loss_fwdbwd(vel.to(dev()), x.to(dev()), 0.7, 1.3)
    v = vel.clone().requires_grad_(True)
    loss, l1, jl1, _ = T.stencil_loss(v, x, 0.7, 1.3, use_curl=False)
    (gref,) = torch.autograd.grad(loss, v)
    assert abs(loss3[0].item() - loss.item()) <= 3e-6 * abs(loss.item())
    assert abs(loss3[1].item() - l1.item()) <= 3e-6 * abs(l1.item()) and abs(loss3[2].item() - jl1.item()) <= 3e-6 * abs(jl1.item())
    assert float((dvel.cpu() - gref).abs().max()) <= 1e-6 * float(gref.abs().max())
    args = ["--synthetic=true", "--use_curl=false", "--batch_size=2", "--num_conv=2", "--max_step=20", "--lr_max=0.0005"]
    args += ["--is_3d=true", "--res_x=16", "--res_y=16", "--res_z=16"] if is3d else ["--res_x=24", "--res_y=32"]
    cfg, _ = C.get_config(args)
    bm = BatchManager(cfg, pool=1)
    tr = (Trainer3 if is3d else Trainer)(cfg, bm)
    assert tr.output_shape[-1] == (3 if is3d else 2)                       # trainer.py:54-55
    var = tr.engine.params.state_dict()
    xb, yb = bm.batch()
    ref = T.generator_loss_and_grads(yb.cpu(), xb.cpu(), var, num_conv=2, use_curl=False)
    tr.train_step(xb, yb)
    got = tr.losses()
    assert abs(got[0] - float(ref[0])) <= 1e-2 * abs(float(ref[0])), (got, float(ref[0]))
    first = got[0]
    for i in range(15):
        tr.train_step(xb, yb)
    assert tr.losses()[0] < first
    vgen = tr.generate_velocity(yb)
    assert vgen.shape == xb.shape


def test_gradient_accumulation_equals_one_large_batch():
    """--grad_accum (strong scaling at a fixed global batch): 2 micro-batches of 2 == the gradient of one batch of 4"""
    from deepfluids_b200 import config as C
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    base = ["--synthetic=true", "--res_x=24", "--res_y=32", "--num_conv=2", "--max_step=50", "--lr_max=0.001"]
    cfg_a, _ = C.get_config(base + ["--batch_size=2", "--grad_accum=2"])
    bm_a = BatchManager(cfg_a, pool=2)
    tr_a = Trainer(cfg_a, bm_a)
    cfg_b, _ = C.get_config(base + ["--batch_size=4"])
    tr_b = Trainer(cfg_b, BatchManager(cfg_b, pool=1))
    assert torch.equal(tr_a.engine.params.data, tr_b.engine.params.data)
    (x0, y0), (x1, y1) = bm_a._pool[0], bm_a._pool[1]
    bm_a._i = 1                                      # the step draws micro-batch 0 itself, then pool[1]
    w0 = tr_a.engine.params.data.clone()
    tr_a.train_step(x0, y0)
    tr_b.train_step(torch.cat([x0, x1]), torch.cat([y0, y1]))
    torch.cuda.synchronize()
    la, lb = tr_a.losses(), tr_b.losses()
    assert abs(la[0] - lb[0]) <= 1e-5 * abs(lb[0]), (la, lb)
    ga, gb = tr_a.engine.params.grad / 2, tr_b.engine.params.grad
    assert rel_l2(ga, gb) <= 2e-3, rel_l2(ga, gb)    # (wgrad reduces with fp32 atomics: order differs between the two)
    da, db = tr_a.engine.params.data - w0, tr_b.engine.params.data - w0
    assert rel_l2(da, db) <= 5e-2 and tr_a.engine.adam_t == 1 and tr_a.step == 1


def test_test_ae_latent_dump_and_decode(tmp_path):
    """Trainer.test_ae (trainer.py:475-583): latent codes of the whole dataset in file order -> code<z>.npz with the
    reference's x / y / p / s / f arrays; then decode-from-code_out.npz."""
    import numpy as np
    from deepfluids_b200 import config as C
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    root = tmp_path / "data" / "toy_ae"
    (root / "v").mkdir(parents=True)
    sims, frames, H, W = 2, 4, 32, 24
    rng = np.random.default_rng(0)
    with open(root / "args.txt", "w") as f:               # AE scene layout (scene/smoke_mov.py-style): <scene>_<frame>.npz
        for k, v in [("num_param", 2), ("p0", "scenes"), ("p1", "frames"), ("min_scenes", 0), ("max_scenes", sims - 1),
                     ("num_scenes", sims), ("min_frames", 0), ("max_frames", frames - 1), ("num_frames", frames),
                     ("num_dof", 1), ("resolution_x", W), ("resolution_y", H)]:
            f.write("%s: %s\n" % (k, v))
    vr = 0.0
    for s_ in range(sims):
        for t in range(frames):
            x = rng.standard_normal((H, W, 2)).astype(np.float32)
            vr = max(vr, float(np.abs(x).max()))
            np.savez_compressed(root / "v" / ("%d_%d.npz" % (s_, t)), x=x, y=(rng.random((1, frames)) * 2 - 1).astype(np.float32))
    with open(root / "v_range.txt", "w") as f:
        f.write("%f\n%f" % (-vr, vr))
    nx = np.cumsum(rng.random((sims, frames)), axis=1)
    np.savez_compressed(root / "n.npz", nx=nx)
    args = ["--arch=ae", "--dataset=toy_ae", "--data_dir=%s" % (tmp_path / "data"), "--res_x=%d" % W, "--res_y=%d" % H,
            "--batch_size=2", "--test_batch_size=4", "--num_conv=2", "--max_step=2", "--num_worker=1"]
    cfg, _ = C.get_config(args)
    cfg.model_dir = str(tmp_path / "run")
    cfg.load_path = ""
    (tmp_path / "run").mkdir()
    bm = BatchManager(cfg)
    tr = Trainer(cfg, bm)
    tr.load_path = str(tmp_path / "run")
    path = tr.test_ae()
    d = np.load(path)
    assert d["x"].shape == (sims * (frames - 1), cfg.z_num) and d["y"].shape == d["x"].shape
    assert d["p"].shape == (sims * (frames - 1), 1) and int(d["s"]) == sims and int(d["f"]) == frames
    np.testing.assert_allclose(d["p"][:, 0], (nx[:, 1:] - nx[:, :-1]).reshape(-1), rtol=1e-6)
    # x / y are the SAME code sequence shifted by one frame inside each simulation
    np.testing.assert_array_equal(d["x"].reshape(sims, frames - 1, -1)[:, 1:], d["y"].reshape(sims, frames - 1, -1)[:, :-1])
    # and they are the encoder's codes of the files in path order
    xs = np.stack([np.load(p)["x"] for p in bm.paths]) / bm.x_range
    z = tr.encode(xs.astype(np.float32)).cpu().numpy().reshape(sims, frames, -1)
    np.testing.assert_allclose(d["x"], z[:, :-1].reshape(-1, cfg.z_num), rtol=1e-4, atol=1e-5)
    # decode branch
    cdir = tmp_path / "codes"
    cdir.mkdir()
    np.savez_compressed(cdir / "code_out.npz", z_out=z[:, :2], z_gt=z[:, 2:])
    tr.config.code_path = str(cdir)
    outs = tr.test_ae()
    assert len(outs) == sims
    v = np.load(outs[1])
    assert v["v"].shape == (2, H, W, 2) and v["v_gt"].shape == (2, H, W, 2) and np.isfinite(v["v"]).all()


@pytest.mark.parametrize("shape,dtype", [((2, 5, 7, 3), torch.float32), ((1, 4, 6, 64), torch.bfloat16), ((2, 3, 4, 5, 3), torch.float32),
                                         ((1, 2, 3, 4, 128), torch.bfloat16), ((1, 3, 3, 1), torch.float32), ((1, 2, 2, 2, 7), torch.bfloat16)])
def test_ops_upscale_forward_and_adjoint_vs_oracle(shape, dtype):
    """ops.upscale / ops.upscale3 (ops.py:66-91) as standalone differentiable layers on dfl_upscale2 / dfl_pool2: bit-exact
    copies forward; the adjoint = sum of the children, vs torch autograd through the oracle's (reference-pinned) upscale"""
    from deepfluids_b200 import ops as O
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).to(dtype)
    nd = len(shape) - 2
    ref_up = R.upscale3 if nd == 3 else R.upscale
    xd = x.to(dev()).requires_grad_(True)
    y = O.upscale3(xd, 2) if nd == 3 else O.upscale(xd, 2)
    assert y.dtype == dtype and torch.equal(y.detach().cpu(), ref_up(x, 2))
    dy = torch.randn(*y.shape, generator=g).to(dtype)
    (gx,) = torch.autograd.grad(y, xd, dy.to(dev()))
    xr = x.float().requires_grad_(True)
    (gr,) = torch.autograd.grad(ref_up(xr, 2), xr, dy.float())
    tol = 1e-6 if dtype == torch.float32 else 8e-3          # bf16: one rounding of the fp32 sum of 4 / 8 children
    assert float((gx.float().cpu() - gr).abs().max()) <= tol * float(gr.abs().max())
    with pytest.raises(NotImplementedError):
        O.upscale(xd, 3) if nd == 2 else O.upscale3(xd, 3)
