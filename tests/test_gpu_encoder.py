"""GPU parity of the encoder / auto-encoder extensions (BASELINE config 5) against the CPU oracle:
channel-blocked inputs (Cin > 128), stride-2 convolutions on the generic per-tap tensor-core kernel (forward with TMA
element strides, data gradient by output parity class, weight gradient), the encoder FC, and the whole AE step.
Tolerances as in test_gpu_kernels.py: bf16 outputs rel-L2 <= 4e-3 vs the fp32 oracle on identical bf16 operands, fp32
weight gradients <= 1e-4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_model as M
from oracle import ref_ops as R
from oracle import ref_train as T


def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def to_blocks(x):
    """[B,..,C] -> channel-blocked [C/128 * B,..,128]"""
    B, C = x.shape[0], x.shape[-1]
    nb = C // 128
    xb = x.reshape(x.shape[:-1] + (nb, 128))
    perm = [xb.dim() - 2] + list(range(0, xb.dim() - 2)) + [xb.dim() - 1]
    return xb.permute(*perm).reshape((nb * B,) + tuple(x.shape[1:-1]) + (128,)).contiguous()


def from_blocks(xb, nb):
    B = xb.shape[0] // nb
    x = xb.reshape((nb, B) + tuple(xb.shape[1:]))
    perm = list(range(1, x.dim() - 1)) + [0, x.dim() - 1]
    return x.permute(*perm).reshape((B,) + tuple(xb.shape[1:-1]) + (nb * 128,)).contiguous()


@pytest.mark.parametrize("shape,nd,nb", [((2, 4, 8, 8), 3, 2), ((1, 6, 10, 9), 3, 3), ((2, 16, 12), 2, 5)])
def test_conv_channel_blocked_input(shape, nd, nb):
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(70)
    cin = nb * 128
    x = (torch.randn(*shape, cin, generator=g) * 0.5).bfloat16()
    w = R.xavier_uniform_((3,) * nd + (cin, 128), g).bfloat16()
    b = torch.randn(128, generator=g) * 0.1
    dy = (torch.randn(*shape, 128, generator=g) * 0.1).bfloat16()
    xin, wt = x.float().requires_grad_(True), w.float().requires_grad_(True)
    y = R.conv_nd(xin, wt, b, 1, R.lrelu)
    ypre = R.conv_nd(xin, wt, b, 1, None)
    gx, gw = torch.autograd.grad(ypre, [xin, wt], dy.float())
    wf, wd = K.pack_conv_weights(w.float().to(dev()))
    xb = to_blocks(x).to(dev())
    out = torch.empty(*shape, 128, dtype=torch.bfloat16, device=dev())
    K.conv3x3(xb, wf, b.to(dev()), out=out, flags=K.CONV_LRELU, nblk=nb)
    assert rel_l2(out.float(), y.detach()) <= 4e-3
    # dgrad block by block with the (v + residual) * lrelu'(mask) epilogue, wgrad block by block
    res = torch.randn(*shape, cin, generator=g).bfloat16()
    msk = torch.randn(*shape, cin, generator=g).bfloat16()
    resb, mskb = to_blocks(res).to(dev()), to_blocks(msk).to(dev())
    dxb = torch.empty_like(xb)
    dw = torch.zeros(3 ** nd, cin, 128, device=dev())
    B = shape[0]
    for ib in range(nb):
        K.conv3x3(dy.to(dev()), wd[ib * 128:(ib + 1) * 128], None, out2=dxb[ib * B:(ib + 1) * B],
                  residual=resb[ib * B:(ib + 1) * B], mask_src=mskb[ib * B:(ib + 1) * B], flags=K.CONV_MASK_AFTER_RESIDUAL)
        K.conv_wgrad_ex(xb[ib * B:(ib + 1) * B], dy.to(dev()), dw[0, ib * 128:], None, 1, 1, cin * 128, 128)
    ref_dx = (gx + res.float()) * torch.where(msk.float() >= 0, 1.0, 0.2)
    assert rel_l2(from_blocks(dxb.float().cpu(), nb), ref_dx) <= 4e-3
    assert rel_l2(dw.view(gw.shape), gw) <= 1e-4


@pytest.mark.parametrize("shape,nd,nb", [((2, 8, 8, 16), 3, 2), ((1, 4, 12, 20), 3, 3), ((2, 16, 24), 2, 2), ((1, 32, 24), 2, 3)])
def test_conv_stride2_fwd_dgrad_wgrad(shape, nd, nb):
    """ch -> ch, k3, stride 2, TF SAME (pad 0 before / 1 after on even sizes): model.py:140,176"""
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.encoder import _same_pad_before
    g = torch.Generator().manual_seed(71)
    ch = nb * 128
    B, fine = shape[0], list(shape[1:])
    coarse = [v // 2 for v in fine]
    taps = 3 ** nd
    x = (torch.randn(*shape, ch, generator=g) * 0.5).bfloat16()
    w = R.xavier_uniform_((3,) * nd + (ch, ch), g).bfloat16()
    b = torch.randn(ch, generator=g) * 0.1
    xin, wt, bt = x.float().requires_grad_(True), w.float().requires_grad_(True), b.clone().requires_grad_(True)
    y = R.conv_nd(xin, wt, bt, 2, R.lrelu)
    ypre = R.conv_nd(xin, wt, bt, 2, None)
    dy = (torch.randn(ypre.shape, generator=g) * 0.1).bfloat16()
    gx, gw, gb = torch.autograd.grad(ypre, [xin, wt, bt], dy.float())
    wf, wd = K.pack_conv_weights(w.float().to(dev()))
    xb = to_blocks(x).to(dev())
    pb = [_same_pad_before(n) for n in fine]
    # ---- forward
    tl = []
    for t in range(taps):
        tt = [(t // 3 ** (nd - 1 - a)) % 3 for a in range(nd)]
        tl.append(([0] * (3 - nd) + [tt[a] - pb[a] for a in range(nd)]) + [t * ch])
    yb = torch.empty([nb * B] + coarse + [128], dtype=torch.bfloat16, device=dev())
    bd = b.to(dev())
    for ob in range(nb):
        K.conv_taps(xb, wf[ob * 128:(ob + 1) * 128], bd[ob * 128:(ob + 1) * 128], yb[ob * B:(ob + 1) * B], None, None, None,
                    [B] + coarse, coarse, ch, 2, tl, 1, [0] * nd, flags=K.CONV_LRELU)
    assert rel_l2(from_blocks(yb.float().cpu(), nb), y.detach()) <= 4e-3
    # ---- data gradient by parity class
    dyb = to_blocks(dy).to(dev())
    dxb = torch.zeros_like(xb)
    for r in range(2 ** nd):
        rr = [(r >> (nd - 1 - a)) & 1 for a in range(nd)]
        tl = []
        for t in range(taps):
            tt = [(t // 3 ** (nd - 1 - a)) % 3 for a in range(nd)]
            if any((rr[a] - tt[a] + pb[a]) % 2 for a in range(nd)):
                continue
            tl.append(([0] * (3 - nd) + [(rr[a] - tt[a] + pb[a]) // 2 for a in range(nd)]) + [(taps - 1 - t) * ch])
        for ib in range(nb):
            K.conv_taps(dyb, wd[ib * 128:(ib + 1) * 128], None, dxb[ib * B:(ib + 1) * B], None, None, None, [B] + coarse,
                        fine, ch, 1, tl, 2, rr)
    assert rel_l2(from_blocks(dxb.float().cpu(), nb), gx) <= 4e-3
    # ---- weight + bias gradient
    dw = torch.zeros(taps, ch, ch, device=dev())
    db = torch.zeros(ch, device=dev())
    for ib in range(nb):
        for ob in range(nb):
            K.conv_wgrad_ex(xb[ib * B:(ib + 1) * B], dyb[ob * B:(ob + 1) * B], dw[0, ib * 128:, ob * 128:],
                            db[ob * 128:(ob + 1) * 128] if ib == 0 else None, 2, pb[0], ch * ch, ch)
    assert rel_l2(dw.view(gw.shape), gw) <= 1e-4
    assert rel_l2(db, gb) <= 1e-4


def test_encoder_glue_kernels():
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(72)
    # pad_cast
    x = torch.randn(2, 3, 4, 5, 3, generator=g)
    out = torch.empty(2, 3, 4, 5, 128, dtype=torch.bfloat16, device=dev())
    K.pad_cast(x.to(dev()), out)
    ref = torch.zeros(2, 3, 4, 5, 128)
    ref[..., :3] = x.bfloat16().float()
    assert torch.equal(out.float().cpu(), ref)
    # add_mask
    a, b, y = (torch.randn(3, 5, 7, 128, generator=g).bfloat16() for _ in range(3))
    o = torch.empty_like(a, device=dev())
    K.add_mask(a.to(dev()), b.to(dev()), y.to(dev()), o)
    refm = ((a.float() + b.float()) * torch.where(y.float() >= 0, 1.0, 0.2)).bfloat16()
    assert torch.equal(o.cpu(), refm)
    K.add_mask(a.to(dev()), None, None, o)
    assert torch.equal(o.cpu(), a)
    # encoder FC forward / backward on a channel-blocked tensor
    B, nb, sp, Z = 3, 3, (2, 2, 3), 16
    flat = (torch.randn(B, *sp, nb * 128, generator=g) * 0.5).bfloat16()
    F = int(np.prod(sp)) * nb * 128
    W = R.xavier_uniform_((F, Z), g)
    bias = torch.randn(Z, generator=g)
    fl = flat.float().reshape(B, -1).requires_grad_(True)
    Wt = W.clone().requires_grad_(True)
    zref = fl @ Wt + bias
    dz = torch.randn(B, Z, generator=g)
    gfl, gW = torch.autograd.grad(zref, [fl, Wt], dz)
    fb = to_blocks(flat).to(dev())
    z = torch.empty(B, Z, device=dev())
    K.enc_fc_fwd(fb, W.to(dev()), bias.to(dev()), z, nb)
    assert rel_l2(z, zref.detach()) <= 1e-5
    dW, db, dfl = torch.empty(F, Z, device=dev()), torch.empty(Z, device=dev()), torch.empty_like(fb)
    K.enc_fc_bwd(fb, W.to(dev()), dz.to(dev()), dW, db, dfl, nb)
    assert rel_l2(dW, gW) <= 1e-5 and rel_l2(db, dz.sum(0)) <= 1e-6
    assert rel_l2(from_blocks(dfl.float().cpu(), nb).reshape(B, -1), gfl) <= 4e-3
    # decoder FC input gradient + loss_p
    Kd, N = 16, 4096
    dout = torch.randn(B, N, generator=g).bfloat16()
    Wd = torch.randn(Kd, N, generator=g)
    dzd = torch.ones(B, Kd, device=dev())
    K.fc_dz(dout.to(dev()), Wd.to(dev()), dzd, accumulate=True)
    assert rel_l2(dzd, 1.0 + dout.float() @ Wd.t()) <= 1e-5
    zz, yl = torch.randn(B, Z, generator=g), torch.randn(B, 2, generator=g)
    dzp, lp = torch.empty(B, Z, device=dev()), torch.empty(1, device=dev())
    K.ae_loss_p(zz.to(dev()), yl.to(dev()), dzp, lp, 0.7)
    zl = zz.clone().requires_grad_(True)
    lref = ((yl - zl[:, -2:]) ** 2).mean()
    (gz,) = torch.autograd.grad(0.7 * lref, zl)
    assert abs(lp.item() - lref.item()) <= 1e-6 and rel_l2(dzp, gz) <= 1e-6


@pytest.mark.parametrize("spatial,nc", [([16, 16, 16], 2), ([32, 24], 3), ([8, 16, 16], 1)])
def test_encoder_forward_vs_oracle(spatial, nc):
    from deepfluids_b200.encoder import EncoderEngine
    nd = len(spatial)
    cin = 3 if nd == 3 else 2
    B = 2
    enc = EncoderEngine(B, spatial + [cin], z_num=16, num_conv=nc, name="AE/enc", device=dev(), seed=9)
    var = enc.params.state_dict()
    assert list(var.keys()) == list(M.encoder_layout(spatial + [cin], num_conv=nc, name="AE/enc")[0].keys())
    x, _ = T.synthetic_batch(B, spatial, seed=4)
    z = enc.forward(x.to(dev()))
    zref = M.encoder_forward(x, var, num_conv=nc, name="AE/enc")
    assert rel_l2(z, zref) <= 2e-2, rel_l2(z, zref)


def test_ae_step_vs_oracle():
    """whole AE3 forward/backward (trainer3.py:240-279) vs the fp32 oracle: outputs tight, gradients within the
    free-running bf16 bound (see test_gpu_trainstep.py for why a free-running gradient comparison is loose)."""
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.encoder import AEEngine
    spatial, B, nc = [16, 16, 16], 2, 2
    ae = AEEngine(B, spatial + [3], z_num=16, num_conv=nc, device=dev(), seed=3)
    var = ae.params.state_dict()
    assert list(var.keys()) == list(M.ae_layout(spatial + [3], num_conv=nc).keys())
    x, _ = T.synthetic_batch(B, spatial, seed=6)
    ylast = torch.rand(B, 2, generator=torch.Generator().manual_seed(1)) * 2 - 1
    total, l1, jl1, lp, g, zref, grads = T.ae_loss_and_grads(x, ylast, var, 2, num_conv=nc)
    ae.zero_grad()
    pot, z = ae.forward(x.to(dev()))
    loss3, dpot, _ = K.stencil_loss_fwdbwd(pot, x.to(dev()))
    lpd = torch.empty(1, device=dev())
    K.ae_loss_p(z, ylast.to(dev()), ae.dz, lpd, 1.0)
    ae.backward(dpot)
    assert rel_l2(z, zref) <= 2e-2
    assert abs(lpd.item() - lp.item()) <= 2e-2 * abs(lp.item())
    assert abs(loss3[0].item() + lpd.item() - total.item()) <= 1e-2 * abs(total.item())
    errs = {k: rel_l2(ae.params.g(k), grads[k]) for k in var if not k.endswith("biases")}
    worst = max(errs, key=errs.get)
    print("AE worst weight-grad rel-L2 %.3e (%s)" % (errs[worst], worst))
    assert errs[worst] <= 1.5e-1, (worst, errs[worst])      # free-running (sign flips, see test_gpu_trainstep); teacher-forced: test_gpu_baseline_sizes
    for k in var:
        assert torch.isfinite(ae.params.g(k)).all(), k


def test_ae2d_step_vs_oracle():
    """2D AE with use_curl (trainer.py:357-396): the decoder emits 2 channels, curl reads channel 0 only"""
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.encoder import AEEngine
    spatial, B, nc = [32, 24], 2, 2
    ae = AEEngine(B, spatial + [2], z_num=16, num_conv=nc, device=dev(), seed=5)
    var = ae.params.state_dict()
    assert list(var.keys()) == list(M.ae_layout(spatial + [2], num_conv=nc).keys())
    x, _ = T.synthetic_batch(B, spatial, seed=8)
    ylast = torch.rand(B, 2, generator=torch.Generator().manual_seed(2)) * 2 - 1
    total, l1, jl1, lp, g, zref, grads = T.ae_loss_and_grads(x, ylast, var, 2, num_conv=nc)
    ae.zero_grad()
    pot, z = ae.forward(x.to(dev()))
    dpot = torch.empty_like(pot)
    loss3, _, _ = K.stencil_loss_fwdbwd(pot, x.to(dev()), dpot=dpot)
    assert float(dpot[..., 1].abs().max()) == 0.0
    lpd = torch.empty(1, device=dev())
    K.ae_loss_p(z, ylast.to(dev()), ae.dz, lpd, 1.0)
    ae.backward(dpot)
    assert rel_l2(z, zref) <= 2e-2
    assert abs(loss3[0].item() + lpd.item() - total.item()) <= 1e-2 * abs(total.item())
    errs = {k: rel_l2(ae.params.g(k), grads[k]) for k in var if k.endswith("weights")}
    worst = max(errs, key=errs.get)
    assert errs[worst] <= 1.5e-1, (worst, errs[worst])      # free-running (sign flips, see test_gpu_trainstep); teacher-forced: test_gpu_baseline_sizes


def test_trainer_ae_api_runs_and_loss_decreases():
    from deepfluids_b200 import config as C
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer3 import Trainer3
    cfg, _ = C.get_config(["--synthetic=true", "--arch=ae", "--is_3d=true", "--res_x=16", "--res_y=16", "--res_z=16",
                           "--batch_size=2", "--num_conv=2", "--max_step=20", "--lr_max=0.0005"])
    bm = BatchManager(cfg, pool=1)
    tr = Trainer3(cfg, bm)
    first = None
    for i in range(20):
        tr.train_step()
        tr.update_lr(i)
        if i == 0:
            first = tr.losses_ae()[0]
    last = tr.losses_ae()[0]
    assert np.isfinite(last) and last < first, (first, last)


def test_ae_use_sparse_vs_oracle():
    """use_sparse=True (config.py:29): sigmoid on z + Bernoulli-KL sparsity term (trainer.py:389-394)"""
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.encoder import AEEngine
    spatial, B, nc = [16, 16, 16], 2, 2
    ae = AEEngine(B, spatial + [3], z_num=16, num_conv=nc, device=dev(), seed=4, use_sparse=True)
    var = ae.params.state_dict()
    x, _ = T.synthetic_batch(B, spatial, seed=6)
    ylast = torch.rand(B, 2, generator=torch.Generator().manual_seed(1))
    total, l1, jl1, lp, g, zref, grads = T.ae_loss_and_grads(x, ylast, var, 2, num_conv=nc, use_sparse=True,
                                                             sparsity=0.05, w5=0.5)
    ae.zero_grad()
    pot, z = ae.forward(x.to(dev()))
    loss3, dpot, _ = K.stencil_loss_fwdbwd(pot, x.to(dev()))
    lpd = torch.empty(1, device=dev())
    K.ae_loss_p(z, ylast.to(dev()), ae.dz, lpd, 1.0)
    ae.backward(dpot, 2, 0.05, 0.5)
    assert rel_l2(z, zref) <= 2e-2 and float(z.min()) > 0 and float(z.max()) < 1
    tot = loss3[0].item() + lpd.item() + 0.5 * ae.loss_kl.item()
    assert abs(tot - total.item()) <= 1e-2 * abs(total.item())
    k = "AE/enc/%d_fc/weights" % (len([v for v in var if "/enc/" in v]) // 2 - 1)
    assert k in var and rel_l2(ae.params.g(k), grads[k]) <= 1e-1


def test_ae_encode_decode_inference():
    """decode-from-z entry (trainer.py:551-552) and latent dump (trainer.py:504) reuse the training engines"""
    from deepfluids_b200 import config as C, kernels as K
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer3 import Trainer3
    cfg, _ = C.get_config(["--synthetic=true", "--arch=ae", "--is_3d=true", "--res_x=16", "--res_y=16", "--res_z=16",
                           "--batch_size=2", "--num_conv=2", "--max_step=4"])
    bm = BatchManager(cfg, pool=2)
    tr = Trainer3(cfg, bm)
    x = torch.cat([bm._pool[0][0], bm._pool[1][0][:1]])           # 3 fields: not a multiple of the batch
    z = tr.encode(x)
    assert z.shape == (3, cfg.z_num) and torch.isfinite(z).all()
    _, z2 = tr.ae.forward(x[:2].contiguous())
    # (the encoder FC reduces over blocks with fp32 atomics: run-to-run differences in the last bits of z)
    assert rel_l2(z[:2], z2) <= 1e-5
    v = tr.decode(z)
    assert v.shape == x.shape and float(K.divergence(v).abs().max()) <= 1e-5
    # decode is a pure function of z: the chunked / padded path equals a direct decoder call on the same codes
    direct = K.curl_fwd(tr.ae.dec.forward(z[:2].contiguous()))
    assert torch.equal(v[:2], direct)
    assert torch.equal(tr.decode(z), v)


def test_conv_taps_paired_bricks_with_an_odd_brick_count():
    """the per-tap kernel pairs bricks (two per schedule step share every weight tile) when a launch has >= 2 bricks per SM; an
    odd count leaves a ghost half in the last pair.  297 bricks of 1 x 8 x 16 voxels, explicit 27-tap list, against the
    tap-window kernel (itself checked against the oracle) on the same operands: same values up to the fp32 summation order."""
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(72)
    shape = (1, 11, 72, 48)
    x = (torch.randn(*shape, 128, generator=g) * 0.5).bfloat16().to(dev())
    w = R.xavier_uniform_((3, 3, 3, 128, 128), g).bfloat16()
    b = (torch.randn(128, generator=g) * 0.1).to(dev())
    wf, _ = K.pack_conv_weights(w.float().to(dev()))
    ref = torch.empty(*shape, 128, dtype=torch.bfloat16, device=dev())
    K.conv3x3(x, wf, b, out=ref, flags=K.CONV_LRELU)
    tl = [[t // 9 - 1, (t // 3) % 3 - 1, t % 3 - 1, t * 128] for t in range(27)]
    out = torch.full(ref.shape, float("nan"), dtype=torch.bfloat16, device=dev())
    K.conv_taps(x, wf, b, out, None, None, None, list(shape), list(shape[1:]), 128, 1, tl, 1, [0, 0, 0], flags=K.CONV_LRELU)
    assert not torch.isnan(out.float()).any()
    assert rel_l2(out.float(), ref.float()) <= 2e-3
    # the last brick (the ghost's partner) and the first one, element-wise: within one bf16 ulp of each other
    for a, r_ in ((out[0, -1, -8:, -16:].float(), ref[0, -1, -8:, -16:].float()), (out[0, 0, :8, :16].float(), ref[0, 0, :8, :16].float())):
        assert float((a - r_).abs().max()) <= 2.0 ** -7 * float(r_.abs().max())
