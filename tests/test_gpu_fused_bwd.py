"""GPU parity of the FUSED loss-stencil + output-conv backward kernel (dfl_lastconv_curl_loss_bwd, north_star's
"curl / velocity-gradient stencils fused into the last generator epilogue and the first backward prologue").

  * against the ORACLE: loss terms (fp64 evaluation), dL/dA vs torch autograd of the reference-pinned curl / jacobian3 / L1
    expression, G_ = curl(A) bit for bit, divergence <= 1e-5; then ds / dw / db vs oracle autograd of the output conv fed
    with that dL/dA;
  * against the un-fused kernel pair (dfl_stencil_loss_fwdbwd + dfl_lastconv_bwd): the same ds bit for bit (identical dL/dA
    planes -> identical im2col operands -> identical tensor-core sums), ds_masked within one bf16 ulp (see _one_ulp);
  * shapes: tiles that overhang H and W, D smaller than the pipeline depth, 148-way splits that start and end inside a
    column, the BASELINE sizes 16 x 64^3 (reduced batch) and 128^3;
  * the trainer's step with and without the fusion lands on the same weights.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_ops as R
from oracle import ref_train as T


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def dev():
    return torch.device("cuda:0")


def _inputs(shape, seed):
    g = torch.Generator().manual_seed(seed)
    B, D, H, W = shape
    pot = torch.randn(B, D, H, W, 3, generator=g) * 0.05
    pot = (pot + torch.roll(pot, 1, 1) + torch.roll(pot, 1, 2) + torch.roll(pot, 1, 3)) / 4
    x, _ = T.synthetic_batch(B, [D, H, W], seed=seed + 1, smooth=1)
    s = (torch.randn(B, D, H, W, 128, generator=g) * 0.5).bfloat16()
    mask = torch.randn(B, D, H, W, 128, generator=g).bfloat16()
    w = R.xavier_uniform_((3, 3, 3, 128, 3), g).bfloat16().float()
    return pot, x, s, mask, w


def _run_fused(pot, x, s, mask, w, w1=1.0, w2=1.0, want=True):
    from deepfluids_b200 import kernels as K
    d = dev()
    ds = torch.full(s.shape, float("nan"), dtype=torch.bfloat16, device=d)
    dsm = torch.full(s.shape, float("nan"), dtype=torch.bfloat16, device=d)
    dw = torch.zeros(w.shape, device=d)
    db = torch.zeros(3, device=d)
    loss3 = torch.zeros(3, device=d)
    dpot = torch.full(pot.shape, float("nan"), device=d) if want else None
    vel = torch.full(pot.shape, float("nan"), device=d) if want else None
    ws = K.lastconv_curl_loss_workspace(d)
    for _ in range(2):                       # twice: the CTA ticket must reset itself
        dw.zero_(); db.zero_()
        K.lastconv_curl_loss_bwd(s.to(d), pot.to(d), x.to(d), w.to(d), mask.to(d), ds, dsm, dw, db, loss3, ws, w1, w2, 1.0,
                                 dpot=dpot, vel=vel)
    torch.cuda.synchronize()
    return ds, dsm, dw, db, loss3, dpot, vel


def _one_ulp(a, b):
    """ds_masked: the fused kernel scales the bf16 image by 0.2 (two roundings), the un-fused one rounds 0.2 * v from fp32:
    equal up to one bf16 ulp, and equal outright in the large majority of the elements"""
    a, b = a.float(), b.float()
    ok = bool(((a - b).abs() <= b.abs() * 2.0 ** -7 + 1e-30).all())
    same = float((a == b).float().mean())
    return ok and same >= 0.85


SHAPES = [(1, 8, 8, 16), (2, 4, 6, 12), (1, 2, 2, 2), (1, 3, 20, 38), (2, 16, 24, 32), (1, 40, 20, 36), (2, 64, 64, 64)]


@pytest.mark.parametrize("shape", SHAPES)
def test_fused_bwd_vs_oracle_and_unfused_pair(shape):
    from deepfluids_b200 import kernels as K
    pot, x, s, mask, w = _inputs(shape, 11 + sum(shape))
    ds, dsm, dw, db, loss3, dpot, vel = _run_fused(pot, x, s, mask, w, 0.7, 1.3)
    # ---- oracle: loss (fp64), dL/dA (fp32 autograd: the sign pattern is an fp32 property), G_
    p = pot.clone().requires_grad_(True)
    loss, l1, jl1, vel_ref = T.stencil_loss(p, x, 0.7, 1.3)
    (gref,) = torch.autograd.grad(loss, p)
    with torch.no_grad():
        l64, l164, jl164, _ = T.stencil_loss(pot.double(), x.double(), 0.7, 1.3)
    assert torch.equal(vel.cpu(), vel_ref.detach())
    assert float(K.divergence(vel).abs().max()) <= 1e-5
    for got, want in zip(loss3.tolist(), (l64.item(), l164.item(), jl164.item())):
        assert abs(got - want) <= 3e-6 * abs(want), (loss3.tolist(), l64.item(), l164.item(), jl164.item())
    scale = float(gref.abs().max())
    assert float((dpot.cpu() - gref).abs().max()) <= 1e-6 * scale + 1e-12
    # ---- oracle: backward of the output conv fed with the bf16 copy of that gradient (the im2col operand is bf16)
    sin = s.float().requires_grad_(True)
    wt = w.clone().requires_grad_(True)
    bt = torch.zeros(3, requires_grad=True)
    y = R.conv_nd(sin, wt, bt, 1, None)
    gx, gw, gb = torch.autograd.grad(y, [sin, wt, bt], gref.bfloat16().float())
    assert rel_l2(ds.float(), gx) <= 4e-3
    assert rel_l2(dsm.float(), gx * torch.where(mask.float() >= 0, 1.0, 0.2)) <= 4e-3
    assert rel_l2(dw, gw) <= 2e-4
    assert float((db.cpu() - gb).abs().max()) <= 1e-3 * float(gref.abs().sum())       # (sums to ~0: compare absolutely)
    # ---- the un-fused kernel pair on the same inputs
    d = dev()
    l3u, dpu, velu = K.stencil_loss_fwdbwd(pot.to(d), x.to(d), 0.7, 1.3, want_vel=True)
    assert torch.equal(dpu, dpot) and torch.equal(velu, vel)
    ds_u, dsm_u = torch.empty_like(ds), torch.empty_like(dsm)
    dw_u, db_u = torch.zeros_like(dw), torch.zeros_like(db)
    K.lastconv_bwd(s.to(d), dpu, w.to(d), mask.to(d), ds_u, dsm_u, dw_u, db_u)
    assert torch.equal(ds_u, ds) and _one_ulp(dsm, dsm_u)
    assert rel_l2(dw, dw_u) <= 1e-5 and abs(l3u[0].item() - loss3[0].item()) <= 2e-6 * abs(l3u[0].item())


def test_fused_bwd_without_optional_outputs_and_null_ds():
    """dpot / vel / ds are optional: without them the kernel writes ds_masked, dw, db and the loss only"""
    from deepfluids_b200 import kernels as K
    pot, x, s, mask, w = _inputs((1, 8, 16, 16), 5)
    ds, dsm, dw, db, loss3, _, _ = _run_fused(pot, x, s, mask, w, want=True)
    d = dev()
    dsm2 = torch.empty_like(dsm)
    dw2, db2, l32 = torch.zeros_like(dw), torch.zeros_like(db), torch.zeros(3, device=d)
    K.lastconv_curl_loss_bwd(s.to(d), pot.to(d), x.to(d), w.to(d), mask.to(d), None, dsm2, dw2, db2, l32,
                             K.lastconv_curl_loss_workspace(d))
    assert torch.equal(dsm2, dsm) and rel_l2(dw2, dw) <= 1e-5 and torch.equal(l32, loss3)


def test_fused_bwd_full_size_128cube_vs_unfused_pair():
    """BASELINE configs[3] grid (one field of 128^3; 2 GB activations): every output of the fused kernel vs the un-fused
    kernel pair (whose stencil half is checked against the oracle at this size in test_gpu_baseline_sizes.py)"""
    from deepfluids_b200 import kernels as K
    d = dev()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, n = 1, 128
    pot = torch.randn(B, n, n, n, 3, device=d, generator=g) * 0.05
    pot = (pot + torch.roll(pot, 1, 1) + torch.roll(pot, 1, 2) + torch.roll(pot, 1, 3)) / 4
    x = torch.randn(B, n, n, n, 3, device=d, generator=g) * 0.05
    s = (torch.randn(B, n, n, n, 128, device=d, generator=g) * 0.5).bfloat16()
    mask = torch.randn(B, n, n, n, 128, device=d, generator=g).bfloat16()
    w = (torch.randn(3, 3, 3, 128, 3, device=d, generator=g) * 0.05).bfloat16().float()
    ds, dsm = torch.empty_like(s), torch.empty_like(s)
    dw, db, l3 = torch.zeros_like(w), torch.zeros(3, device=d), torch.zeros(3, device=d)
    dpot, vel = torch.empty_like(pot), torch.empty_like(pot)
    K.lastconv_curl_loss_bwd(s, pot, x, w, mask, ds, dsm, dw, db, l3, K.lastconv_curl_loss_workspace(d), dpot=dpot, vel=vel)
    l3u, dpu, velu = K.stencil_loss_fwdbwd(pot, x, want_vel=True)
    assert torch.equal(dpu, dpot) and torch.equal(velu, vel)
    assert abs(l3u[0].item() - l3[0].item()) <= 2e-6 * abs(l3u[0].item())
    ds_u, dsm_u = torch.empty_like(s), torch.empty_like(s)
    dw_u, db_u = torch.zeros_like(w), torch.zeros(3, device=d)
    K.lastconv_bwd(s, dpu, w, mask, ds_u, dsm_u, dw_u, db_u)
    assert torch.equal(ds_u, ds) and _one_ulp(dsm, dsm_u)
    assert rel_l2(dw, dw_u) <= 1e-4          # both reduce 2 M voxels with fp32 atomics in different orders


def test_trainer_gradients_and_loss_with_and_without_the_fusion():
    """the trainer's step body on the same weights and batch, fused vs un-fused: same loss, same flat gradient buffer up to
    bf16-ulp effects (ds_masked rounding) and the order of the fp32 atomics; then both train (loss falls)"""
    from deepfluids_b200 import config as C
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer3 import Trainer3
    args = ["--synthetic=true", "--is_3d=true", "--res_x=32", "--res_y=16", "--res_z=16", "--batch_size=2", "--num_conv=2",
            "--max_step=20", "--lr_max=0.001"]
    grads, losses, final = [], [], []
    for fused in ("1", "0"):
        os.environ["DFL_FUSED_LOSS"] = fused
        os.environ["DFL_CUDA_GRAPH"] = "0"
        try:
            cfg, _ = C.get_config(args)
            bm = BatchManager(cfg, pool=1)
            tr = Trainer3(cfg, bm)
            x, y = bm.batch()
            assert (tr._fused_args(x) is not None) == (fused == "1")
            tr._step_body_a(x, y)
            torch.cuda.synchronize()
            grads.append(tr.engine.params.grad.clone())
            losses.append(tr._loss3.tolist())
            first = losses[-1][0]
            for i in range(10):
                tr.train_step(x, y)
            final.append(tr.losses()[0])
            assert final[-1] < first
        finally:
            os.environ.pop("DFL_FUSED_LOSS", None)
            os.environ.pop("DFL_CUDA_GRAPH", None)
    assert abs(losses[0][0] - losses[1][0]) <= 2e-6 * abs(losses[1][0]), losses
    assert rel_l2(grads[0], grads[1]) <= 5e-3, rel_l2(grads[0], grads[1])
    # (after ten sign-like early Adam steps at lr 1e-3 the two runs are a few % apart in loss: both fell, asserted above)


# ------------------------------------------------------------------ 2D: the same entry point with ndim = 2
SHAPES_2D = [(2, 16, 32), (1, 8, 16), (3, 5, 7), (1, 2, 2), (2, 20, 38), (4, 128, 96), (2, 128, 128)]


def _inputs_2d(shape, seed):
    g = torch.Generator().manual_seed(seed)
    B, H, W = shape
    pot = torch.randn(B, H, W, 1, generator=g) * 0.05
    pot = (pot + torch.roll(pot, 1, 1) + torch.roll(pot, 1, 2)) / 3
    x, _ = T.synthetic_batch(B, [H, W], seed=seed + 1, smooth=1)
    s = (torch.randn(B, H, W, 128, generator=g) * 0.5).bfloat16()
    mask = torch.randn(B, H, W, 128, generator=g).bfloat16()
    w = R.xavier_uniform_((3, 3, 128, 1), g).bfloat16().float()
    return pot, x, s, mask, w


@pytest.mark.parametrize("shape", SHAPES_2D)
def test_fused_bwd_2d_vs_oracle_and_unfused_pair(shape):
    """2D (stream function psi, C = 1): the builder warps of the output conv's backward compute curl, both Jacobians, the
    loss terms and dL/dpsi on the tile's 15 x 23 footprint.  Oracle: loss (fp64), dL/dpsi (fp32 autograd), G_ bit for bit;
    un-fused pair (dfl_stencil_loss_fwdbwd + dfl_lastconv_bwd): every output bit for bit (same arithmetic, same kernel body)."""
    from deepfluids_b200 import kernels as K
    d = dev()
    pot, x, s, mask, w = _inputs_2d(shape, 23 + sum(shape))
    ds = torch.full(s.shape, float("nan"), dtype=torch.bfloat16, device=d)
    dsm = torch.full(s.shape, float("nan"), dtype=torch.bfloat16, device=d)
    dw, db, loss3 = torch.zeros(w.shape, device=d), torch.zeros(1, device=d), torch.zeros(3, device=d)
    dpot = torch.full(pot.shape, float("nan"), device=d)
    vel = torch.full(x.shape, float("nan"), device=d)
    ws = K.lastconv_curl_loss_workspace(d)
    for _ in range(2):                       # twice: the CTA ticket must reset itself
        dw.zero_(); db.zero_()
        K.lastconv_curl_loss_bwd(s.to(d), pot.to(d), x.to(d), w.to(d), mask.to(d), ds, dsm, dw, db, loss3, ws, 0.7, 1.3, 1.0,
                                 dpot=dpot, vel=vel)
    torch.cuda.synchronize()
    p = pot.clone().requires_grad_(True)
    loss, l1, jl1, vel_ref = T.stencil_loss(p, x, 0.7, 1.3)
    (gref,) = torch.autograd.grad(loss, p)
    with torch.no_grad():
        l64, l164, jl164, _ = T.stencil_loss(pot.double(), x.double(), 0.7, 1.3)
    assert torch.equal(vel.cpu(), vel_ref.detach())
    if shape[1] > 2:
        assert float(K.divergence(vel).abs().max()) <= 1e-5
    for got, want in zip(loss3.tolist(), (l64.item(), l164.item(), jl164.item())):
        assert abs(got - want) <= 3e-6 * abs(want), (loss3.tolist(), l64.item(), l164.item(), jl164.item())
    scale = float(gref.abs().max())
    assert float((dpot.cpu() - gref).abs().max()) <= 1e-6 * scale + 1e-12
    # output conv backward fed with the bf16 copy of that gradient (torch's CPU conv backward rejects the 2 x 2 image: that
    # shape is covered by the comparison with the un-fused pair below)
    if shape[1] * shape[2] >= 16:
        sin, wt, bt = s.float().requires_grad_(True), w.clone().requires_grad_(True), torch.zeros(1, requires_grad=True)
        y = R.conv_nd(sin, wt, bt, 1, None)
        gx, gw, gb = torch.autograd.grad(y, [sin, wt, bt], gref.bfloat16().float())
        assert rel_l2(ds.float(), gx) <= 4e-3
        assert rel_l2(dsm.float(), gx * torch.where(mask.float() >= 0, 1.0, 0.2)) <= 4e-3
        assert rel_l2(dw, gw) <= 2e-4
        assert float((db.cpu() - gb).abs().max()) <= 1e-3 * float(gref.abs().sum())
    # the un-fused pair
    l3u, dpu, velu = K.stencil_loss_fwdbwd(pot.to(d), x.to(d), 0.7, 1.3, want_vel=True)
    assert torch.equal(dpu, dpot) and torch.equal(velu, vel)
    ds_u, dsm_u = torch.empty_like(ds), torch.empty_like(dsm)
    dw_u, db_u = torch.zeros_like(dw), torch.zeros_like(db)
    K.lastconv_bwd(s.to(d), dpu, w.to(d), mask.to(d), ds_u, dsm_u, dw_u, db_u)
    assert torch.equal(ds_u, ds) and torch.equal(dsm_u, dsm)
    assert rel_l2(dw, dw_u) <= 1e-5 and abs(l3u[0].item() - loss3[0].item()) <= 2e-6 * abs(l3u[0].item())
    # optional outputs off: same ds_masked, dw and loss
    dsm2, dw2, db2, l32 = torch.empty_like(dsm), torch.zeros_like(dw), torch.zeros_like(db), torch.zeros(3, device=d)
    K.lastconv_curl_loss_bwd(s.to(d), pot.to(d), x.to(d), w.to(d), mask.to(d), None, dsm2, dw2, db2, l32, ws, 0.7, 1.3)
    assert torch.equal(dsm2, dsm) and rel_l2(dw2, dw) <= 1e-5 and torch.equal(l32, loss3)


def test_trainer_2d_step_fused_equals_unfused(monkeypatch):
    """one 2D train-step body with and without the fusion: same loss terms, same gradients (the weight-gradient atomics
    reorder sums: 1e-5)"""
    import importlib
    C = importlib.import_module("deepfluids_b200.config")
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    grads, losses = [], []
    monkeypatch.setenv("DFL_CUDA_GRAPH", "0")
    for flag in ("1", "0"):
        monkeypatch.setenv("DFL_FUSED_LOSS", flag)
        cfg, _ = C.get_config(["--synthetic=true", "--is_3d=false", "--res_x=32", "--res_y=48", "--batch_size=4", "--num_conv=2",
                               "--max_step=10"])
        bm = BatchManager(cfg, pool=1)
        tr = Trainer(cfg, bm)
        x, y = bm.batch()
        assert (tr._fused_args(x) is not None) == (flag == "1")
        tr._step_body_a(x, y)
        torch.cuda.synchronize()
        grads.append(tr.engine.params.grad.clone())
        losses.append(tr._loss3.clone())
    assert torch.allclose(losses[0], losses[1], rtol=2e-6, atol=0)
    assert float((grads[0] - grads[1]).abs().max()) <= 1e-5 * float(grads[1].abs().max())
