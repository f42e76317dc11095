"""GPU parity tests of the individual sm_100a kernels against the CPU oracle (oracle/), called through the C-ABI.

Tolerances (stated per test): fp32 stencil kernels reproduce the reference's own rounding order -> curl/jacobian
bit-exact, fused-loss gradient within 1e-6 abs of autograd on the oracle; bf16 tensor-core convolutions are compared
with the fp32 oracle evaluated on the SAME bf16-rounded inputs/weights -> rel-L2 <= 2e-3 (fp32 accumulation, bf16
output rounding only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_ops as R
from oracle import ref_train as T


def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------ stencils
def test_golden_stencil2d(golden_dir):
    from deepfluids_b200 import kernels as K
    g = np.load(golden_dir + "/stencil2d.npz")
    psi = torch.from_numpy(g["psi"]).to(dev())
    vel = torch.from_numpy(g["vel"]).to(dev())
    assert np.array_equal(K.curl_fwd(psi).cpu().numpy(), g["curl"])          # bit-exact
    jac, vort = K.jacobian_fwd(vel)
    assert np.array_equal(jac.cpu().numpy(), g["jac"]) and np.array_equal(vort.cpu().numpy(), g["vort"])
    div = K.divergence(torch.from_numpy(g["curl"]).to(dev()))
    assert np.array_equal(div.cpu().numpy(), g["div_of_curl"])
    assert float(div.abs().max()) <= 1e-5
    x = torch.from_numpy(g["x"]).to(dev())
    loss3, dpsi, v = K.stencil_loss_fwdbwd(psi, x, want_vel=True)
    assert np.array_equal(v.cpu().numpy(), g["curl"])
    l = loss3.cpu().numpy()
    np.testing.assert_allclose(l, [g["loss"], g["loss_l1"], g["loss_j_l1"]], rtol=2e-6)
    np.testing.assert_allclose(dpsi.cpu().numpy(), g["dpsi"], atol=1e-7, rtol=1e-5)


def test_golden_stencil3d(golden_dir):
    from deepfluids_b200 import kernels as K
    g = np.load(golden_dir + "/stencil3d.npz")
    A = torch.from_numpy(g["A"]).to(dev())
    vel = torch.from_numpy(g["vel"]).to(dev())
    assert np.array_equal(K.curl_fwd(A).cpu().numpy(), g["curl_of_A"])
    jac, c = K.jacobian_fwd(vel)
    assert np.array_equal(jac.cpu().numpy(), g["jac"]) and np.array_equal(c.cpu().numpy(), g["curl_of_vel"])
    div = K.divergence(torch.from_numpy(g["curl_of_A"]).to(dev()))
    assert np.array_equal(div.cpu().numpy(), g["div_of_curl"])
    x = torch.from_numpy(g["x"]).to(dev())
    loss3, dA, v = K.stencil_loss_fwdbwd(A, x, want_vel=True)
    assert np.array_equal(v.cpu().numpy(), g["curl_of_A"])
    np.testing.assert_allclose(loss3.cpu().numpy(), [g["loss"], g["loss_l1"], g["loss_j_l1"]], rtol=2e-6)
    np.testing.assert_allclose(dA.cpu().numpy(), g["dA"], atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("shape", [(3, 2, 2), (2, 40, 33), (1, 128, 96), (2, 5, 70)])
def test_stencil2d_vs_oracle(shape):
    from deepfluids_b200 import kernels as K
    B, H, W = shape
    g = torch.Generator().manual_seed(1)
    psi = torch.randn(B, H, W, 1, generator=g)
    x = torch.randn(B, H, W, 2, generator=g)
    p = psi.clone().requires_grad_(True)
    loss, l1, jl1, G = T.stencil_loss(p, x, 0.7, 1.3)
    (dp,) = torch.autograd.grad(loss, p)
    loss3, dpsi, v = K.stencil_loss_fwdbwd(psi.to(dev()), x.to(dev()), 0.7, 1.3, want_vel=True)
    assert torch.equal(v.cpu(), G.detach())
    np.testing.assert_allclose(loss3.cpu().numpy(), [loss.item(), l1.item(), jl1.item()], rtol=3e-6)
    np.testing.assert_allclose(dpsi.cpu().numpy(), dp.numpy(), atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("shape", [(2, 2, 2, 2), (1, 7, 13, 30), (2, 16, 16, 16), (1, 40, 12, 29), (1, 3, 64, 64),
                                   (1, 9, 37, 70), (2, 20, 40, 136), (1, 33, 130, 128), (3, 5, 3, 4), (1, 2, 11, 2)])
def test_stencil3d_vs_oracle(shape):
    from deepfluids_b200 import kernels as K
    B, D, H, W = shape
    g = torch.Generator().manual_seed(2)
    A = torch.randn(B, D, H, W, 3, generator=g)
    x = torch.randn(B, D, H, W, 3, generator=g)
    a = A.clone().requires_grad_(True)
    loss, l1, jl1, G = T.stencil_loss(a, x, 1.0, 0.5)
    (dA_ref,) = torch.autograd.grad(loss, a)
    loss3, dA, v = K.stencil_loss_fwdbwd(A.to(dev()), x.to(dev()), 1.0, 0.5, want_vel=True)
    assert torch.equal(v.cpu(), G.detach())
    np.testing.assert_allclose(loss3.cpu().numpy(), [loss.item(), l1.item(), jl1.item()], rtol=3e-6)
    np.testing.assert_allclose(dA.cpu().numpy(), dA_ref.numpy(), atol=1e-7, rtol=1e-5)


@pytest.mark.parametrize("shape", [(3, 64, 64, 64), (1, 128, 128, 128), (2, 21, 50, 66)])
def test_stencil3d_lean_matches_generic_kernel(shape, monkeypatch):
    """The all-fp32 3D fast path (dfl_stencil3_lean.cu: persistent CTAs, 2 voxels per thread) against the generic
    z-march kernel at BASELINE sizes: identical G_ (bit-exact), loss within 1e-6, dL/dA within fp32 rounding."""
    from deepfluids_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(*shape, 3, device=dev(), generator=g)
    x = torch.randn(*shape, 3, device=dev(), generator=g)
    l_fast, d_fast, v_fast = K.stencil_loss_fwdbwd(A, x, 0.8, 1.1, want_vel=True)
    l_fast, d_fast, v_fast = l_fast.clone(), d_fast.clone(), v_fast.clone()
    monkeypatch.setenv("DFL_STENCIL_GENERIC", "1")
    l_gen, d_gen, v_gen = K.stencil_loss_fwdbwd(A, x, 0.8, 1.1, want_vel=True)
    torch.cuda.synchronize()
    monkeypatch.delenv("DFL_STENCIL_GENERIC")
    assert torch.equal(v_fast, v_gen)
    np.testing.assert_allclose(l_fast.cpu().numpy(), l_gen.cpu().numpy(), rtol=1e-6)
    scale = float(d_gen.abs().max())
    assert float((d_fast - d_gen).abs().max()) <= 1e-5 * scale
    # every voxel written exactly once: no NaN / untouched entries
    d2 = torch.full_like(d_fast, float("nan"))
    K.stencil_loss_fwdbwd(A, x, 0.8, 1.1, dpot=d2)
    assert bool(torch.isfinite(d2).all()) and torch.equal(d2, d_fast)


def test_curl_divergence_free_full_size():
    """north_star acceptance: div(curl(A)) <= 1e-5 at BASELINE sizes (size-independent property)."""
    from deepfluids_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(2, 128, 128, 128, 3, device=dev(), generator=g)
    assert float(K.divergence(K.curl_fwd(A)).abs().max()) <= 1e-5
    psi = torch.randn(64, 128, 96, 1, device=dev(), generator=g)
    assert float(K.divergence(K.curl_fwd(psi)).abs().max()) <= 1e-5


def test_stencil_loss_directional_derivative_full_size():
    """Full-size (64^3) size-independent property: along the direction sign(dL/dA) the finite-difference slope of
    the fused kernel's loss equals sum|dL/dA| (L1 kinks flip ~eps of the signs -> 2% tolerance)."""
    from deepfluids_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(4)
    A = torch.randn(1, 64, 64, 64, 3, device=dev(), generator=g)
    x = torch.randn(1, 64, 64, 64, 3, device=dev(), generator=g)
    _, dA, _ = K.stencil_loss_fwdbwd(A, x)
    dirn = torch.sign(dA)
    eps = 1e-3
    lp = K.stencil_loss_fwdbwd(A + eps * dirn, x)[0][0].item()
    lm = K.stencil_loss_fwdbwd(A - eps * dirn, x)[0][0].item()
    fd = (lp - lm) / (2 * eps)
    an = float(dA.double().abs().sum())
    assert an > 0.1 and abs(fd - an) <= 2e-2 * an


# ------------------------------------------------------------------------------------------ convolution
def _conv_case(shape, seed, nd):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(*shape, 128, generator=g) * 0.5).bfloat16()
    w = R.xavier_uniform_((3,) * nd + (128, 128), g).bfloat16()
    b = (torch.randn(128, generator=g) * 0.1)
    return x, w, b


@pytest.mark.parametrize("shape", [(1, 8, 8, 8), (2, 4, 6, 10), (1, 16, 16, 16), (1, 3, 5, 37)])
def test_conv3d_fwd_lrelu(shape):
    from deepfluids_b200 import kernels as K
    x, w, b = _conv_case(shape, 10, 3)
    ref = R.conv_nd(x.float(), w.float(), b, 1, R.lrelu)
    wf, _ = K.pack_conv_weights(w.float().to(dev()))
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=dev())
    K.conv3x3(x.to(dev()), wf, b.to(dev()), out=out, flags=K.CONV_LRELU)
    assert rel_l2(out.float(), ref) <= 4e-3   # bf16 output rounding: 2^-9 relative per element


@pytest.mark.parametrize("shape", [(2, 8, 6), (1, 16, 12), (3, 33, 20), (1, 128, 96)])
def test_conv2d_fwd_lrelu(shape):
    from deepfluids_b200 import kernels as K
    x, w, b = _conv_case(shape, 11, 2)
    ref = R.conv_nd(x.float(), w.float(), b, 1, R.lrelu)
    wf, _ = K.pack_conv_weights(w.float().to(dev()))
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=dev())
    K.conv3x3(x.to(dev()), wf, b.to(dev()), out=out, flags=K.CONV_LRELU)
    assert rel_l2(out.float(), ref) <= 4e-3


def test_conv3d_residual_upsample_epilogue():
    from deepfluids_b200 import kernels as K
    shape = (2, 4, 8, 8)
    x, w, b = _conv_case(shape, 12, 3)
    g = torch.Generator().manual_seed(5)
    x0 = (torch.randn(*shape, 128, generator=g)).bfloat16()
    y = R.conv_nd(x.float(), w.float(), b, 1, R.lrelu)
    ref2 = R.upscale3(y + x0.float(), 2)
    wf, _ = K.pack_conv_weights(w.float().to(dev()))
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=dev())
    out2 = torch.empty(ref2.shape, dtype=torch.bfloat16, device=dev())
    K.conv3x3(x.to(dev()), wf, b.to(dev()), out=out, out2=out2, residual=x0.to(dev()),
              flags=K.CONV_LRELU | K.CONV_OUT2_UPSAMPLE)
    assert rel_l2(out.float(), y) <= 4e-3
    assert rel_l2(out2.float(), ref2) <= 4e-3


def test_conv2d_residual_upsample_epilogue():
    from deepfluids_b200 import kernels as K
    shape = (2, 8, 6)
    x, w, b = _conv_case(shape, 13, 2)
    g = torch.Generator().manual_seed(6)
    x0 = (torch.randn(*shape, 128, generator=g)).bfloat16()
    y = R.conv_nd(x.float(), w.float(), b, 1, R.lrelu)
    ref2 = R.upscale(y + x0.float(), 2)
    wf, _ = K.pack_conv_weights(w.float().to(dev()))
    out2 = torch.empty(ref2.shape, dtype=torch.bfloat16, device=dev())
    K.conv3x3(x.to(dev()), wf, b.to(dev()), out2=out2, residual=x0.to(dev()),
              flags=K.CONV_LRELU | K.CONV_OUT2_UPSAMPLE)
    assert rel_l2(out2.float(), ref2) <= 4e-3


@pytest.mark.parametrize("shape,nd", [((1, 8, 8, 8), 3), ((2, 5, 6, 9), 3), ((2, 16, 12), 2)])
def test_conv_dgrad_and_wgrad_vs_autograd(shape, nd):
    """dgrad (with lrelu-derivative mask + residual add epilogues) and wgrad/bias-grad vs torch autograd on the oracle."""
    from deepfluids_b200 import kernels as K
    x, w, b = _conv_case(shape, 20 + nd, nd)
    g = torch.Generator().manual_seed(7)
    dy = (torch.randn(*shape, 128, generator=g) * 0.1).bfloat16()
    xin = x.float().requires_grad_(True)
    wt = w.float().requires_grad_(True)
    bt = b.clone().requires_grad_(True)
    y = R.conv_nd(xin, wt, bt, 1, None)
    gx, gw, gb = torch.autograd.grad(y, [xin, wt, bt], dy.float())
    _, wd = K.pack_conv_weights(w.float().to(dev()))
    # plain dgrad
    dx = torch.empty(x.shape, dtype=torch.bfloat16, device=dev())
    K.conv3x3(dy.to(dev()), wd, None, out=dx)
    assert rel_l2(dx.float(), gx) <= 4e-3
    # dgrad * lrelu'(mask) and dgrad + residual
    mask_src = torch.randn(*shape, 128, generator=g).bfloat16()
    res = torch.randn(*shape, 128, generator=g).bfloat16()
    dxm = torch.empty_like(dx)
    dxr = torch.empty_like(dx)
    K.conv3x3(dy.to(dev()), wd, None, out=dxm, mask_src=mask_src.to(dev()))
    K.conv3x3(dy.to(dev()), wd, None, out2=dxr, residual=res.to(dev()))
    slope = torch.where(mask_src.float() >= 0, 1.0, 0.2)
    assert rel_l2(dxm.float(), gx * slope) <= 4e-3
    assert rel_l2(dxr.float(), gx + res.float()) <= 4e-3
    # wgrad (fp32 accumulate in TMEM, fp32 atomics) + bias grad
    dw = torch.zeros(w.shape, dtype=torch.float32, device=dev())
    db = torch.zeros(128, dtype=torch.float32, device=dev())
    db2 = torch.zeros(128, dtype=torch.float32, device=dev())
    K.conv3x3_wgrad(x.to(dev()), dy.to(dev()), dw, db2)       # bias gradient fused into a free TMEM slot
    K.bias_grad(dy.to(dev()), db)                             # standalone column-sum kernel
    assert rel_l2(dw, gw) <= 1e-4
    assert rel_l2(db, gb) <= 1e-4
    assert rel_l2(db2, gb) <= 1e-4


def test_conv_linearity_full_size():
    """Full-size (128^3, B=1) size-independent property: conv(a*x1 + x2) == a*conv(x1) + conv(x2) without bias/act
    (checked on a sub-sampled set of voxels), and agreement with the oracle on one boundary slab."""
    from deepfluids_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(8)
    shape = (1, 128, 128, 128, 128)
    x1 = (torch.randn(shape, device=dev(), generator=g) * 0.5).bfloat16()
    w = R.xavier_uniform_((3, 3, 3, 128, 128), torch.Generator().manual_seed(9)).bfloat16()
    wf, _ = K.pack_conv_weights(w.float().to(dev()))
    y1 = torch.empty(shape, dtype=torch.bfloat16, device=dev())
    K.conv3x3(x1, wf, None, out=y1)
    y2 = torch.empty_like(y1)
    K.conv3x3((x1.float() * 2).bfloat16(), wf, None, out=y2)     # exact scaling by 2 in bf16
    assert torch.equal((y1.float() * 2).bfloat16(), y2)
    # boundary slab z in [0,3): oracle on x[:, :4]
    ref = R.conv_nd(x1[:, :4].float().cpu(), w.float(), torch.zeros(128), 1, None)[:, :3]
    assert rel_l2(y1[:, :3].float(), ref) <= 4e-3


# ------------------------------------------------------------------------------------------ edge layers
@pytest.mark.parametrize("B,K_,N", [(8, 3, 6144), (4, 16, 65536)])
def test_fc_fwd_bwd(B, K_, N):
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(30)
    z = torch.rand(B, K_, generator=g) * 2 - 1
    W = R.xavier_uniform_((K_, N), g)
    b = torch.randn(N, generator=g) * 0.1
    ref = R.linear(z, W, b)
    out = K.fc_fwd(z.to(dev()), W.to(dev()), b.to(dev()))
    assert rel_l2(out.float(), ref) <= 4e-3
    out32 = K.fc_fwd(z.to(dev()), W.to(dev()), b.to(dev()), out_dtype=torch.float32)
    assert rel_l2(out32, ref) <= 1e-6
    dout = (torch.randn(B, N, generator=g)).bfloat16()
    dW = torch.empty(K_, N, device=dev())
    db = torch.empty(N, device=dev())
    K.fc_bwd(z.to(dev()), dout.to(dev()), dW, db)
    assert rel_l2(dW, z.t() @ dout.float()) <= 1e-5
    assert rel_l2(db, dout.float().sum(0)) <= 1e-5


@pytest.mark.parametrize("shape,nd,cout", [((2, 4, 6, 11), 3, 3), ((1, 8, 8, 8), 3, 3), ((1, 5, 20, 37), 3, 1), ((3, 16, 12), 2, 1),
                                           ((2, 9, 21), 2, 2), ((2, 40, 33), 2, 3), ((2, 40, 20, 37), 3, 3), ((1, 2, 2, 2), 3, 2),
                                           ((1, 33, 9, 16), 3, 1)])
def test_lastconv_tensorcore_fwd_and_fused_bwd(shape, nd, cout):
    """plane-GEMM + shift-sum forward (and the tap-window N=16 forward the fp32-grade path still uses) and the fused
    im2col-GEMM backward vs autograd on the oracle (bf16 operands, fp32 acc).  (2,40,20,37): 720 (column, plane) units
    over 148 CTAs = z-segments that start and end inside a column."""
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(41)
    x = (torch.randn(*shape, 128, generator=g) * 0.5).bfloat16()
    w = R.xavier_uniform_((3,) * nd + (128, cout), g).bfloat16().float()     # bf16-representable weights
    b = torch.randn(cout, generator=g) * 0.1
    xin = x.float().requires_grad_(True)
    wt = w.clone().requires_grad_(True)
    bt = b.clone().requires_grad_(True)
    y = R.conv_nd(xin, wt, bt, 1, None)
    dy = (torch.randn(y.shape, generator=g) * 1e-3).bfloat16().float()       # bf16-representable gradient
    gx, gw, gb = torch.autograd.grad(y, [xin, wt, bt], dy)
    w16 = K.pack_lastconv_weights(w.to(dev()))
    out = K.lastconv_fwd_tc(x.to(dev()), w16, b.to(dev()), cout)
    assert rel_l2(out, y.detach()) <= 1e-5
    out2 = torch.full(y.shape, float("nan"), device=dev())                   # every output voxel must be written
    K.lastconv_fwd(x.to(dev()), w.to(dev()), b.to(dev()), out=out2)
    assert rel_l2(out2, y.detach()) <= 1e-5
    out3 = K.lastconv_fwd(x.to(dev()), w.to(dev()), None)                   # bias is optional
    assert rel_l2(out3, y.detach() - b) <= 1e-5
    mask_src = torch.randn(*shape, 128, generator=g).bfloat16()
    ds = torch.empty(x.shape, dtype=torch.bfloat16, device=dev())
    dsm = torch.empty_like(ds)
    dw = torch.zeros(w.shape, device=dev())
    db = torch.zeros(cout, device=dev())
    K.lastconv_bwd(x.to(dev()), dy.to(dev()), w.to(dev()), mask_src.to(dev()), ds, dsm, dw, db)
    assert rel_l2(ds.float(), gx) <= 4e-3
    assert rel_l2(dsm.float(), gx * torch.where(mask_src.float() >= 0, 1.0, 0.2)) <= 4e-3
    assert rel_l2(dw, gw) <= 1e-4
    assert rel_l2(db, gb) <= 1e-5


@pytest.mark.parametrize("shape,nd,cout", [((1, 128, 128, 128), 3, 3), ((64, 128, 96), 2, 1)])
def test_lastconv_fwd_full_size_two_formulations_agree(shape, nd, cout):
    """BASELINE-size grids: the z-marching plane-GEMM kernel and the tap-window N=16 kernel (different tilings, different
    summation orders, both fp32 accumulate of the same bf16 products) must agree to fp32 rounding."""
    from deepfluids_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(43)
    x = (torch.randn(*shape, 128, device=dev(), generator=g) * 0.5).bfloat16()
    w = (torch.randn((3,) * nd + (128, cout), device=dev(), generator=g) * 0.05).bfloat16().float()
    b = torch.randn(cout, device=dev(), generator=g) * 0.1
    ref = K.lastconv_fwd_tc(x, K.pack_lastconv_weights(w), b, cout)
    out = torch.full_like(ref, float("nan"))
    K.lastconv_fwd(x, w, b, out=out)
    assert rel_l2(out, ref) <= 1e-5
    assert float((out - ref).abs().max()) <= 1e-4


@pytest.mark.parametrize("cshape,nd", [((2, 3, 4, 5), 3), ((2, 6, 7), 2)])
def test_pool_mask(cshape, nd):
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(50)
    fine = (cshape[0],) + tuple(2 * s for s in cshape[1:])
    gfine = torch.randn(*fine, 128, generator=g).bfloat16()
    mask_src = torch.randn(*cshape, 128, generator=g).bfloat16()
    # adjoint of nearest upsampling = sum over children (autograd through the oracle's upscale)
    u = torch.zeros(*cshape, 128, requires_grad=True)
    up = R.upscale3(u, 2) if nd == 3 else R.upscale(u, 2)
    (ref,) = torch.autograd.grad(up, u, gfine.float())
    ds = torch.empty(*cshape, 128, dtype=torch.bfloat16, device=dev())
    dm = torch.empty_like(ds)
    K.pool_mask(gfine.to(dev()), mask_src.to(dev()), ds, dm)
    assert rel_l2(ds.float(), ref) <= 4e-3
    assert rel_l2(dm.float(), ref * torch.where(mask_src.float() >= 0, 1.0, 0.2)) <= 4e-3


def test_adam_matches_tf_semantics():
    from deepfluids_b200 import kernels as K
    import math
    g = torch.Generator().manual_seed(60)
    p0 = torch.randn(10007, generator=g)
    var = {"w": p0.clone()}
    opt = T.TFAdam(var, 0.5, 0.999, 1e-8)
    p = p0.to(dev())
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    lr = 1e-3
    for t in range(1, 4):
        gr = torch.randn(10007, generator=g)
        opt.step(var, {"w": gr}, lr)
        lr_t = lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.5 ** t)
        K.adam_step(p, gr.to(dev()), m, v, lr_t, 0.5, 0.999, 1e-8)
    np.testing.assert_allclose(p.cpu().numpy(), var["w"].numpy(), rtol=1e-5, atol=1e-7)


def test_cabi_error_paths():
    """argument errors come back as negative codes with a message (no CPU fallback, no crash, no silent success)"""
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.cabi import DflError
    x = torch.zeros(1, 8, 8, 100, dtype=torch.bfloat16, device=dev())       # Cin = 100: unsupported
    w = torch.zeros(128, 900, dtype=torch.bfloat16, device=dev())
    out = torch.zeros(1, 8, 8, 128, dtype=torch.bfloat16, device=dev())
    with pytest.raises(DflError, match="Cin must be"):
        K.conv3x3(x, w, None, out=out)
    with pytest.raises(DflError, match="no output buffer"):
        K.conv3x3(torch.zeros(1, 8, 8, 128, dtype=torch.bfloat16, device=dev()), torch.zeros(128, 9 * 128, dtype=torch.bfloat16, device=dev()))
    with pytest.raises(DflError, match="extent must be >= 2"):
        K.curl_fwd(torch.zeros(1, 1, 8, 1, device=dev()))
    with pytest.raises(DflError, match="K <= 16"):
        K.fc_fwd(torch.zeros(2, 17, device=dev()), torch.zeros(17, 64, device=dev()), torch.zeros(64, device=dev()))
    with pytest.raises(TypeError):
        K.curl_fwd(torch.zeros(1, 8, 8, 1, dtype=torch.float16, device=dev()))
    with pytest.raises(AssertionError):
        K.curl_fwd(torch.zeros(1, 8, 8, 1))                                  # CPU tensor: refused, never computed on the host
