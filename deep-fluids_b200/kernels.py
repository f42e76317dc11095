"""Thin torch-tensor front end of the C-ABI (one function per entry point of include/deepfluids_b200.h).

torch is plumbing here: it owns the device memory and the stream; every call passes raw pointers through
ctypes to hand-written sm_100a kernels.  Nothing in this module computes on the CPU or through torch ops.
"""
import contextlib
import ctypes as C
import os

import torch

from . import cabi
from .cabi import (F32, BF16, CONV_LRELU, CONV_OUT2_UPSAMPLE, CONV_MASK_AFTER_RESIDUAL, CONV_SPLIT_IO, check,
                   dims_array)

_DT = {torch.float32: F32, torch.bfloat16: BF16}


class _Prof(object):
    """Launch counter (bench.py's `gpu_launches` claim) and optional per-kernel CUDA-event timing."""
    launches = 0
    # when a list: (kernel name, start event, end event, ALGORITHMIC work, EXECUTED work) tuples are appended.  The two differ
    # only for the phase-decomposed upsample-conv launches, which are booked with the dense layer's algorithmic FLOPs
    events = None

    # DFL_NVTX=1: every launch (and the trainers' step phases, see `nvtx_range`) is wrapped in an NVTX range, so that
    # `ncu --nvtx --nvtx-include "conv3x3_fwd/"` / an Nsight Systems timeline can address them by name.  Off by default: the
    # push / pop pair costs ~1 us per launch on the host, and captured graphs replay without host-side ranges anyway.
    nvtx = os.environ.get("DFL_NVTX", "0") == "1"

    @classmethod
    def timed(cls, name, work, fn, exec_work=None):
        cls.launches += 1
        if cls.nvtx:
            torch.cuda.nvtx.range_push(name)
            try:
                return cls._timed(name, work, fn, exec_work)
            finally:
                torch.cuda.nvtx.range_pop()
        return cls._timed(name, work, fn, exec_work)

    @classmethod
    def _timed(cls, name, work, fn, exec_work=None):
        if cls.events is None:
            return fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        cls.events.append((name, s, e, work, work if exec_work is None else exec_work))


PROF = _Prof


@contextlib.contextmanager
def nvtx_range(name):
    """NVTX range around a phase of the step (forward / loss+first backward / backward / exchange / optimizer) when DFL_NVTX=1"""
    if not PROF.nvtx:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "deepfluids_b200 kernels need contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("unsupported dtype %s (float32 / bfloat16 only)" % t.dtype)


def _spatial(t):
    """channels-last tensor [B,(D,)H,W,C] -> (dims array, ndim)"""
    nd = t.dim() - 2
    assert nd in (2, 3), "expected NHWC or NDHWC tensor"
    return dims_array(t.shape[:-1]), nd


_DET = {"ws": None}


def set_deterministic(on, device=None):
    """dfl_set_deterministic: weight / bias gradients of the 128 -> 128 layers reduce their split-K partial sums in a fixed
    order (through a workspace this module owns) instead of fp32 atomics.  Process-wide; launches that use the workspace must
    stay on one stream."""
    lib = cabi.lib()
    if not on:
        check(lib.dfl_set_deterministic(None, 0))
        _DET["ws"] = None
        return
    n = int(lib.dfl_deterministic_workspace_bytes())
    ws = torch.empty(n, dtype=torch.uint8, device=device or torch.device("cuda", torch.cuda.current_device()))
    check(lib.dfl_set_deterministic(_p(ws), n))
    _DET["ws"] = ws


def deterministic():
    return _DET["ws"] is not None


# ------------------------------------------------------------------ stencils
def curl_fwd(pot):
    d, nd = _spatial(pot)
    vel = torch.empty(pot.shape[:-1] + (nd,), dtype=pot.dtype, device=pot.device)
    PROF.timed("curl_fwd", 0.0, lambda: check(cabi.lib().dfl_curl_fwd(_p(pot), _p(vel), d, nd, pot.shape[-1], _dt(pot), _st())))
    return vel


def jacobian_fwd(vel, want_jac=True, want_aux=True):
    d, nd = _spatial(vel)
    assert vel.shape[-1] == nd
    jac = torch.empty(vel.shape[:-1] + (nd * nd,), dtype=vel.dtype, device=vel.device) if want_jac else None
    aux = torch.empty(vel.shape[:-1] + (1 if nd == 2 else 3,), dtype=vel.dtype, device=vel.device) if want_aux else None
    PROF.timed("jacobian_fwd", 0.0, lambda: check(cabi.lib().dfl_jacobian_fwd(_p(vel), _p(jac), _p(aux), d, nd, _dt(vel), _st())))
    return jac, aux


def divergence(vel):
    d, nd = _spatial(vel)
    out = torch.empty((vel.shape[0],) + tuple(s - 1 for s in vel.shape[1:-1]) + (1,), dtype=vel.dtype, device=vel.device)
    PROF.timed("divergence", 0.0, lambda: check(cabi.lib().dfl_divergence(_p(vel), _p(out), d, nd, _dt(vel), _st())))
    return out


def curl_bwd(dvel, pot_channels=None):
    """dpot = curl^T(dvel) (fp32); 2D: `pot_channels` channels (gradient in channel 0), 3D: 3"""
    d, nd = _spatial(dvel)
    assert dvel.dtype == torch.float32 and dvel.shape[-1] == nd
    cs = 3 if nd == 3 else int(pot_channels or 1)
    dpot = torch.empty(dvel.shape[:-1] + (cs,), dtype=torch.float32, device=dvel.device)
    PROF.timed("curl_bwd", 0.0, lambda: check(cabi.lib().dfl_curl_bwd(_p(dvel), _p(dpot), d, nd, cs, _st())))
    return dpot


def jacobian_bwd(djac, daux):
    """dvel = J^T djac + aux^T daux (fp32; either may be None)"""
    ref = djac if djac is not None else daux
    d, nd = _spatial(ref)
    assert all(t is None or t.dtype == torch.float32 for t in (djac, daux))
    dvel = torch.empty(ref.shape[:-1] + (nd,), dtype=torch.float32, device=ref.device)
    PROF.timed("jacobian_bwd", 0.0, lambda: check(cabi.lib().dfl_jacobian_bwd(_p(djac), _p(daux), _p(dvel), d, nd, _st())))
    return dvel


def mse_loss(dvals, target, scale=1.0, want_grad=True):
    """LSGAN term: (mean((d - target)^2) as a 1-element fp32 tensor, scale * dloss/dd or None)"""
    assert dvals.dtype == torch.float32
    loss = torch.empty(1, dtype=torch.float32, device=dvals.device)
    dd = torch.empty_like(dvals) if want_grad else None
    PROF.timed("mse_loss", 0.0, lambda: check(cabi.lib().dfl_mse_loss(_p(dvals), float(target), _p(loss), _p(dd), dvals.numel(), float(scale), _st())))
    return loss, dd


_L1_WS = {}


def l1_loss(a, b, scale=1.0, dd=None, accumulate=False, want_grad=True):
    """(mean|a - b| as a 1-element tensor, dd (+)= scale * sgn(a - b) / n): the un-fused L1 term (use_curl=False)"""
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape
    ws = _L1_WS.get(a.device)
    if ws is None:
        ws = _L1_WS[a.device] = torch.zeros(cabi.lib().dfl_l1_loss_workspace_bytes(), dtype=torch.uint8, device=a.device)
    loss = torch.empty(1, dtype=torch.float32, device=a.device)
    if dd is None and want_grad:
        assert not accumulate
        dd = torch.empty_like(a)
    PROF.timed("l1_loss", 0.0, lambda: check(cabi.lib().dfl_l1_loss(_p(a), _p(b), _p(loss), _p(dd), a.numel(), float(scale),
                                                                  int(accumulate), _p(ws), _st())))
    return loss, dd


def velocity_loss_fwdbwd(vel, x, w1=1.0, w2=1.0, dvel=None, loss3=None):
    """use_curl=False (trainer.py:141-144,170-172): the network output IS the velocity G_.
    loss = w1*mean|G_ - x| + w2*mean|J(G_) - J(x)|;  -> (loss3 [total, l1, jl1], d loss / d G_).  Un-fused sequence of the
    standalone kernels: jacobian (x2), L1 (x2), jacobian adjoint."""
    vel, x = vel.contiguous(), x.contiguous()
    jv, _ = jacobian_fwd(vel, want_aux=False)
    jx, _ = jacobian_fwd(x, want_aux=False)
    jl1, gj = l1_loss(jv, jx, w2)
    g = jacobian_bwd(gj, None)
    if dvel is not None:
        dvel.copy_(g)
        g = dvel
    l1, _ = l1_loss(vel, x, w1, dd=g, accumulate=True)
    if loss3 is None:
        loss3 = torch.empty(3, dtype=torch.float32, device=vel.device)
    loss3[1:2].copy_(l1)
    loss3[2:3].copy_(jl1)
    loss3[0:1].copy_(w1 * l1 + w2 * jl1)
    return loss3, g


def stencil_loss_fwdbwd(pot, x, w1=1.0, w2=1.0, grad_scale=1.0, want_vel=False, dpot=None, loss3=None, workspace=None):
    """-> (loss3 [total,l1,jl1] float32 device tensor, dpot, vel|None).  If `dpot` is given with more than one channel in
    2D, the gradient is written full-shape (channel 0 = d/d psi, the rest 0)."""
    d, nd = _spatial(x)
    l = cabi.lib()
    nb = l.dfl_stencil_loss_workspace_bytes(d, nd)
    if workspace is None or workspace.numel() < nb:
        workspace = torch.empty(nb, dtype=torch.uint8, device=x.device)
    if dpot is None:
        dpot = torch.empty(x.shape[:-1] + (1 if nd == 2 else 3,), dtype=pot.dtype, device=x.device)
    if loss3 is None:
        loss3 = torch.empty(3, dtype=torch.float32, device=x.device)
    vel = torch.empty(x.shape, dtype=pot.dtype, device=x.device) if want_vel else None
    nvox = x.numel() // x.shape[-1]
    work = nvox * (pot.shape[-1] * pot.element_size() + x.shape[-1] * x.element_size()
                   + dpot.shape[-1] * dpot.element_size())          # algorithmic bytes: read pot + x, write dpot
    PROF.launches += 1                                              # + the tiny finalize kernel
    PROF.timed("stencil_fused", work, lambda: check(l.dfl_stencil_loss_fwdbwd_ex(
        _p(pot), _p(x), _p(dpot), _p(vel), _p(loss3), _p(workspace), d, nd, pot.shape[-1], dpot.shape[-1], w1, w2,
        grad_scale, _dt(pot), _dt(x), _st())))
    return loss3, dpot, vel


# ------------------------------------------------------------------ FC
def fc_fwd(z, W, bias, out_dtype=torch.bfloat16, out=None):
    B, K = z.shape
    N = W.shape[1]
    if out is None:
        out = torch.empty(B, N, dtype=out_dtype, device=z.device)
    PROF.timed("fc_fwd", 0.0, lambda: check(cabi.lib().dfl_fc_fwd(_p(z), _p(W), _p(bias), _p(out), B, K, N, _dt(out), _st())))
    return out


def fc_bwd(z, dout, dW, db):
    B, K = z.shape
    N = dW.shape[1]
    PROF.timed("fc_bwd", 0.0, lambda: check(cabi.lib().dfl_fc_bwd(_p(z), _p(dout), _p(dW), _p(db), B, K, N, _dt(dout), _st())))


def gemm(a, b, bias=None, out=None, trans_a=False, trans_b=False, accumulate=False):
    """fp32 C[M,N] = op(a) op(b) (+ bias) (+ C): the general fully connected layer (dfl_gemm_f32)"""
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2
    M, Kd = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    assert (b.shape[1] if trans_b else b.shape[0]) == Kd, "gemm: inner dimensions differ"
    if out is None:
        assert not accumulate
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    PROF.timed("gemm_f32", 0.0, lambda: check(cabi.lib().dfl_gemm_f32(_p(a), _p(b), _p(bias), _p(out), M, N, Kd, int(trans_a),
                                                                  int(trans_b), int(accumulate), _st())))
    return out


def colsum(x):
    out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    PROF.timed("colsum_f32", 0.0, lambda: check(cabi.lib().dfl_colsum_f32(_p(x), _p(out), x.shape[0], x.shape[1], _st())))
    return out


ACT_NONE, ACT_LRELU, ACT_ELU = 0, 1, 2


def bn_act_fwd(x, gamma, beta, moving_mean, moving_var, eps, decay, training, act):
    """slim.batch_norm + activation on [M, N] fp32 -> (y, save_mean, save_rstd); moving statistics updated in place"""
    M, N = x.shape
    y = torch.empty_like(x)
    sm = torch.empty(N, dtype=torch.float32, device=x.device) if training else None
    sr = torch.empty(N, dtype=torch.float32, device=x.device) if training else None
    PROF.timed("bn_act_fwd", 0.0, lambda: check(cabi.lib().dfl_bn_act_fwd(_p(x), _p(gamma), _p(beta), _p(moving_mean), _p(moving_var), _p(y), _p(sm), _p(sr), M, N,
                                        float(eps), float(decay), int(bool(training)), int(act), _st())))
    return y, sm, sr


def bn_act_bwd(x, y, dy, gamma, save_mean, save_rstd, act, want_dx=True):
    M, N = x.shape
    dx = torch.empty_like(x) if want_dx else None
    dgamma = torch.empty(N, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(N, dtype=torch.float32, device=x.device)
    PROF.timed("bn_act_bwd", 0.0, lambda: check(cabi.lib().dfl_bn_act_bwd(_p(x), _p(y), _p(dy), _p(gamma), _p(save_mean), _p(save_rstd), _p(dx), _p(dgamma), _p(dbeta), M, N,
                                        int(act), _st())))
    return dx, dgamma, dbeta


def dropout(x, keep_prob, seed, offset):
    y = torch.empty_like(x)
    PROF.timed("dropout", 0.0, lambda: check(cabi.lib().dfl_dropout(_p(x), _p(y), x.numel(), float(keep_prob), int(seed), int(offset), _st())))
    return y


# ------------------------------------------------------------------ conv
def pack_conv_weights(w, w_fwd=None, w_dgrad=None):
    """w fp32 TF layout [k,(k,)k,Cin,Cout] -> (bf16 [Cout, taps*Cin], bf16 [Cin, taps*Cout])"""
    cin, cout = w.shape[-2], w.shape[-1]
    taps = w.numel() // (cin * cout)
    if w_fwd is None:
        w_fwd = torch.empty(cout, taps * cin, dtype=torch.bfloat16, device=w.device)
    if w_dgrad is None:
        w_dgrad = torch.empty(cin, taps * cout, dtype=torch.bfloat16, device=w.device)
    PROF.timed("pack_conv_weights", 0.0, lambda: check(cabi.lib().dfl_pack_conv_weights(_p(w), _p(w_fwd), _p(w_dgrad), taps, cin, cout, _st())))
    return w_fwd, w_dgrad


def pack_phase_weights(w, w_fwd, w_dgrad):
    """fp32 TF-layout [3,(3,)3,Cin,Cout] -> operands of the phase-decomposed upsample-conv (dfl_pack_phase_weights)"""
    nd = w.dim() - 2
    PROF.timed("pack_conv_weights", 0.0, lambda: check(cabi.lib().dfl_pack_phase_weights(_p(w), _p(w_fwd), _p(w_dgrad), nd, w.shape[-2], w.shape[-1], _st())))


def gather_stride2(fine, coarse):
    """coarse = fine[:, ::2, ::2(, ::2), :]  (bf16 [.., 128])"""
    d, nd = _spatial(coarse)
    PROF.timed("gather_stride2", 0.0, lambda: check(cabi.lib().dfl_gather_stride2(_p(fine), _p(coarse), d, nd, _st())))


def phase_wgrad(dy_fine, s_coarse, t_scratch, dw, alg_flops=None):
    """weight gradient of a phase-decomposed upsample-conv: 4^nd-tap stride-2 tensor-core correlation + fold onto dw"""
    nd = dy_fine.dim() - 2
    t_scratch.zero_()
    flops = 2.0 * (s_coarse.numel() // 128) * 128 * 128 * (4 ** nd)
    PROF.timed("wgrad_tc", alg_flops if alg_flops is not None else flops, lambda: check(cabi.lib().dfl_phase_wgrad(
        _p(dy_fine), _p(s_coarse), _p(t_scratch), dims_array(dy_fine.shape[:-1]), dims_array(s_coarse.shape[:-1]), nd, _st())),
        exec_work=flops)
    PROF.timed("phase_wgrad_fold", 0.0, lambda: check(cabi.lib().dfl_phase_wgrad_fold(_p(t_scratch), _p(dw), nd, dw.shape[-2], dw.shape[-1], _st())))


def pack_conv_weights_multi(ptr_table, n_layers, taps, cin, cout):
    """all same-shape layers in one launch; ptr_table = int64 device tensor [3, n_layers] of {w, w_fwd, w_dgrad} addresses"""
    PROF.timed("pack_conv_weights", 0.0, lambda: check(cabi.lib().dfl_pack_conv_weights_multi(
        _p(ptr_table), n_layers, taps, cin, cout, _st())))


def conv3x3(x, w_packed, bias=None, out=None, out2=None, residual=None, mask_src=None, flags=0, nblk=1):
    """tcgen05 implicit-GEMM conv; see dfl_conv3x3_fwd in include/deepfluids_b200.h.  `nblk` > 1: x is a channel-blocked
    tensor [nblk*B,(D,)H,W,128] holding nblk*128 input channels."""
    nd = x.dim() - 2
    d = dims_array((x.shape[0] // nblk,) + tuple(x.shape[1:-1]))
    cin, cout = x.shape[-1] * nblk, w_packed.shape[0]
    assert x.dtype == torch.bfloat16 and w_packed.dtype == torch.bfloat16 and x.shape[0] % nblk == 0
    flops = 2.0 * (x.numel() // cin) * cin * cout * (3 ** nd)     # algorithmic MACs x 2 (dense taps)
    PROF.timed("conv_tc", flops, lambda: check(cabi.lib().dfl_conv3x3_fwd(
        _p(x), _p(w_packed), _p(bias), _p(out), _p(out2), _p(residual), _p(mask_src), d, nd, cin, cout, flags, _st())))


def pack_conv_weights_ex(w, w_fwd, w_dgrad, cin_ld):
    """as pack_conv_weights, forward operand laid out for `cin_ld` (>= Cin) input channels"""
    cin, cout = w.shape[-2], w.shape[-1]
    taps = w.numel() // (cin * cout)
    PROF.timed("pack_conv_weights_ex", 0.0, lambda: check(cabi.lib().dfl_pack_conv_weights_ex(_p(w), _p(w_fwd), _p(w_dgrad), taps, cin, cout, cin_ld, _st())))


def conv_taps(x, w_rows, bias, out, out2, residual, mask_src, tile_dims, out_dims, cin, in_stride, taps, out_stride,
              out_off, flags=0, alg_flops=None):
    """generic per-tap tensor-core conv (dfl_conv_taps).  x: channel-blocked input [nblk*B,(D,)H,W,128]; w_rows: bf16
    2-D view [128, w_ld] of the packed weight rows of this output block; taps: list of (dz, dy, dx, kcol)."""
    nd = x.dim() - 2
    tarr = (C.c_int32 * (4 * len(taps)))(*[int(v) for t in taps for v in t])
    oarr = (C.c_int32 * 3)(*([int(v) for v in out_off] + [0] * (3 - len(out_off))))
    assert w_rows.shape[0] == 128 and w_rows.stride(1) == 1
    flops = 2.0 * float(torch.tensor(tile_dims).prod()) * cin * 128 * len(taps)
    # alg_flops (phase-decomposed upsample-conv): the roofline numerator stays the dense layer's FLOPs
    PROF.timed("conv_tap", flops if alg_flops is None else alg_flops, lambda: check(cabi.lib().dfl_conv_taps(
        _p(x), C.c_void_p(w_rows.data_ptr()), _p(bias), _p(out), _p(out2), _p(residual), _p(mask_src),
        dims_array(x.shape[:-1]), dims_array(tile_dims), dims_array(out_dims), nd, cin, in_stride, len(taps), tarr,
        out_stride, oarr, w_rows.stride(0), flags, _st())), exec_work=flops)


def conv_wgrad_ex(x, dpre, dw_ptr_tensor, db, in_stride, pad, dw_tap_stride, dw_row_stride):
    """weight gradient of one (input block, output block) pair; dw_ptr_tensor: fp32 tensor whose data_ptr is the
    address of element (tap 0, ci 0, co 0) of the destination block."""
    nd = x.dim() - 2
    flops = 2.0 * (dpre.numel() // 128) * 128 * 128 * (3 ** nd)
    PROF.timed("wgrad_tc", flops, lambda: check(cabi.lib().dfl_conv_wgrad_ex(
        _p(x), _p(dpre), C.c_void_p(dw_ptr_tensor.data_ptr()), _p(db), dims_array(x.shape[:-1]),
        dims_array(dpre.shape[:-1]), nd, in_stride, pad, dw_tap_stride, dw_row_stride, _st())))


def pad_cast(x, out):
    """fp32 [..., cin] -> bf16 [..., 128] zero padded"""
    PROF.timed("pad_cast", 0.0, lambda: check(cabi.lib().dfl_pad_cast(_p(x), _p(out), x.numel() // x.shape[-1], x.shape[-1], _st())))


def add_mask(a, b, y, out):
    """out = (a + b) * lrelu'(y)   (b, y optional)"""
    PROF.timed("add_mask", 0.0, lambda: check(cabi.lib().dfl_add_mask(_p(a), _p(b), _p(y), _p(out), a.numel(), _st())))


def enc_fc_fwd(flat, W, bias, z, nblk):
    B = flat.shape[0] // nblk
    V = flat.numel() // (flat.shape[0] * 128)
    PROF.timed("enc_fc_fwd", 0.0, lambda: check(cabi.lib().dfl_enc_fc_fwd(_p(flat), _p(W), _p(bias), _p(z), B, V, nblk, W.shape[1], _st())))


def enc_fc_bwd(flat, W, dz, dW, db, dflat, nblk):
    B = flat.shape[0] // nblk
    V = flat.numel() // (flat.shape[0] * 128)
    PROF.timed("enc_fc_bwd", 0.0, lambda: check(cabi.lib().dfl_enc_fc_bwd(_p(flat), _p(W), _p(dz), _p(dW), _p(db), _p(dflat), B, V, nblk, W.shape[1], _st())))


def fc_dz(dout, W, dz, accumulate=False):
    B, N = dout.shape
    PROF.timed("fc_dz", 0.0, lambda: check(cabi.lib().dfl_fc_dz(_p(dout), _p(W), _p(dz), B, W.shape[0], N, _dt(dout), 1 if accumulate else 0, _st())))


def ae_loss_p(z, y_last, dz, loss_p, scale):
    B, Z = z.shape
    PROF.timed("ae_loss_p", 0.0, lambda: check(cabi.lib().dfl_ae_loss_p(_p(z), _p(y_last), _p(dz), _p(loss_p), B, Z, y_last.shape[1], scale, _st())))


def ae_sigmoid(z_lin, z):
    PROF.timed("ae_sigmoid", 0.0, lambda: check(cabi.lib().dfl_ae_sigmoid(_p(z_lin), _p(z), z_lin.numel(), _st())))


def ae_sparse_bwd(z, dz, dz_lin, loss_kl, p_num, rho, w5):
    B, Z = z.shape
    PROF.timed("ae_sparse_bwd", 0.0, lambda: check(cabi.lib().dfl_ae_sparse_bwd(_p(z), _p(dz), _p(dz_lin), _p(loss_kl), B, Z, p_num, rho, w5, _st())))


# ------------------------------------------------------------------ fp32-grade mode (bf16x3 split operands)
_SPLIT_MAP = (C.c_int32 * 3)(0, 1, 0)


def pack_conv_weights_split(w, w_fwd, w_dgrad):
    """fp32 TF-layout weights -> split tensor-core operands [rows][taps*384] (see dfl_pack_conv_weights_split)"""
    cin, cout = w.shape[-2], w.shape[-1]
    taps = w.numel() // (cin * cout)
    PROF.timed("pack_conv_weights_split", 0.0, lambda: check(cabi.lib().dfl_pack_conv_weights_split(_p(w), _p(w_fwd), _p(w_dgrad), taps, cin, cout, _st())))


def conv3x3_split(x2, w_split, bias=None, out=None, out2=None, residual=None, mask_src=None, flags=0, cout=128):
    """fp32-grade conv: x2 = (hi, lo) pair [2*B,(D,)H,W,128]; bf16 outputs / residual are pairs too; cout < 128 selects
    the output-conv variant (fp32 out)."""
    nd = x2.dim() - 2
    d = dims_array((x2.shape[0] // 2,) + tuple(x2.shape[1:-1]))
    flops = 2.0 * (x2.numel() // 256) * 384 * cout * (3 ** nd)
    fl = flags | (CONV_SPLIT_IO if cout == 128 else 0)
    PROF.timed("conv_tc", flops, lambda: check(cabi.lib().dfl_conv3x3_fwd_ex(
        _p(x2), _p(w_split), _p(bias), _p(out), _p(out2), _p(residual), _p(mask_src), d, nd, 384, cout, fl, _SPLIT_MAP, 2,
        _st())))


def split_f32(x, out, cpad=None):
    """fp32 [..., c] -> (hi, lo) pair written to out [2, ..., cpad]"""
    c = x.shape[-1]
    PROF.timed("split_f32", 0.0, lambda: check(cabi.lib().dfl_split_f32(_p(x), _p(out), x.numel() // c, c, cpad or c, _st())))


def merge_split(x2, out):
    PROF.timed("merge_split", 0.0, lambda: check(cabi.lib().dfl_merge_split(_p(x2), _p(out), out.numel(), _st())))


def pool_mask_split(g2, mask_src, ds2, dmasked2):
    ref = ds2 if ds2 is not None else dmasked2
    nd = ref.dim() - 2
    d = dims_array((ref.shape[0] // 2,) + tuple(ref.shape[1:-1]))
    PROF.timed("pool_mask_split", 0.0, lambda: check(cabi.lib().dfl_pool_mask_split(_p(g2), _p(mask_src), _p(ds2), _p(dmasked2), d, nd, _st())))


def conv3x3_wgrad(x, dpre, dw, db=None):
    """dw += x^T (x) dpre per tap; db += column sums of dpre (optional, free in the kernel)"""
    d, nd = _spatial(x)
    flops = 2.0 * (x.numel() // x.shape[-1]) * x.shape[-1] * dpre.shape[-1] * (3 ** nd)
    PROF.timed("wgrad_tc", flops, lambda: check(cabi.lib().dfl_conv3x3_wgrad(
        _p(x), _p(dpre), _p(dw), _p(db), d, nd, x.shape[-1], dpre.shape[-1], _st())))


def conv3x3_wgrad_split(x2, dpre2, dw, db=None):
    """fp32-grade dw += x^T (x) dP on (hi, lo) pairs [2B,...,128]: three operand combinations in ONE launch"""
    nd = x2.dim() - 2
    d = dims_array((x2.shape[0] // 2,) + tuple(x2.shape[1:-1]))
    flops = 2.0 * 3 * (x2.numel() // 256) * 128 * 128 * (3 ** nd)
    PROF.timed("wgrad_tc", flops, lambda: check(cabi.lib().dfl_conv3x3_wgrad_split(
        _p(x2), _p(dpre2), _p(dw), _p(db), d, nd, _st())))


def bias_grad(dpre, db):
    PROF.timed("bias_grad", 0.0, lambda: check(cabi.lib().dfl_bias_grad(_p(dpre), _p(db), dpre.numel() // dpre.shape[-1], _st())))


def pack_lastconv_weights(w, w16=None):
    """w fp32 TF layout [k,(k,)k,128,C] -> bf16 [16, taps*128] (rows >= C stay zero): operand of the small-Cout conv."""
    cin, cout = w.shape[-2], w.shape[-1]
    taps = w.numel() // (cin * cout)
    if w16 is None:
        w16 = torch.zeros(16, taps * cin, dtype=torch.bfloat16, device=w.device)
    PROF.timed("pack_conv_weights", 0.0, lambda: check(cabi.lib().dfl_pack_conv_weights(_p(w), _p(w16), None, taps, cin, cout, _st())))
    return w16


def lastconv_fwd_tc(x, w16, bias, cout, out=None):
    """128 -> cout (1..3) output conv on the tap-window tensor-core kernel (N = 16 variant); out fp32 [.., cout]."""
    d, nd = _spatial(x)
    if out is None:
        out = torch.empty(x.shape[:-1] + (cout,), dtype=torch.float32, device=x.device)
    flops = 2.0 * (x.numel() // 128) * 128 * cout * (3 ** nd)
    PROF.timed("lastconv_fwd_tc", flops, lambda: check(cabi.lib().dfl_conv3x3_fwd(
        _p(x), _p(w16), _p(bias), _p(out), None, None, None, d, nd, 128, cout, 0, _st())))
    return out


def lastconv_fwd(x, w, bias, out=None):
    """128 -> cout (1..3) output conv (model.py:42,84): one GEMM per input plane + shift-sum (z-marching tensor-core
    kernel); w = the fp32 TF-layout variable [3,(3,)3,128,cout]; out fp32 [.., cout]."""
    d, nd = _spatial(x)
    cout = w.shape[-1]
    if out is None:
        out = torch.empty(x.shape[:-1] + (cout,), dtype=torch.float32, device=x.device)
    flops = 2.0 * (x.numel() // 128) * 128 * cout * (3 ** nd)
    PROF.timed("lastconv_fwd_tc", flops, lambda: check(cabi.lib().dfl_lastconv_fwd(
        _p(x), _p(w), _p(bias), _p(out), d, nd, cout, _st())))
    return out


def lastconv_bwd(s, dout, w, mask_src, ds, ds_masked, dw, db):
    """fused dgrad + wgrad + bias-grad of the output conv (tensor cores)"""
    d, nd = _spatial(s)
    PROF.timed("lastconv_bwd_tc", 0.0, lambda: check(cabi.lib().dfl_lastconv_bwd(
        _p(s), _p(dout), _p(w), _p(mask_src), _p(ds), _p(ds_masked), _p(dw), _p(db), d, nd, w.shape[-1], _st())))


def lastconv_curl_loss_workspace(device):
    """zeroed workspace of the fused kernel (per-CTA loss partials + the CTA ticket); allocate ONCE per trainer"""
    return torch.zeros(cabi.lib().dfl_lastconv_curl_loss_workspace_bytes(), dtype=torch.uint8, device=device)


def lastconv_curl_loss_bwd(s, pot, x, w, mask_src, ds, ds_masked, dw, db, loss3, workspace, w1=1.0, w2=1.0, grad_scale=1.0,
                           dpot=None, vel=None):
    """FUSED curl + Jacobian-L1 loss + adjoints + output-conv backward: see dfl_lastconv_curl_loss_bwd.
    3D: pot, x [B,D,H,W,3];  2D: pot [B,H,W,1] (stream function), x [B,H,W,2]"""
    d, nd = _spatial(s)
    assert pot.dtype == torch.float32 and x.dtype == torch.float32
    assert (nd == 3 and pot.shape[-1] == 3 and x.shape[-1] == 3) or (nd == 2 and pot.shape[-1] == 1 and x.shape[-1] == 2)
    nvox = x.numel() // x.shape[-1]
    # algorithmic bytes: read s + mask, write ds + ds_masked (bf16 x 128), read the potential and the target
    work = nvox * (128 * 2 * 4 + 4 * (pot.shape[-1] + x.shape[-1]))
    PROF.timed("lastconv_bwd_fused", work, lambda: check(cabi.lib().dfl_lastconv_curl_loss_bwd(
        _p(s), _p(pot), _p(x), _p(w), _p(mask_src), _p(ds), _p(ds_masked), _p(dw), _p(db), _p(dpot), _p(vel), _p(loss3),
        _p(workspace), d, nd, float(w1), float(w2), float(grad_scale), _st())))


def upscale2(x):
    """nearest x2 of a channels-last tensor [B,(D,)H,W,C] (fp32 / bf16)"""
    d, nd = _spatial(x)
    out = torch.empty((x.shape[0],) + tuple(2 * int(e) for e in x.shape[1:-1]) + (x.shape[-1],), dtype=x.dtype, device=x.device)
    PROF.timed("upscale2", 0.0, lambda: check(cabi.lib().dfl_upscale2(_p(x), _p(out), d, nd, x.shape[-1], _dt(x), _st())))
    return out


def pool2(g):
    """adjoint of upscale2: sum of the 2^nd children"""
    nd = g.dim() - 2
    out = torch.empty((g.shape[0],) + tuple(int(e) // 2 for e in g.shape[1:-1]) + (g.shape[-1],), dtype=g.dtype, device=g.device)
    d, _ = _spatial(out)
    PROF.timed("pool2", 0.0, lambda: check(cabi.lib().dfl_pool2(_p(g), _p(out), d, nd, g.shape[-1], _dt(g), _st())))
    return out


def pool_mask(g, mask_src, ds, dmasked, addend=None):
    """g: fine-grid gradient [B,(2D,)2H,2W,128]; ds / dmasked: coarse [B,(D,)H,W,128]; addend: optional coarse term"""
    ref = ds if ds is not None else dmasked
    d, nd = _spatial(ref)
    if addend is None:
        PROF.timed("pool_mask", 0.0, lambda: check(cabi.lib().dfl_pool_mask(_p(g), _p(mask_src), _p(ds), _p(dmasked), d, nd, _st())))
    else:
        PROF.timed("pool_mask", 0.0, lambda: check(cabi.lib().dfl_pool_mask_add(_p(g), _p(addend), _p(mask_src), _p(ds), _p(dmasked), d, nd, _st())))


# ------------------------------------------------------------------ optimizer / misc
def adam_step_dev(param, grad, m, v, lr_t_dev, beta1, beta2, eps=1e-8, grad_scale=1.0):
    """Adam / GD with the step size in device memory (CUDA-graph replayable)."""
    PROF.timed("adam_step_dev", 0.0, lambda: check(cabi.lib().dfl_adam_step_dev(_p(param), _p(grad), _p(m), _p(v), param.numel(), _p(lr_t_dev), beta1, beta2, eps,
                                       grad_scale, _st())))


def adam_step(param, grad, m, v, lr_t, beta1, beta2, eps=1e-8, grad_scale=1.0):
    PROF.timed("adam_step", 0.0, lambda: check(cabi.lib().dfl_adam_step(_p(param), _p(grad), _p(m), _p(v), param.numel(), lr_t, beta1, beta2, eps,
                                   grad_scale, _st())))


def cast_f32_bf16(a, out=None):
    if out is None:
        out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    PROF.timed("cast_f32_bf16", 0.0, lambda: check(cabi.lib().dfl_cast_f32_bf16(_p(a), _p(out), a.numel(), _st())))
    return out
