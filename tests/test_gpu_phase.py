"""Phase-decomposed upsample-conv (DFL_PHASE_UPCONV=1; reference model.py:76-79 followed by :67-69): the first conv of every
block after the first reads upscale(s), so per output phase it is a 2x2(x2)-tap convolution on the COARSE tensor with
pre-summed weights (8/27 of the dense FLOPs in 3D).  The identity itself is asserted on the oracle in
tests/test_oracle.py::test_upsample_conv_equals_phase_conv; here the kernels:
  * the phase forward == the dense tensor-core conv on the up-sampled tensor (same bf16 operands; the summed weights are
    rounded once instead of three times, so agreement is at bf16-rounding level, not bit level);
  * the whole generator with the flag: potential / loss vs the fp32 oracle, teacher-forced gradients of every layer vs the
    oracle (<= 2e-2 weights / 5e-2 biases: the bounds of the dense path), and vs the dense-path engine on the same weights.
"""
import os
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


def _engine(phase, *a, **k):
    from deepfluids_b200.engine import GeneratorEngine
    os.environ["DFL_PHASE_UPCONV"] = "1" if phase else "0"
    try:
        return GeneratorEngine(*a, **k)
    finally:
        os.environ.pop("DFL_PHASE_UPCONV", None)


@pytest.mark.parametrize("spatial,num_conv,B", [([32, 24], 2, 2), ([16, 16, 32], 2, 2), ([64, 48], 4, 1), ([32, 32, 32], 4, 1)])
def test_phase_generator_vs_oracle_and_dense_path(spatial, num_conv, B):
    from deepfluids_b200 import kernels as K
    nd = len(spatial)
    cout = 3 if nd == 3 else 1
    eng = _engine(True, B, spatial + [cout], z_dim=3, num_conv=num_conv, device=dev(), seed=17)
    ref_eng = _engine(False, B, spatial + [cout], z_dim=3, num_conv=num_conv, device=dev(), seed=17)
    assert eng.phase and not ref_eng.phase and torch.equal(eng.params.data, ref_eng.params.data)
    x, y = T.synthetic_batch(B, spatial, seed=5)
    outs = []
    for e in (eng, ref_eng):
        pot = e.forward(y.to(dev()))
        loss3, dpot, _ = K.stencil_loss_fwdbwd(pot, x.to(dev()))
        e.zero_grad()
        e.backward(dpot)
        outs.append((pot.clone(), loss3.clone(), dpot.clone()))
    var = eng.params.state_dict()
    pot_ref = M.generator_forward(y, var, spatial + [cout], num_conv=num_conv)
    loss_ref = T.stencil_loss(pot_ref, x)[0]
    e_pot, e_dense = rel_l2(outs[0][0], pot_ref), rel_l2(outs[0][0], outs[1][0])
    assert e_pot <= 1e-2 and e_dense <= 1e-2, (e_pot, e_dense)
    assert abs(outs[0][1][0].item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item())
    # teacher-forced backward: every layer fed with what the device stored (the oracle runs the DENSE layer on upscale(s))
    acts = {"x0": [t.float().cpu() for t in eng.x0], "y": [[t.float().cpu() for t in row] for row in eng.y], "s": eng.s.float().cpu()}
    # (lrelu masks from the STORED outputs: the phase layers' pre-activations differ from a dense re-evaluation at the level
    # of the weights' bf16 rounding -- summed-then-rounded vs rounded-then-summed -- which would flip ~1e-3 of the re-computed
    # signs, each flip a 5x change of that element)
    tf = T.teacher_forced_backward(y, var, acts, outs[0][2].cpu(), num_conv=num_conv, operand_round=M.bf16_round_ste,
                                   mask_from_acts=True)
    errs = OrderedDict((k, rel_l2(eng.params.g(k), tf[k])) for k in list(var)[:-1])
    ew = max(v for k, v in errs.items() if k.endswith("weights"))
    eb = max(v for k, v in errs.items() if k.endswith("biases"))
    # and the dense-path engine's gradients (free-running against each other: same masks up to bf16 noise)
    ed = max(rel_l2(eng.params.g(k), ref_eng.params.g(k)) for k in var if k.endswith("weights"))
    print("phase %s nc=%d: pot vs oracle %.2e vs dense %.2e | teacher-forced W %.2e b %.2e | vs dense grads %.2e" % (
        spatial, num_conv, e_pot, e_dense, ew, eb, ed))
    assert ew <= 1.2e-2 and eb <= 2e-2, errs      # measured <= 7.1e-3 / 1.0e-2
    assert ed <= 1.5e-1


@pytest.mark.parametrize("nd", [2, 3])
def test_phase_forward_kernel_equals_dense_conv_on_upsampled_input(nd):
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(3 + nd)
    coarse = [6, 8, 12] if nd == 3 else [10, 12]
    B, F = 2, 128
    s = (torch.randn([B] + coarse + [F], generator=g) * 0.5).bfloat16()
    w = R.xavier_uniform_((3,) * nd + (F, F), g)
    b = torch.randn(F, generator=g) * 0.1
    up = (R.upscale3 if nd == 3 else R.upscale)(s.float(), 2).bfloat16().to(dev()).contiguous()
    wf, _ = K.pack_conv_weights(w.to(dev()))
    dense = torch.empty_like(up)
    K.conv3x3(up, wf, b.to(dev()), out=dense, flags=K.CONV_LRELU)
    P = 2 ** nd
    wfp = torch.empty(P, F, P * F, dtype=torch.bfloat16, device=dev())
    wdp = torch.empty(F, P * P * F, dtype=torch.bfloat16, device=dev())
    K.pack_phase_weights(w.to(dev()), wfp, wdp)
    out = torch.full_like(up, float("nan"))
    fine = [2 * v for v in coarse]
    for r in range(P):
        rr = [(r >> (nd - 1 - a)) & 1 for a in range(nd)]
        taps = []
        for o in range(P):
            ob = [(o >> (nd - 1 - a)) & 1 for a in range(nd)]
            off = [(ob[a] - 1) if rr[a] == 0 else ob[a] for a in range(nd)]
            taps.append([0] * (3 - nd) + [2 * v for v in off] + [o * F])
        K.conv_taps(up, wfp[r], b.to(dev()), out, None, None, None, [B] + coarse, fine, F, 2, taps, 2, rr, flags=K.CONV_LRELU)
    assert torch.isfinite(out.float()).all()                      # every fine voxel written
    oracle = R.conv_nd(up.float().cpu(), w.bfloat16().float(), b, 1, R.lrelu)
    assert rel_l2(out, oracle) <= 6e-3 and rel_l2(out, dense) <= 6e-3, (rel_l2(out, oracle), rel_l2(out, dense))
    # data gradient: dS = pool-free coarse gradient of sum(out * gy) w.r.t. s, vs oracle autograd through upscale + conv
    gy = (torch.randn(up.shape, generator=g) * 0.1).bfloat16()
    taps = []
    for r in range(P):
        rr = [(r >> (nd - 1 - a)) & 1 for a in range(nd)]
        for o in range(P):
            ob = [(o >> (nd - 1 - a)) & 1 for a in range(nd)]
            off = [(ob[a] - 1) if rr[a] == 0 else ob[a] for a in range(nd)]
            taps.append([0] * (3 - nd) + [rr[a] - 2 * off[a] for a in range(nd)] + [(r * P + o) * F])
    ds = torch.empty_like(s, device=dev())
    K.conv_taps(gy.to(dev()), wdp, None, ds, None, None, None, [B] + coarse, coarse, F, 2, taps, 1, [0] * nd)
    sc = s.float().clone().requires_grad_(True)
    yc = R.conv_nd((R.upscale3 if nd == 3 else R.upscale)(sc, 2), w.bfloat16().float(), b, 1, None)
    (gs,) = torch.autograd.grad(yc, sc, gy.float())
    assert rel_l2(ds, gs) <= 6e-3, rel_l2(ds, gs)
