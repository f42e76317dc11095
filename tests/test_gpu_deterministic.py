"""Deterministic weight / bias gradients (dfl_set_deterministic): the split-K partial sums of the tensor-core weight-gradient
kernel and of the bias-gradient kernel are reduced in a fixed order instead of fp32 atomics.

  * bit-identical results over repeated launches for every operand mode of the kernel (3D / 2D brick mode, stride-2 tap-list
    mode, the 4^nd-tap phase correlation, the split-operand fp32-grade mode, sizes with empty slabs);
  * same values as the atomic reduction up to fp32 summation order (1e-5 relative to max|dw|) and as oracle autograd (1e-4
    rel-L2, the tolerance of the atomic path);
  * engine level (DFL_DETERMINISTIC=1): two backward passes of the generator from the same state give bit-identical
    gradients for all 128 -> 128 conv layers, phase-decomposed ones included.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_ops as R


def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture
def deterministic():
    from deepfluids_b200 import kernels as K
    K.set_deterministic(True, dev())
    yield K
    K.set_deterministic(False)


def _case(shape, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, 128, generator=g).bfloat16()
    dy = (torch.randn(*shape, 128, generator=g) * 0.1).bfloat16()
    return x, dy


@pytest.mark.parametrize("shape", [(2, 16, 16, 16), (1, 5, 6, 9), (3, 40, 24), (1, 2, 2, 2), (2, 32, 32, 32)])
def test_wgrad_deterministic_bitwise_and_equal_to_atomic_path(shape, deterministic):
    K = deterministic
    nd = len(shape) - 1
    x, dy = _case(shape, 100 + sum(shape))
    xd, dyd = x.to(dev()), dy.to(dev())
    runs = []
    for _ in range(3):
        dw = torch.zeros((3,) * nd + (128, 128), device=dev())
        db = torch.zeros(128, device=dev())
        db2 = torch.zeros(128, device=dev())
        K.conv3x3_wgrad(xd, dyd, dw, db)
        K.bias_grad(dyd, db2)
        runs.append((dw, db, db2))
    for dw, db, db2 in runs[1:]:
        assert torch.equal(dw, runs[0][0]) and torch.equal(db, runs[0][1]) and torch.equal(db2, runs[0][2])
    # accumulate semantics: a second launch into the same buffers doubles them exactly
    dw, db = runs[0][0].clone(), runs[0][1].clone()
    K.conv3x3_wgrad(xd, dyd, dw, db)
    assert torch.equal(dw, 2 * runs[0][0]) and torch.equal(db, 2 * runs[0][1])
    K.set_deterministic(False)
    dwa = torch.zeros_like(dw)
    dba = torch.zeros(128, device=dev())
    K.conv3x3_wgrad(xd, dyd, dwa, dba)
    K.set_deterministic(True, dev())
    assert float((dwa - runs[0][0]).abs().max()) <= 1e-5 * float(dwa.abs().max())
    assert float((dba - runs[0][1]).abs().max()) <= 1e-5 * float(dba.abs().max())
    if x.numel() <= 2 * 16 ** 3 * 128:
        wt = torch.zeros((3,) * nd + (128, 128), requires_grad=True)
        bt = torch.zeros(128, requires_grad=True)
        y = R.conv_nd(x.float(), wt, bt, 1, None)
        gw, gb = torch.autograd.grad(y, [wt, bt], dy.float())
        assert rel_l2(runs[0][0], gw) <= 1e-4 and rel_l2(runs[0][1], gb) <= 1e-4 and rel_l2(runs[0][2], gb) <= 1e-4


def test_wgrad_deterministic_stride2_phase_and_split_modes(deterministic):
    K = deterministic
    g = torch.Generator().manual_seed(5)
    # stride-2 convolution's weight gradient (tap-list mode, in_stride 2): x on the fine grid, dpre on the coarse grid
    xf = torch.randn(2, 16, 16, 16, 128, generator=g).bfloat16().to(dev())
    dc = (torch.randn(2, 8, 8, 8, 128, generator=g) * 0.1).bfloat16().to(dev())
    outs = []
    for _ in range(2):
        dw = torch.zeros(3, 3, 3, 128, 128, device=dev())
        db = torch.zeros(128, device=dev())
        K.conv_wgrad_ex(xf, dc, dw, db, 2, 0, 128 * 128, 128)
        outs.append((dw, db))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    K.set_deterministic(False)
    dwa, dba = torch.zeros_like(outs[0][0]), torch.zeros(128, device=dev())
    K.conv_wgrad_ex(xf, dc, dwa, dba, 2, 0, 128 * 128, 128)
    K.set_deterministic(True, dev())
    assert float((dwa - outs[0][0]).abs().max()) <= 1e-5 * float(dwa.abs().max())
    assert float((dba - outs[0][1]).abs().max()) <= 1e-5 * float(dba.abs().max())
    # phase-decomposed upsample-conv: 4^3-tap stride-2 correlation + fold
    outs = []
    for _ in range(2):
        dw = torch.zeros(3, 3, 3, 128, 128, device=dev())
        t = torch.empty(64, 128, 128, device=dev())
        K.phase_wgrad(xf, dc, t, dw)
        outs.append(dw)
    assert torch.equal(outs[0], outs[1])
    K.set_deterministic(False)
    dwa = torch.zeros_like(outs[0])
    K.phase_wgrad(xf, dc, torch.empty(64, 128, 128, device=dev()), dwa)
    K.set_deterministic(True, dev())
    assert float((dwa - outs[0]).abs().max()) <= 1e-5 * float(dwa.abs().max())
    # split-operand (fp32-grade) mode: (hi, lo) pairs stacked along the batch axis, three operand combinations
    x2 = torch.randn(2 * 2, 12, 20, 128, generator=g).bfloat16().to(dev())
    d2 = (torch.randn(2 * 2, 12, 20, 128, generator=g) * 0.1).bfloat16().to(dev())
    outs = []
    for _ in range(2):
        dw = torch.zeros(3, 3, 128, 128, device=dev())
        db = torch.zeros(128, device=dev())
        K.conv3x3_wgrad_split(x2, d2, dw, db)
        outs.append((dw, db))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    K.set_deterministic(False)
    dwa, dba = torch.zeros_like(outs[0][0]), torch.zeros(128, device=dev())
    K.conv3x3_wgrad_split(x2, d2, dwa, dba)
    K.set_deterministic(True, dev())
    assert float((dwa - outs[0][0]).abs().max()) <= 1e-5 * float(dwa.abs().max())
    assert float((dba - outs[0][1]).abs().max()) <= 1e-5 * float(dba.abs().max())


def test_set_deterministic_rejects_a_small_workspace():
    from deepfluids_b200 import cabi
    import ctypes as C
    lib = cabi.lib()
    need = int(lib.dfl_deterministic_workspace_bytes())
    assert need > 0
    buf = torch.empty(1024, dtype=torch.uint8, device=dev())
    assert lib.dfl_set_deterministic(C.c_void_p(buf.data_ptr()), 1024) != 0
    assert b"workspace" in lib.dfl_last_error()
    assert lib.dfl_set_deterministic(None, 0) == 0


def test_engine_backward_deterministic(monkeypatch):
    """generator backward twice from the same state: the whole flat gradient buffer bit-identical (3D, two levels, so the
    phase-decomposed layer, its bias_grad and the output conv's dW / db are on the path)"""
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.engine import GeneratorEngine
    monkeypatch.setenv("DFL_DETERMINISTIC", "1")
    try:
        eng = GeneratorEngine(2, [16, 16, 16, 3], z_dim=3, filters=128, num_conv=2, repeat=0, device=dev(), seed=3)
        assert K.deterministic() and eng._side is None
        g = torch.Generator().manual_seed(1)
        z = (torch.rand(2, 3, generator=g) * 2 - 1).to(dev())
        dpot = torch.randn(2, 16, 16, 16, 3, generator=g).to(dev()) * 0.01
        grads = []
        for _ in range(2):
            eng.zero_grad()
            eng.forward(z)
            eng.backward(dpot.clone())
            grads.append(eng.params.grad.clone())
        # every variable: FC, the 128 -> 128 layers (phase-decomposed one included) and the 128 -> 3 output conv
        assert torch.equal(grads[0], grads[1])
        for k in eng.params.table:
            if k.endswith("weights"):
                assert float(eng.params._view(grads[0], k).abs().max()) > 0, k
    finally:
        K.set_deterministic(False)


@pytest.mark.parametrize("is_3d", [True, False])
def test_training_runs_are_bit_reproducible(monkeypatch, is_3d):
    """DFL_DETERMINISTIC=1: two fresh trainers (same seed, same synthetic batches) land on bit-identical parameters after
    three optimizer steps -- through the fused first-backward kernel (3D) / the fused 2D output-conv backward, the
    tensor-core weight gradients, the phase-decomposed layers and the captured graph"""
    from deepfluids_b200 import config as C
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    from deepfluids_b200.trainer3 import Trainer3
    monkeypatch.setenv("DFL_DETERMINISTIC", "1")
    args = (["--synthetic=true", "--is_3d=true", "--res_x=32", "--res_y=16", "--res_z=16", "--batch_size=2", "--num_conv=2",
             "--max_step=20", "--lr_max=0.001"] if is_3d else
            ["--synthetic=true", "--is_3d=false", "--res_x=32", "--res_y=48", "--batch_size=4", "--num_conv=2", "--max_step=20",
             "--lr_max=0.001"])
    finals = []
    try:
        for _ in range(2):
            cfg, _u = C.get_config(args)
            bm = BatchManager(cfg, pool=2)
            tr = (Trainer3 if is_3d else Trainer)(cfg, bm)
            assert K.deterministic() and tr._fused_args(tr.x) is not None
            for _i in range(3):
                tr.train_step()
            torch.cuda.synchronize()
            finals.append((tr.engine.params.data.clone(), tr.losses()))
        assert torch.equal(finals[0][0], finals[1][0])
        assert finals[0][1] == finals[1][1]
    finally:
        K.set_deterministic(False)


@pytest.mark.parametrize("is_3d", [True, False])
def test_autoencoder_training_runs_are_bit_reproducible(monkeypatch, is_3d):
    """arch=ae under DFL_DETERMINISTIC=1: encoder (stride-2 weight gradients over channel blocks, the split-K encoder FC),
    decoder and the fused loss kernel -- two fresh trainers, three steps, identical parameters"""
    from deepfluids_b200 import config as C
    from deepfluids_b200 import kernels as K
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    from deepfluids_b200.trainer3 import Trainer3
    monkeypatch.setenv("DFL_DETERMINISTIC", "1")
    args = (["--synthetic=true", "--arch=ae", "--is_3d=true", "--res_x=16", "--res_y=16", "--res_z=16", "--batch_size=2",
             "--num_conv=2", "--max_step=20", "--lr_max=0.001"] if is_3d else
            ["--synthetic=true", "--arch=ae", "--is_3d=false", "--res_x=32", "--res_y=48", "--batch_size=4", "--num_conv=2",
             "--max_step=20", "--lr_max=0.001"])
    finals = []
    try:
        for _ in range(2):
            cfg, _u = C.get_config(args)
            bm = BatchManager(cfg, pool=2)
            tr = (Trainer3 if is_3d else Trainer)(cfg, bm)
            assert K.deterministic()
            for _i in range(3):
                tr.train_step()
            torch.cuda.synchronize()
            finals.append(tr.engine.params.data.clone())
        assert torch.equal(finals[0], finals[1])
    finally:
        K.set_deterministic(False)


def test_gemm_f32_split_k_is_off_in_deterministic_mode():
    """dfl_gemm_f32 splits K over atomics when M x N cannot fill the chip (the ops-level encoder FC: [B, 36 864+] x [.., 16]); under
    dfl_set_deterministic one CTA per output tile walks K in order: repeated launches are bit-identical and equal to the split
    result to fp32 summation accuracy"""
    from deepfluids_b200 import kernels as K
    g = torch.Generator().manual_seed(5)
    a = torch.randn(8, 36864, generator=g).to(dev())
    b = (torch.randn(36864, 16, generator=g) * 0.01).to(dev())
    bias = torch.randn(16, generator=g).to(dev())
    ref = (a.double() @ b.double() + bias.double()).float()
    split = K.gemm(a, b, bias)
    K.set_deterministic(True, dev())
    try:
        outs = [K.gemm(a, b, bias).clone() for _ in range(3)]
    finally:
        K.set_deterministic(False)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert float((outs[0] - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert float((split - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
