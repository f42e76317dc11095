"""model-level `reuse=True` on the fused engines = tf.variable_scope(reuse=True) (reference model.py:15,57,121,158,192,206):
the SAME variables are applied again -- also at another batch size -- and an update of the variables is seen by every
engine of the scope (no snapshot copies)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("nd", [2, 3])
def test_generator_reuse_shares_the_variables_across_batch_sizes(nd):
    from deepfluids_b200 import model as M
    M.reset()
    gen = M.GeneratorBE if nd == 2 else M.GeneratorBE3
    shape = [32, 48, 1] if nd == 2 else [16, 16, 16, 3]
    g = torch.Generator().manual_seed(0)
    z = (torch.rand(4, 3, generator=g) * 2 - 1).to(dev())
    out4, names = gen(z, 128, shape, num_conv=2)
    out4 = out4.clone()
    out2, names2 = gen(z[:2], 128, shape, num_conv=2, reuse=True)
    out2 = out2.clone()                  # (the engines return their output buffer)
    assert names2 == names
    e4, e2 = M.get_engine("G", nd), M._ENGINES[("G", nd, ("batch", 2))]
    assert e2.params is e4.params, "reuse=True must share the flat parameter buffer, not copy it"
    assert rel_l2(out2, out4[:2]) < 1e-3
    # the variables change (an optimizer step, a checkpoint load): every engine of the scope sees it
    e4.params.data.mul_(1.25)
    out2b, _ = gen(z[:2], 128, shape, num_conv=2, reuse=True)
    out2b = out2b.clone()
    out4b, _ = gen(z, 128, shape, num_conv=2, reuse=True)
    assert rel_l2(out2b, out4b[:2]) < 1e-3
    assert rel_l2(out2b, out2) > 1e-2
    # reuse=False creates fresh variables and drops the siblings of the old ones
    gen(z, 128, shape, num_conv=2)
    assert ("G", nd, ("batch", 2)) not in M._ENGINES
    M.reset()


def test_encoder_and_ae_reuse_share_the_variables():
    from deepfluids_b200 import model as M
    M.reset()
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(4, 32, 48, 2, generator=g) * 2 - 1).to(dev())
    z4, _ = M.EncoderBE(x, 128, 16, num_conv=2)
    z4 = z4.clone()
    z2, _ = M.EncoderBE(x[:2].contiguous(), 128, 16, num_conv=2, reuse=True)
    assert M._ENGINES[("enc", 2, "enc", ("batch", 2))].params is M.get_engine("enc", 2, "enc").params
    assert rel_l2(z2, z4[:2]) < 2e-3
    out4, c4, _ = M.AE(x, 128, 16, num_conv=2)
    out4, c4 = out4.clone(), c4.clone()
    out2, c2, _ = M.AE(x[:2].contiguous(), 128, 16, num_conv=2, reuse=True)
    assert M._ENGINES[("AE", 2, "ae", ("batch", 2))].params is M.get_engine("AE", 2, "ae").params
    assert rel_l2(c2, c4[:2]) < 2e-3 and rel_l2(out2, out4[:2]) < 5e-3
    M.reset()
