"""CPU tests of arch=nn: the oracle (oracle/ref_nn.py) against the committed golden vectors generated under the tf shim
from the reference's own model.NN / Trainer.build_model_nn (oracle/make_golden_nn.py), an independent witness for the
batch-norm restatement (torch.nn.functional.batch_norm), and the host-side data pipeline (data_nn.BatchManager mirrors
reference data_nn.py:12-168)."""
import argparse
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_nn as N

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from nn_helpers import _golden, _write_codes, _cfg  # noqa: E402

def test_oracle_matches_reference_wiring_golden():
    z, (B, F, Z, P, W), var, masks = _golden()
    assert list(var.keys()) == list(N.nn_layout(Z + P, F, Z).keys())
    ovar = {k: v.clone() for k, v in var.items()}
    y_ = N.nn_forward(torch.from_numpy(z["in/x"]), ovar, True, masks[0], keep_prob=float(z["keep"]))
    loss, grads, yw_ = N.nn_loss_and_grads(torch.from_numpy(z["in/xw"]), torch.from_numpy(z["in/yw"]), ovar, P, float(z["rescale"]),
                                           masks[1:], keep_prob=float(z["keep"]))
    assert np.array_equal(y_.numpy(), z["y_"]) and np.array_equal(yw_.numpy(), z["yw_"])
    assert float(loss) == float(z["loss"])
    for k, g in grads.items():
        assert np.allclose(g.numpy(), z["grad/" + k], rtol=0, atol=1e-6 * max(np.abs(z["grad/" + k]).max(), 1e-30)), k
    for k in ovar:
        if not N.is_trainable(k):
            assert np.array_equal(ovar[k].numpy(), z["stats_after/" + k]), k
    ytw_ = N.rollout(torch.from_numpy(z["in/xtw"]), ovar, P, float(z["rescale"]), False)
    assert np.array_equal(ytw_.numpy(), z["ytw_"])


@pytest.mark.parametrize("B,n", [(7, 5), (2, 3), (64, 32)])
def test_batch_norm_restatement_vs_torch_witness(B, n):
    g = torch.Generator().manual_seed(B * 100 + n)
    x = torch.randn(B, n, generator=g, dtype=torch.float64).requires_grad_(True)
    gamma = (1 + 0.3 * torch.randn(n, generator=g, dtype=torch.float64)).requires_grad_(True)
    beta = torch.randn(n, generator=g, dtype=torch.float64).requires_grad_(True)
    mm, mv = torch.randn(n, generator=g, dtype=torch.float64), 1 + torch.rand(n, generator=g, dtype=torch.float64)
    mm2, mv2 = mm.clone(), mv.clone()
    dy = torch.randn(B, n, generator=g, dtype=torch.float64)
    y = N.batch_norm(x, gamma, beta, mm, mv, True, 1e-5, 0.9, act=None)
    gx = torch.autograd.grad(y, [x, gamma, beta], dy)
    # torch's momentum is the weight of the NEW statistic: 1 - decay
    yw = torch.nn.functional.batch_norm(x, mm2, mv2, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    gw = torch.autograd.grad(yw, [x, gamma, beta], dy)
    assert torch.allclose(y, yw, atol=1e-12) and torch.allclose(mm, mm2, atol=1e-12) and torch.allclose(mv, mv2, atol=1e-12)
    for a, b in zip(gx, gw):
        assert torch.allclose(a, b, atol=1e-10)
    yi = N.batch_norm(x, gamma, beta, mm, mv, False, 1e-5, 0.9, act=None)
    yiw = torch.nn.functional.batch_norm(x, mm2, mv2, gamma, beta, training=False, eps=1e-5)
    assert torch.allclose(yi, yiw, atol=1e-12)


def test_dropout_mask_restatement_statistics_and_offsets():
    m = N.dropout_mask(123, 0, (1000, 100), 0.1)
    assert abs(float(m.float().mean()) - 0.1) < 3e-3
    # the stream is indexed by (seed, offset + element): a later window of the same stream is the tail of the earlier one
    a = N.dropout_mask(7, 0, (50,), 0.3)
    b = N.dropout_mask(7, 20, (30,), 0.3)
    assert torch.equal(a[20:], b)
    assert not torch.equal(N.dropout_mask(8, 0, (50,), 0.3), a)
    x = torch.randn(1000, 100)
    y = N.dropout(x, 0.1, m)
    assert torch.equal(y[m], (x / 0.1)[m]) and float(y[~m].abs().max()) == 0.0


