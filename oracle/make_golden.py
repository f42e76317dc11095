"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz IN THE BUILD CONTAINER.

Runs the reference's own `ops.py` (imported read-only from /root/reference through oracle/tf_shim.py,
source unchanged) on seeded inputs and stores inputs + outputs as small fixtures, so the pinned
results travel to the GPU box where /root/reference does not exist.  Also asserts, while
generating, that `oracle/ref_ops.py` reproduces the reference bit-exactly and that the reference's
numpy twins (ops.py:305-374) agree -- i.e. this script is what pins the oracle.

    python -m oracle.make_golden            # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import tf_shim  # noqa: E402
from oracle import ref_ops as R  # noqa: E402
from oracle import ref_train as T  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def main(reference_root="/root/reference", out_dir=os.path.join(ROOT, "tests", "golden")):
    os.makedirs(out_dir, exist_ok=True)
    ops = tf_shim.import_reference_ops(reference_root)
    g = torch.Generator().manual_seed(20260925)

    # ---------------- 2D stencils: curl, jacobian, divergence, lrelu, upscale ----------------
    B, H, W = 2, 12, 10
    psi = torch.randn(B, H, W, 1, generator=g)
    vel = torch.randn(B, H, W, 2, generator=g)
    c_ref = ops.curl(psi)
    j_ref, w_ref = ops.jacobian(vel, "NHWC")
    d_ref = ops.divergence(c_ref)
    l_ref = ops.lrelu(vel)
    u_ref = ops.upscale(vel, 2)
    assert torch.equal(c_ref, R.curl(psi))
    jj, ww = R.jacobian(vel)
    assert torch.equal(j_ref, jj) and torch.equal(w_ref, ww)
    assert torch.equal(d_ref, R.divergence(c_ref))
    assert torch.equal(l_ref, R.lrelu(vel))
    assert torch.equal(u_ref, R.upscale(vel, 2))
    # the reference's numpy twins (ops.py:305-324) are a second witness
    assert np.array_equal(ops.curl_np(_np(psi)), _np(c_ref))
    assert np.array_equal(ops.vort_np(_np(vel)), _np(w_ref))
    assert float(d_ref.abs().max()) <= 1e-5

    # loss + autograd gradient wrt the stream function, reference ops verbatim (trainer.py:140-172)
    x2 = ops.curl(torch.randn(B, H, W, 1, generator=g))
    x2 = x2 / x2.abs().max()
    p = psi.clone().requires_grad_(True)
    G_ = ops.curl(p)
    l1 = torch.mean(torch.abs(G_ - x2))
    jl1 = torch.mean(torch.abs(ops.jacobian(G_, "NHWC")[0] - ops.jacobian(x2, "NHWC")[0]))
    loss = l1 * 1.0 + jl1 * 1.0
    (dpsi,) = torch.autograd.grad(loss, p)
    lo, l1o, jl1o, go = T.stencil_loss(psi, x2)
    assert torch.equal(go, G_.detach()) and float(lo) == float(loss)
    np.savez_compressed(os.path.join(out_dir, "stencil2d.npz"),
                        psi=_np(psi), vel=_np(vel), curl=_np(c_ref), jac=_np(j_ref), vort=_np(w_ref),
                        div_of_curl=_np(d_ref), lrelu=_np(l_ref), upscale=_np(u_ref),
                        x=_np(x2), loss=np.float32(loss.item()), loss_l1=np.float32(l1.item()),
                        loss_j_l1=np.float32(jl1.item()), dpsi=_np(dpsi))

    # ---------------- 3D stencils ----------------
    B, D, H, W = 2, 6, 10, 8
    A = torch.randn(B, D, H, W, 3, generator=g)
    v3 = torch.randn(B, D, H, W, 3, generator=g)
    j3_ref, c3_ref = ops.jacobian3(v3)
    _, cA = ops.jacobian3(A)
    d3_ref = ops.divergence3(cA)
    u3_ref = ops.upscale3(v3, 2)
    j3, c3 = R.jacobian3(v3)
    assert torch.equal(j3_ref, j3) and torch.equal(c3_ref, c3)
    assert torch.equal(d3_ref, R.divergence3(cA))
    assert torch.equal(u3_ref, R.upscale3(v3, 2))
    jn, cn = ops.jacobian_np3(_np(v3))
    assert np.array_equal(jn, _np(j3_ref)) and np.array_equal(cn, _np(c3_ref))
    assert float(d3_ref.abs().max()) <= 1e-5

    x3 = ops.jacobian3(torch.randn(B, D, H, W, 3, generator=g))[1]
    x3 = x3 / x3.abs().max()
    a = A.clone().requires_grad_(True)
    _, G3 = ops.jacobian3(a)                                   # trainer3.py:18
    l1 = torch.mean(torch.abs(G3 - x3))                        # trainer3.py:49
    jl1 = torch.mean(torch.abs(ops.jacobian3(G3)[0] - ops.jacobian3(x3)[0]))
    loss = l1 * 1.0 + jl1 * 1.0
    (dA,) = torch.autograd.grad(loss, a)
    lo, _, _, go = T.stencil_loss(A, x3)
    assert torch.equal(go, G3.detach()) and float(lo) == float(loss)
    np.savez_compressed(os.path.join(out_dir, "stencil3d.npz"),
                        A=_np(A), vel=_np(v3), jac=_np(j3_ref), curl_of_vel=_np(c3_ref), curl_of_A=_np(cA),
                        div_of_curl=_np(d3_ref), upscale3=_np(u3_ref),
                        x=_np(x3), loss=np.float32(loss.item()), loss_l1=np.float32(l1.item()),
                        loss_j_l1=np.float32(jl1.item()), dA=_np(dA))
    # ---------------- boundary: the reference's flag surface (config.py:16-70), names and defaults ----------------
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("_dfl_reference_config", reference_root + "/config.py")
    cfgmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfgmod)
    ref_cfg, _ = cfgmod.parser.parse_known_args([])
    with open(os.path.join(out_dir, "reference_config_defaults.json"), "w") as f:
        json.dump(dict(sorted(vars(ref_cfg).items())), f, indent=1)
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main()
