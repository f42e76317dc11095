"""__graft_entry__.smoke(): one tiny generator train step on cuda:0, checked against the CPU oracle."""
import torch


def run_smoke():
    from . import kernels as K
    from .engine import GeneratorEngine
    from oracle import ref_model as M          # checker only (allowed importer: smoke)
    from oracle import ref_train as T

    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    B, spatial = 2, [16, 16, 16]
    eng = GeneratorEngine(B, spatial + [3], z_dim=3, num_conv=2, device=dev, seed=7)
    x, y = T.synthetic_batch(B, spatial, seed=5)
    pot = eng.forward(y.to(dev))
    # the loss stencil (curl, Jacobians, L1 means, adjoints) runs in the PROLOGUE of the output conv's backward kernel:
    # between lastconv_fwd_tc_kernel and lastconv_bwd_fused_kernel nothing is launched
    loss3 = torch.zeros(3, dtype=torch.float32, device=dev)
    eng.zero_grad()
    eng.backward(None, fused=dict(x=x.to(dev), w1=1.0, w2=1.0, loss3=loss3, workspace=K.lastconv_curl_loss_workspace(dev)))
    torch.cuda.synchronize()
    var = eng.params.state_dict()
    loss, _, _, _, pot_ref, grads = T.generator_loss_and_grads(y, x, var, num_conv=2)
    e_pot = float((pot.cpu() - pot_ref).norm() / pot_ref.norm())
    e_loss = abs(loss3[0].item() - loss.item()) / abs(loss.item())
    k = "G/1_conv/weights"
    e_g = float((eng.params.g(k).cpu() - grads[k]).norm() / grads[k].norm())
    print("smoke: rel-L2(potential)=%.2e  rel(loss)=%.2e  rel-L2(dW1)=%.2e" % (e_pot, e_loss, e_g))
    assert e_pot < 3e-2 and e_loss < 3e-2 and e_g < 1e-1, "smoke parity failed"
    eng.adam_step(1e-4)
    torch.cuda.synchronize()
    print("smoke OK")
