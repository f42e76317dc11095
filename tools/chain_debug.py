"""Where does the backward-chain error of the bf16 path come from?  Compares dL/d(pre-activation) of every conv layer
(GPU engine) against autograd on the oracle run with bf16 activation storage, driven by the same upstream gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
from deepfluids_b200 import kernels as K
from deepfluids_b200.engine import GeneratorEngine
from oracle import ref_model as M, ref_train as T, ref_ops as R

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())

dev = torch.device("cuda:0")
spatial, num_conv, B = [int(v) for v in os.environ.get("SP", "32,24").split(",")], int(os.environ.get("NC", "4")), 2
nd = len(spatial); cout = 3 if nd == 3 else 1
eng = GeneratorEngine(B, spatial + [cout], z_dim=3, num_conv=num_conv, device=dev, seed=11)
x, y = T.synthetic_batch(B, spatial, seed=3)
pot = eng.forward(y.to(dev))
loss3, dpot, _ = K.stencil_loss_fwdbwd(pot, x.to(dev))
eng.zero_grad(); eng.debug = OrderedDict(); eng.backward(dpot)
var = eng.params.state_dict()
leaves = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in var.items())
keep = []
pot_o = M.generator_forward(y, leaves, spatial + [cout], num_conv=num_conv, store=M.bf16_round_ste, keep=keep)
gy = torch.autograd.grad(pot_o, keep, dpot.cpu(), retain_graph=True)      # dL/dy_k (post-activation, stored)
gw = torch.autograd.grad(pot_o, list(leaves.values()), dpot.cpu())
names = [cn for row in eng.conv_names for cn in row]
print("pot rel", rel(pot, pot_o.detach()))
for k, (cn, yk, g) in enumerate(zip(names, keep, gy)):
    dpre_ref = g * torch.where(yk.detach() >= 0, 1.0, 0.2)
    mism = float((torch.sign(yk.detach()) != torch.sign(eng.y[k // num_conv][k % num_conv].float().cpu())).float().mean())
    print("%-12s dpre rel-L2 %.3e   y rel-L2 %.3e  sign(y) mismatch %.2e   dW rel %.3e" % (
        cn, rel(eng.debug[cn], dpre_ref), rel(eng.y[k // num_conv][k % num_conv], yk.detach()), mism,
        rel(eng.params.g(cn + "/weights"), gw[list(leaves).index(cn + "/weights")])))
