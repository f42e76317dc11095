"""Vendor bar (informative): the same generator train step -- FC, num_conv x [3x3(x3) conv + lrelu] per level,
residual add + nearest x2 upsample, output conv, curl, Jacobian-L1 loss, backward, Adam -- written in plain PyTorch
(cuDNN / cuBLAS kernels, bf16 autocast, channels-last) and timed the way bench.py times ours (CUDA events, warm-up,
synthetic data of the BASELINE shapes).  It is NOT the reference (TensorFlow 1.15) and none of this repo's kernels run in it;
it answers "what does the stock library path reach on this GPU".

    python tools/cudnn_bar.py --workload c4 [--steps 10 --warmup 3] [--device cuda|cpu] [--tiny]

Prints one JSON line {"impl": "torch-cudnn", "workload": ..., "value": fields/s, "ms_per_step": ...}."""
import argparse
import json
import math
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

WORKLOADS = {  # name: (spatial (D,)H,W, batch)  -- BASELINE.json configs[1..3] at the per-GPU batch bench.py uses
    "c2": ((128, 96), 64),
    "c3": ((64, 64, 64), 16),
    "c4": ((128, 128, 128), 4),
}


def lrelu(x):
    return F.leaky_relu(x, 0.2)                           # ops.py:9-10 (tf.maximum(x, 0.2 x)); the library's fused form


class Generator(nn.Module):
    """GeneratorBE / GeneratorBE3 (model.py:5-87), channels-first inside torch; filters = 128, num_conv = 4"""

    def __init__(self, spatial, cout, z_dim=3, filters=128, num_conv=4):
        super().__init__()
        self.nd = len(spatial)
        self.rep = int(math.log2(max(spatial))) - 2
        self.x0 = [s // 2 ** (self.rep - 1) for s in spatial]
        self.filters, self.num_conv = filters, num_conv
        conv = nn.Conv3d if self.nd == 3 else nn.Conv2d
        self.fc = nn.Linear(z_dim, int(torch.tensor(self.x0).prod()) * filters)
        self.convs = nn.ModuleList([conv(filters, filters, 3, padding=1) for _ in range(self.rep * num_conv)])
        self.last = conv(filters, cout, 3, padding=1)

    def forward(self, z):
        x = self.fc(z).view(z.shape[0], *self.x0, self.filters)
        x = x.permute(0, self.nd + 1, *range(1, self.nd + 1))                       # -> channels first
        x = x.contiguous(memory_format=torch.channels_last_3d if self.nd == 3 else torch.channels_last)
        x0 = x
        k = 0
        for i in range(self.rep):
            for _ in range(self.num_conv):
                x = lrelu(self.convs[k](x))
                k += 1
            x = x + x0
            if i < self.rep - 1:
                x = F.interpolate(x, scale_factor=2, mode="nearest")
                x0 = x
        return self.last(x)


def fdiff(f, axis):
    """replicate-last forward difference (ops.py:205-262)"""
    n = f.shape[axis]
    d = f.narrow(axis, 1, n - 1) - f.narrow(axis, 0, n - 1)
    return torch.cat([d, d.narrow(axis, n - 2, 1)], dim=axis)


def curl_and_jacobian(pot, x):
    """channels-first fp32: velocity G = curl(pot), loss = mean|G - x| + mean|J(G) - J(x)| (trainer.py:140-172, trainer3.py:18-51)"""
    nd = pot.dim() - 2
    if nd == 2:                                           # axes: 2 = y, 3 = x
        psi = pot[:, 0:1]
        G = torch.cat([fdiff(psi, 2), -fdiff(psi, 3)], dim=1)
        axes = (3, 2)
    else:                                                 # axes: 2 = z, 3 = y, 4 = x; A = (Au, Av, Aw)
        Au, Av, Aw = pot[:, 0:1], pot[:, 1:2], pot[:, 2:3]
        G = torch.cat([fdiff(Aw, 3) - fdiff(Av, 2), fdiff(Au, 2) - fdiff(Aw, 4), fdiff(Av, 4) - fdiff(Au, 3)], dim=1)
        axes = (4, 3, 2)
    jac = lambda v: torch.cat([fdiff(v, a) for a in axes], dim=1)
    return (G - x).abs().mean() + (jac(G) - jac(x)).abs().mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--tiny", action="store_true", help="shrink the grid 8x per axis and the batch to 2 (CPU smoke test of this script)")
    a = ap.parse_args()
    spatial, B = WORKLOADS[a.workload]
    if a.tiny:
        spatial, B = tuple(max(8, s // 8) for s in spatial), 2
    dev = torch.device(a.device)
    torch.backends.cudnn.benchmark = True                 # let the library pick its fastest algorithms
    nd = len(spatial)
    torch.manual_seed(123)
    g = Generator(spatial, 1 if nd == 2 else 3).to(dev)
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.999), eps=1e-8, fused=(dev.type == "cuda"))
    z = torch.rand(B, 3, device=dev) * 2 - 1
    x = torch.randn(B, nd, *spatial, device=dev).clamp_(-1, 1)
    amp = dev.type == "cuda"

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            pot = g(z)
        loss = curl_and_jacobian(pot.float(), x)
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
    else:
        t0 = time.time()
        for _ in range(a.steps):
            loss = step()
        ms = (time.time() - t0) * 1e3 / a.steps
    loss = float(loss.detach())
    assert loss == loss, "NaN loss"
    print(json.dumps({"impl": "torch-cudnn" if amp else "torch-cpu", "workload": a.workload, "tiny": a.tiny, "spatial": spatial, "batch": B,
                      "value": B / (ms * 1e-3), "unit": "fields/s", "ms_per_step": ms, "steps": a.steps, "warmup": a.warmup,
                      "dtype": "bf16 autocast, fp32 master weights" if amp else "f32", "params": sum(p.numel() for p in g.parameters())}))


if __name__ == "__main__":
    main()
