"""Probe: does the relative placement of the four 2 GB streams of the fused first-backward kernel (s, mask read; ds, ds_masked
written) change its speed (DRAM bank / channel conflicts between streams at the same relative offset)?  Also times plain torch
elementwise kernels with 2-4 streams as a calibration of what the memory system gives to multi-stream access.
    python tools/hbm_pattern_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfluids_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=6, warm=2):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(warm + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    d = torch.device("cuda:0")
    shape = (4, 128, 128, 128)
    n = 4 * 128 ** 3 * 128
    g = torch.Generator(device="cuda").manual_seed(1)
    pot = torch.randn(*shape, 3, device=d, generator=g) * 0.05
    x = torch.randn(*shape, 3, device=d, generator=g) * 0.05
    w = torch.randn(3, 3, 3, 128, 3, device=d, generator=g) * 0.05
    dw, db, l3 = torch.zeros_like(w), torch.zeros(3, device=d), torch.zeros(3, device=d)
    ws = K.lastconv_curl_loss_workspace(d)
    for skew in (0, 4096 + 256, 65536 + 1024, (1 << 20) + 8192 + 512, (3 << 20) + 32768 + 2048):
        pool = torch.empty(4 * (n + (8 << 20)), dtype=torch.bfloat16, device=d)
        views = []
        for k in range(4):
            off = k * (n + (4 << 20)) + k * (skew // 2)
            views.append(pool[off:off + n].view(*shape, 128))
        s, mask, ds, dsm = views
        s.normal_(0, 0.5); mask.normal_()
        t = timeit(lambda: K.lastconv_curl_loss_bwd(s, pot, x, w, mask, ds, dsm, dw, db, l3, ws))
        print("fused, stream skew %8d B: %.3f ms  %.0f GB/s" % (skew, t * 1e3, (n // 128) * 1048 / t / 1e9), flush=True)
        del pool, views, s, mask, ds, dsm
    a = torch.randn(n, device=d, dtype=torch.bfloat16); b = torch.randn_like(a); c = torch.empty_like(a); e = torch.empty_like(a)
    t = timeit(lambda: c.copy_(a)); print("copy          (1r 1w): %.3f ms %.0f GB/s" % (t * 1e3, 2 * n * 2 / t / 1e9))
    t = timeit(lambda: torch.add(a, b, out=c)); print("add           (2r 1w): %.3f ms %.0f GB/s" % (t * 1e3, 3 * n * 2 / t / 1e9))
    t = timeit(lambda: torch._foreach_copy_([c, e], [a, b])); print("foreach copy  (2r 2w): %.3f ms %.0f GB/s" % (t * 1e3, 4 * n * 2 / t / 1e9))


if __name__ == "__main__":
    main()
