"""Times the fused 3D curl + Jacobian-L1 loss kernel alone (CUDA events, L2 flushed between launches) on the BASELINE
grids, fast path vs generic kernel, and prints achieved algorithmic GB/s (36 B / voxel: read A, read x, write dL/dA)
against MEASURED_PEAKS.json's HBM number.   python tools/stencil_bench.py [--json out.json]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfluids_b200 import kernels as K  # noqa: E402


def peak_gbs():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs", "hbm_copy_gbs", "hbm"):
            if k in d:
                v = d[k]
                return float(v["value"] if isinstance(v, dict) else v)
        return float(next(v for k, v in d.items() if "hbm" in k.lower() and isinstance(v, (int, float))))
    except Exception:
        return 6584.5


def time_one(shape, generic, iters=10):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(*shape, 3, device=dev, generator=g)
    x = torch.randn(*shape, 3, device=dev, generator=g)
    dA = torch.empty_like(A)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if generic:
        os.environ["DFL_STENCIL_GENERIC"] = "1"
    else:
        os.environ.pop("DFL_STENCIL_GENERIC", None)
    for _ in range(3):
        K.stencil_loss_fwdbwd(A, x, dpot=dA)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.stencil_loss_fwdbwd(A, x, dpot=dA)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    os.environ.pop("DFL_STENCIL_GENERIC", None)
    ts.sort()
    t = ts[len(ts) // 2]
    nbytes = A.numel() * 4 * 3
    return t, nbytes / t / 1e9


def main():
    peak = peak_gbs()
    out = {"peak_gbs": peak, "cases": []}
    cases = (("c3 64^3 B16", (16, 64, 64, 64)), ("c4 128^3 B4", (4, 128, 128, 128)), ("128^3 B1", (1, 128, 128, 128)))
    if "--profile" in sys.argv:          # one case, fast path only (for ncu)
        cases = cases[1:2]
    for name, shape in cases:
        for generic in ((False,) if "--profile" in sys.argv else (True, False)):
            t, gbs = time_one(shape, generic)
            rec = {"case": name, "kernel": "generic z-march" if generic else "lean (2 voxels/thread, persistent)",
                   "us": t * 1e6, "algorithmic_GBs": gbs, "frac_of_hbm_peak": gbs / peak}
            out["cases"].append(rec)
            print(rec, flush=True)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
