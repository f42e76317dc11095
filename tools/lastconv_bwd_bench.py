"""Times the first backward kernel(s) of the 3D step alone (CUDA events, L2 flushed between launches) at 4 x 128^3 and
16 x 64^3:   fused  = dfl_lastconv_curl_loss_bwd (loss stencil in the prologue of the output conv's backward)
             pair   = dfl_stencil_loss_fwdbwd (+ finalize) then dfl_lastconv_bwd
and prints algorithmic GB/s (read s + mask, write ds + ds_masked: 4 x 256 B / voxel; read A + x: 24 B / voxel) against
MEASURED_PEAKS.json's HBM number.    python tools/lastconv_bwd_bench.py [--profile fused|pair] [--json out.json]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfluids_b200 import kernels as K  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run(shape, mode, iters=8, warm=2):
    d = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    pot = torch.randn(*shape, 3, device=d, generator=g) * 0.05
    x = torch.randn(*shape, 3, device=d, generator=g) * 0.05
    s = (torch.randn(*shape, 128, device=d, generator=g) * 0.5).bfloat16()
    mask = torch.randn(*shape, 128, device=d, generator=g).bfloat16()
    w = (torch.randn(3, 3, 3, 128, 3, device=d, generator=g) * 0.05)
    ds, dsm = torch.empty_like(s), torch.empty_like(s)
    dw, db, l3 = torch.zeros_like(w), torch.zeros(3, device=d), torch.zeros(3, device=d)
    dpot = torch.empty_like(pot)
    ws = K.lastconv_curl_loss_workspace(d)
    wss = torch.empty(1 << 16, dtype=torch.uint8, device=d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)

    def once():
        if mode == "fused":
            K.lastconv_curl_loss_bwd(s, pot, x, w, mask, ds, dsm, dw, db, l3, ws)
        else:
            K.stencil_loss_fwdbwd(pot, x, dpot=dpot, loss3=l3, workspace=wss)
            K.lastconv_bwd(s, dpot, w, mask, ds, dsm, dw, db)

    ts = []
    for i in range(warm + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        once()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    t = ts[len(ts) // 2]
    nvox = pot.numel() // 3
    return t, nvox * (4 * 256 + 24) / t / 1e9


def main():
    peak = peak_gbs()
    if "--profile" in sys.argv:
        mode = sys.argv[sys.argv.index("--profile") + 1]
        run((4, 128, 128, 128), mode, iters=1, warm=1)
        return
    out = {"peak_gbs": peak, "cases": []}
    for name, shape in (("c4 128^3 B4", (4, 128, 128, 128)), ("c3 64^3 B16", (16, 64, 64, 64))):
        row = {"case": name}
        for mode in (("fused",) if "--fused-only" in sys.argv else ("fused", "pair")):
            t, gbs = run(shape, mode)
            row[mode] = {"ms": t * 1e3, "algorithmic_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
        out["cases"].append(row)
        print("%-12s " % name + "   ".join("%s %.3f ms (%.0f GB/s, %.3f of HBM)" % (
            m, row[m]["ms"], row[m]["algorithmic_gbs"], row[m]["frac_of_hbm_peak"]) for m in ("fused", "pair") if m in row))
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
