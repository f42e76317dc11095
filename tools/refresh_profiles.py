"""Copies the outputs of tools/final_measure.sh (gpurun_out/<NAME>/) into profiles/ under the round's names and re-derives the
two text summaries (kernel shares from the ncu launch list, ncu metrics of the fused first-backward kernel).  Runs here (no GPU).
    python tools/refresh_profiles.py r2final4 [--round r02]"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    name = sys.argv[1]
    rnd = sys.argv[sys.argv.index("--round") + 1] if "--round" in sys.argv else "r02"
    src = os.path.join(ROOT, "gpurun_out", name)
    dst = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(src)):
        if f.startswith("bench_") and f.endswith(".json"):
            shutil.copy(os.path.join(src, f), os.path.join(dst, "%s_final_%s" % (rnd, f)))
    shutil.copy(os.path.join(src, "launches_default_c4.csv"), os.path.join(dst, "%s_final_launches_default_c4.csv" % rnd))
    shutil.copy(os.path.join(src, "sanitizer_summary.txt"), os.path.join(dst, "%s_sanitizer_summary.txt" % rnd))
    shutil.copy(os.path.join(src, "lastconv_bwd_bench.json"), os.path.join(dst, "%s_final_lastconv_bwd_bench.json" % rnd))
    log = open(os.path.join(src, "pytest_gpu.log")).read().splitlines()
    keep = [l for l in log if re.search(r"passed|failed", l)][-1:]
    keep += [l.lstrip(".") for l in log if re.search(r"(64|128)\^3|stencil |recipe|phase \[|pot |AE worst|enc |dec W", l)]
    open(os.path.join(dst, "%s_final_pytest_gpu.txt" % rnd), "w").write("\n".join(keep) + "\n")

    # ---- kernel shares from the launch list
    rows = list(csv.reader(open(os.path.join(src, "launches_default_c4.csv"))))
    i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, order = collections.defaultdict(lambda: [0, 0.0]), []
    for r in rows[i0 + 1:]:
        if len(r) <= vi:
            continue
        v, u = float(r[vi].replace(",", "")), r[ui]
        ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki])).replace("dfl::", "")
        agg[short][0] += 1
        agg[short][1] += ms
        order.append(short)
    ours = {k: v for k, v in agg.items() if not k.startswith("at::") and "elementwise" not in k}
    tot = sum(v[1] for v in ours.values())
    out = ["ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu` (gpu__time_duration.sum, --clock-control none; "
           "cold-cache, serialised):",
           "aggregated over all launches of the library's kernels in the capture (%d launches, %.1f ms)" % (
               sum(v[0] for v in ours.values()), tot), ""]
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append("%-46s %4d launches %9.3f ms  %5.1f %%" % (k[:46], v[0], v[1], 100 * v[1] / tot))
    seq = []
    for i in [i for i, n in enumerate(order) if n.startswith("lastconv_fwd_tc_kernel")]:
        j, between = i + 1, []
        while j < len(order) and not order[j].startswith("lastconv_bwd"):
            between.append(order[j])
            j += 1
        seq.append(between)
    out += ["", "launches between lastconv_fwd_tc_kernel and the following lastconv_bwd* kernel, per occurrence: %s" % seq,
            "(the first occurrence is the eager warm-up pass, in which the loss workspace is allocated and zeroed once; the "
            "captured steps have none)"]
    open(os.path.join(dst, "%s_final_kernel_shares_c4.txt" % rnd), "w").write("\n".join(out) + "\n")

    # ---- ncu summary of the fused first-backward kernel
    rep = os.path.join("gpurun_out", name, "fused.ncu-rep")
    k = list(json.loads(subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True,
                                       cwd=ROOT).stdout).values())[0][0]
    raw = subprocess.run("ncu -i %s --page raw --csv" % rep, shell=True, capture_output=True, text=True, cwd=ROOT).stdout
    rr = list(csv.reader(raw.splitlines()))
    m = dict(zip(rr[0], rr[-1]))
    stalls = []
    for h in rr[0]:
        if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio") and "_not_issued" not in h:
            try:
                stalls.append((float(m[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    nvox, t = 4 * 128 ** 3, k["gpu__time_duration.sum [ms]"] * 1e-3
    dram = k["dram__bytes_read.sum [Gbyte]"] + k["dram__bytes_write.sum [Gbyte]"]
    lines = ["lastconv_bwd_fused_kernel (dfl_lastconv_curl_loss_bwd): ncu --set full --clock-control none, one launch at 4 x 128^3 "
             "(tools/lastconv_bwd_bench.py --profile fused)",
             "raw report: %s (scratch); summarised with tools/ncu_summary.py + `ncu --page raw --csv`" % rep, ""]
    lines += ["%-80s %s" % kv for kv in k.items()]
    lines += ["", "algorithmic bytes: %d voxels x 1048 B = %.2f GB -> %.0f GB/s at this duration; DRAM traffic %.2f GB = %.2f x "
              "algorithmic" % (nvox, nvox * 1048 / 1e9, nvox * 1048 / t / 1e9, dram, dram / (nvox * 1048 / 1e9)),
              "IPC (sm__inst_executed.avg.per_cycle_elapsed) %s; issue slots busy %s %%" % (
                  m.get("sm__inst_executed.avg.per_cycle_elapsed"), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active")),
              "warp stall reasons (warps stalled per issue-active cycle):"]
    lines += ["   %-28s %.2f" % (n, v) for v, n in stalls[:8]]
    lines += ["", "reading: DESIGN.md 4d (per-role accounting of the source page: profiles/r02b_ncu_fused_roles.txt; A/B of the variants:",
              "profiles/r02b_fused_ab.txt).  CUDA-event time of the same launch outside the profiler: %s_final_lastconv_bwd_bench.json." % rnd]
    open(os.path.join(dst, "%s_ncu_lastconv_bwd_fused.txt" % rnd), "w").write("\n".join(lines) + "\n")
    for f in sorted(os.listdir(src)):
        if f.startswith("bench_") and f.endswith(".json"):
            try:
                d = json.loads(open(os.path.join(src, f)).read().strip().splitlines()[-1])
                print("%-44s %10.2f %s  e2e %10.2f  %8.2f ms/step" % (f, d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"]))
            except Exception as e:
                print(f, "UNREADABLE", e)


if __name__ == "__main__":
    main()
