"""Prints one line per bench JSON (value, ms/step, clocks, conv / wgrad TFLOP/s) and the top kernels.  python tools/bench_summary.py DIR [N]"""
import glob, json, os, sys
d = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for f in sorted(glob.glob(os.path.join(d, "bench_*.json"))):
    try:
        j = json.load(open(f))
    except Exception as e:
        print(os.path.basename(f), "ERR", e); continue
    if "roofline" not in j:
        print(os.path.basename(f), j.get("impl", ""), round(j.get("value", 0), 4), j.get("unit", ""), "(no roofline: reference arm)")
        continue
    r = j["roofline"]
    print(os.path.basename(f), round(j["value"], 1), "fields/s", round(j["ms_per_step"], 3), "ms  e2e", round(j["e2e"]["value"], 1), "clk", j["clocks"]["sm_mhz"],
          "conv", round(r["achieved"]), "wgrad", round(r["others"]["wgrad_tc_kernel"]["achieved"]), "stencil GB/s", round(r["others"]["stencil_fused_kernel"]["achieved"]))
    for k, v in list(r.get("kernel_ms", {}).items())[:n]:
        print(f"   {k:28s} {v['ms_per_step']:9.3f} ms  n={v['launches_per_step']}")
