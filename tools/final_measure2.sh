#!/bin/bash
# Lean one-GPU evidence batch (last session of round 2): GPU test suite, smoke, the bench lines, the first-backward kernel alone,
# the ncu launch list of the default bench command, one `ncu --set full` capture of the fused first-backward kernel, the three
# sanitizers on that kernel's tests.  Output: gpurun_out/$1/ (default r2b).     usage (under gpurun): bash tools/final_measure2.sh [outdir]
o=gpurun_out/${1:-r2b}; mkdir -p $o
python -m pytest tests -m gpu -q -s > $o/pytest_gpu.log 2>&1; tail -3 $o/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/smoke.log 2>&1; tail -2 $o/smoke.log
python bench.py > $o/bench_c4_n1.json 2> $o/bench_c4_n1.err
for w in c3 c5 c2; do python bench.py --workload $w --no-cpu > $o/bench_${w}_n1.json 2> $o/bench_${w}_n1.err; done
python bench.py --workload c2 --precision bf16 --no-cpu > $o/bench_c2_bf16_n1.json 2> $o/bench_c2_bf16_n1.err
python tools/lastconv_bwd_bench.py --json $o/lastconv_bwd_bench.json > $o/lastconv_bwd_bench.txt 2>&1; cat $o/lastconv_bwd_bench.txt
for f in $o/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-32s %10.2f %s  e2e %10.2f  %.2f ms/step  clk %s" % (sys.argv[1].split("/")[-1], d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], "UNREADABLE", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_default_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu > $o/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lastconv_bwd_fused -c 1 -s 1 -o $o/fused python tools/lastconv_bwd_bench.py --profile fused > $o/ncu_fused.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 150 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_fused_bwd.py -x -q -k "shape0 or shape1 or shape3 or optional" > $o/${tool}_fused_bwd.log 2>&1
  echo "== $tool fused_bwd: $(grep -E 'passed|failed|error' $o/${tool}_fused_bwd.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK SUMMARY' $o/${tool}_fused_bwd.log | tail -1)"
done | tee $o/sanitizer_summary.txt
