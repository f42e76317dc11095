#!/bin/bash
# sweeps the wgrad slab heuristic constant (DFL_WGRAD_SLAB_K; 1e9 = always the SM-count cap) on the 2D and 3D benches
for k in 0.5 1 2 4 8 1000000000; do
  for wl in "c2 --precision bf16" "c2" "c3"; do
    v=$(DFL_WGRAD_SLAB_K=$k timeout 200 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f fields/s %.3f ms wgrad_frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['others']['wgrad_tc_kernel']['frac']))")
    echo "K=$k workload=$wl : $v"
  done
done
