#!/bin/bash
# A/B of fused first-backward kernel variants (built by tools/build_fused_variants.sh) on one box:
#   usage (under gpurun): bash tools/fused_ab.sh OUTDIR [variant names ...]     ("main" = the library in lib/)
o=gpurun_out/$1; shift; mkdir -p $o
for v in "$@"; do
  if [ "$v" = main ]; then lib=$PWD/deep-fluids_b200/lib/libdeepfluids_b200.so; else lib=$PWD/deep-fluids_b200/lib/variants/lib_fused_$v.so; fi
  echo "== $v"
  DFL_LIB_PATH=$lib timeout 300 python tools/lastconv_bwd_bench.py --fused-only --json $o/bwd_bench_$v.json 2>&1 | tail -2
done
