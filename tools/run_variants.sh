#!/bin/bash
for v in 00 01 10 11; do
  r=$(DFL_LIB_PATH=$PWD/deep-fluids_b200/lib/variants/lib_$v.so timeout 200 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f fields/s %.2f ms conv_frac %.3f conv_TF %.0f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['clocks']['sm_mhz']))")
  echo "variant issue/inc=$v : $r"
done
