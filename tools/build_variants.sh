#!/bin/bash
# builds A/B variants of the conv kernel's MMA-issue path into deep-fluids_b200/lib/variants/ (select with DFL_LIB_PATH)
set -e
cd "$(dirname "$0")/../deep-fluids_b200"
mkdir -p lib/variants build/var
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
OTHERS="build/dfl_api.o build/dfl_stencil.o build/dfl_stencil3_lean.o build/dfl_wgrad_tc.o build/dfl_lastconv_tc.o build/dfl_lastconv_fwd_tc.o build/dfl_edge.o"
for v in "0 0" "0 1" "1 0" "1 1"; do
  set -- $v
  nvcc $F -DDFL_MMA_ISSUE=$1 -DDFL_DESC_INC=$2 -c csrc/dfl_conv_tc.cu -o build/var/conv_$1$2.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/lib_$1$2.so build/var/conv_$1$2.o $OTHERS -lcudart_static -ldl -lpthread -lrt
done
ls -la lib/variants
