#!/bin/bash
# builds A/B variants of the fused first-backward kernel into deep-fluids_b200/lib/variants/ (select with DFL_LIB_PATH):
#   usage: tools/build_fused_variants.sh NAME "-DMACRO ..." [NAME2 "-D..."] ...
set -e
cd "$(dirname "$0")/../deep-fluids_b200"
make -j8 >/dev/null
mkdir -p lib/variants build/var
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
OTHERS=$(ls build/*.o | grep -v dfl_lastconv_bwd_fused.o)
while [ $# -ge 2 ]; do
  nvcc $F $2 -c csrc/dfl_lastconv_bwd_fused.cu -o build/var/fused_$1.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/lib_fused_$1.so build/var/fused_$1.o $OTHERS -lcudart_static -ldl -lpthread -lrt
  shift 2
done
ls lib/variants
