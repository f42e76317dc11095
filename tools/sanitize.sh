#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck) over the hand-rolled tcgen05 / TMA / mbarrier / named-barrier pipelines
# at small shapes: per-kernel conv, wgrad, output conv forward + un-fused backward, the fused first-backward kernel, the lean
# stencil, the phase-decomposed layer, the paired-brick per-tap kernel.  Writes one summary per tool under gpurun_out/sanitizer/.   usage: tools/sanitize.sh
out=gpurun_out/sanitizer; mkdir -p $out
SEL='test_conv3d_fwd_lrelu or test_conv2d_fwd_lrelu or test_conv_dgrad_and_wgrad_vs_autograd or test_conv3d_residual_upsample_epilogue or lastconv_tensorcore_fwd_and_fused_bwd or test_golden_stencil3d or test_stencil3d_vs_oracle'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -x -q -k "$SEL" > $out/${tool}_kernels.log 2>&1
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_fused_bwd.py -x -q -k "shape0 or shape1 or shape3 or optional" > $out/${tool}_fused_bwd.log 2>&1
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_phase.py -x -q -k "forward_kernel" > $out/${tool}_phase.log 2>&1
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_encoder.py -x -q -k "paired" > $out/${tool}_paired.log 2>&1
  for f in kernels fused_bwd phase paired; do
    echo "== $tool $f: $(grep -E 'passed|failed|error' $out/${tool}_$f.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK SUMMARY' $out/${tool}_$f.log | tail -1)"
  done
done | tee $out/summary.txt
