#!/bin/bash
# Per-role cycle counters (-DHDN_TC_TIMING) of the tcgen05 kernels on representative convolutions: where a stage's
# cycles go (copy wait / barrier / issue / operand-ring wait / transform, MMA issuer waits, epilogue).  Rebuilds the
# library with the counters on the box (seconds), bf16 and bf16x3, transform forms 0 and 2.
mkdir -p gpurun_out
HDN_NVCC_EXTRA=-DHDN_TC_TIMING python -c "import h_denseunet_b200._lib as L; L.build(force=True)" > gpurun_out/timing_build.log 2>&1
for fx in 2 0; do
  for c in dense2_x1 dense4_x1 dense2_x2 3dconv_up4 fianl_conv; do
    for w in fprop dgrad wgrad; do
      for prec in 1 2; do
        echo "== $c $w precision=$prec HDN_TC_FASTX=$fx"
        HDN_TC_FASTX=$fx timeout 120 python scripts/prof_conv.py $c $w 1 $prec 2>&1 | grep "^\[\|TFLOP" | tail -6
      done
    done
  done
done | tee gpurun_out/role_timing.txt
python -c "import h_denseunet_b200._lib as L; L.build(force=True)" >> gpurun_out/timing_build.log 2>&1
