#!/bin/bash
# Round 2, call J: per-role cycle counters (-DHDN_TC_TIMING) of the fprop / dgrad kernel in its current form.
mkdir -p gpurun_out
HDN_NVCC_EXTRA=-DHDN_TC_TIMING python -c "import h_denseunet_b200._lib as L; L.build(force=True)" > gpurun_out/r2j_build.log 2>&1
for cw in "fianl_conv fprop" "fianl_conv dgrad" "3dconv_up4 fprop" "dense2_x2 fprop" "dense2_x1 fprop" "dense2_x1 dgrad" "dense4_x1 dgrad" "conv_up4 dgrad"; do
  set -- $cw
  for prec in 1 2; do
    echo "== $1 $2 precision=$prec"
    timeout 120 python scripts/prof_conv.py $1 $2 1 $prec 2>&1 | grep "^\[\|TFLOP" | tail -8
  done
done > gpurun_out/r2j_role_timing.txt 2>&1
python -c "import h_denseunet_b200._lib as L; L.build(force=True)" >> gpurun_out/r2j_build.log 2>&1
echo done > gpurun_out/r2j_status.txt
