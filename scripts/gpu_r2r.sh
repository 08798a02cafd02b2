#!/bin/bash
# Round 2, call R: per-role cycle counters of the fprop / dgrad kernel after the issuer change (timing build, variants/).
mkdir -p gpurun_out
lib=$PWD/h-denseunet_b200/variants/libhdn_timing.so
for cw in "fianl_conv fprop" "fianl_conv dgrad" "3dconv_up4 fprop" "dense2_x2 fprop" "dense2_x1 fprop" "dense2_x1 dgrad" "conv_up4 dgrad"; do
  set -- $cw
  for prec in 1 2; do
    echo "== $1 $2 precision=$prec"
    HDN_LIB=$lib timeout 120 python scripts/prof_conv.py $1 $2 1 $prec 2>&1 | grep "^\[\|TFLOP" | tail -8
  done
done > gpurun_out/r2r_role_timing.txt 2>&1
echo done > gpurun_out/r2r_status.txt
