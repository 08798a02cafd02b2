#!/bin/bash
set -x
mkdir -p gpurun_out
for c in 3dconv_up4 conv_up4 fianl_conv dense2_x2 dense2_x1 dense4_x1 dense4_x2 3ddense2_x2; do
  for w in fprop dgrad wgrad; do timeout 120 python scripts/prof_conv.py $c $w 3 2>&1 | tail -1; done
done | tee gpurun_out/conv_times.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_up4_fprop python scripts/prof_conv.py 3dconv_up4 fprop 1 > gpurun_out/ncu1.log 2>&1; tail -3 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_d4x1_dgrad python scripts/prof_conv.py dense4_x1 dgrad 1 > gpurun_out/ncu2.log 2>&1; tail -3 gpurun_out/ncu2.log
