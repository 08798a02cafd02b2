#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc_kernel -s 2 -c 1 -o gpurun_out/prof_up4_wgrad python scripts/prof_conv.py 3dconv_up4 wgrad 1 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_up4_fprop2 python scripts/prof_conv.py 3dconv_up4 fprop 1 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
