#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q --timeout 60 --timeout-method thread 2>&1 | tail -3
for c in 3dconv_up4 fianl_conv dense2_x2 dense2_x1 dense4_x1 3ddense2_x2; do
  for w in dgrad wgrad; do timeout 120 python scripts/prof_conv.py $c $w 3 2>&1 | tail -1; done
done | tee gpurun_out/conv_times_v4.txt
