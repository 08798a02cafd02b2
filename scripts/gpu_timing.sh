#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q --timeout 60 --timeout-method thread 2>&1 | tail -3
for c in 3dconv_up4 dense2_x2 dense2_x1; do
  for w in fprop dgrad; do echo "== $c $w"; timeout 120 python scripts/prof_conv.py $c $w 1 2>&1 | grep "^\[\|TFLOP" | tail -4; done
done | tee gpurun_out/role_timing.txt
