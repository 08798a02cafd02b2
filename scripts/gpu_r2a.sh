#!/bin/bash
# Round 2, call A: tc2 weight gradient (bf16 pre-pass + TMA + tcgen05) -- correctness of both shared-memory layouts
# against the fp32 FMA kernel and the first-generation kernel, then per-layer timings gen1 / tc2-planes / tc2-sw128.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_env.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py -q -k "tc2" 2>&1 | tail -40 > gpurun_out/r2a_tc2_tests.txt
for c in fianl_conv 3dconv_up4 conv_up4 dense2_x2 3ddense2_x2 dense4_x2; do
  for cfg in "0 0" "1 0" "1 1"; do
    set -- $cfg
    echo "== $c wgrad HDN_WGRAD_TC2=$1 HDN_TC2_LAYOUT=$2"
    HDN_WGRAD_TC2=$1 HDN_TC2_LAYOUT=$2 timeout 180 python scripts/prof_conv.py $c wgrad 5 1 2>&1 | tail -2
  done
done > gpurun_out/r2a_wgrad_times.txt 2>&1
echo done > gpurun_out/r2a_status.txt
