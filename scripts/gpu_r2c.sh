#!/bin/bash
# Round 2, call C: staging-race fix (sliding window tests), gradient error map at 128x128x8, tc2 MMA issue order A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity2.py -q -k "sliding or dp_reduce or moving or wce" 2>&1 | tail -30 > gpurun_out/r2c_pytest_sel.txt
for ord in 0 1; do
  for c in fianl_conv 3dconv_up4 dense2_x2; do
    echo "== $c wgrad tc2 HDN_TC2_ORDER=$ord"
    HDN_TC2_ORDER=$ord timeout 180 python scripts/prof_conv.py $c wgrad 5 1 2>&1 | tail -1
  done
done > gpurun_out/r2c_order_times.txt 2>&1
HDN_GE_SIZE=128 HDN_GE_WORST=30 timeout 1500 python tests/grad_errors.py fp32 mixed > gpurun_out/r2c_grad_errors_128.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -q -k tc2 2>&1 | tail -3 > gpurun_out/r2c_tc2_tests.txt
echo done > gpurun_out/r2c_status.txt
