#!/bin/bash
# Round 2, call D: bf16x3 with IEEE-half tails (correctness + gradient error map), tc2 MMA-rate experiments, slice reuse diff,
# post-processing tests, the tcgen05 unit tests of the whole file.
mkdir -p gpurun_out
python -c "import h_denseunet_b200._lib as L; L.build()" > gpurun_out/r2d_build.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_tc.py -q -s -k "half_tails" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2d_half_tails.txt
timeout 1200 python -m pytest tests/test_gpu_post.py tests/test_gpu_parity2.py -q -s -k "post or sliding" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2d_post_sliding.txt
for dbg in 0 1 3 5; do
  for c in fianl_conv dense2_x2; do
    echo "== $c wgrad tc2 HDN_TC2_DEBUG=$dbg"
    HDN_TC2_DEBUG=$dbg timeout 180 python scripts/prof_conv.py $c wgrad 5 1 2>&1 | tail -1
  done
done > gpurun_out/r2d_tc2_debug_times.txt 2>&1
HDN_TC_TAIL16=1 HDN_GE_SIZE=128 HDN_GE_WORST=12 timeout 1500 python tests/grad_errors.py mixed > gpurun_out/r2d_grad_errors_128_tail16.txt 2>&1
timeout 2400 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -8 > gpurun_out/r2d_pytest_tc_all.txt
echo done > gpurun_out/r2d_status.txt
