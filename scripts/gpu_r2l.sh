#!/bin/bash
# Round 2, call L: swizzled K-major operand form (tests + A/B timings), GPU sample pipeline tests, headline-size layer tests,
# dgrad timings with the read-modify-write epilogue, headline bench with both operand forms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x -k "swizzle128 or producer_forms" 2>&1 | tail -8 > gpurun_out/r2l_pytest_sw128.txt
timeout 900 python -m pytest tests/test_gpu_augment.py -q 2>&1 | tail -15 > gpurun_out/r2l_pytest_augment.txt
timeout 1200 python -m pytest tests/test_gpu_parity2.py -q -k headline 2>&1 | tail -15 > gpurun_out/r2l_pytest_headline.txt
for sw in 0 1; do
  for c in 3dconv_up4 fianl_conv dense2_x2 dense2_x1 conv_up4 dense4_x2; do
    for ps in fprop dgrad; do
      echo "== sw128=$sw $c $ps x3"; HDN_TC_SW128=$sw timeout 180 python scripts/prof_conv.py $c $ps 5 2 2>&1 | tail -1
    done
  done
done > gpurun_out/r2l_sw128_times.txt 2>&1
HDN_TC_SW128=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2l_bench_sw0.json 2> gpurun_out/r2l_bench_sw0_err.txt
HDN_TC_SW128=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2l_bench_sw1.json 2> gpurun_out/r2l_bench_sw1_err.txt
du -sk gpurun_out > gpurun_out/r2l_status.txt
