#!/bin/bash
# Round 2, call U: kernel instantiated per (pass, fold, operand path) against the call-I kernel on one box; fold on/off;
# corrected tcgen05.mma issue-rate micro-benchmark; tcgen05 unit tests.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -6 > gpurun_out/r2u_pytest_tc.txt
V=h-denseunet_b200/variants
for v in head callI headfold; do
  lib=$PWD/h-denseunet_b200/libhdn.so; [ $v = callI ] && lib=$PWD/$V/libhdn_callI.so
  f=0; [ $v = headfold ] && f=1
  for c in 3dconv_up4 fianl_conv dense2_x2 dense2_x1 conv_up4; do
    for ps in fprop dgrad; do
      echo "== variant=$v $c $ps x3"; HDN_TC_X3FOLD=$f HDN_LIB=$lib timeout 180 python scripts/prof_conv.py $c $ps 5 2 2>&1 | tail -1
    done
  done
done > gpurun_out/r2u_variants.txt 2>&1
HDN_TC_X3FOLD=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2u_bench_fold0.json 2> gpurun_out/r2u_bench_fold0_err.txt
HDN_TC_X3FOLD=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2u_bench_fold1.json 2> gpurun_out/r2u_bench_fold1_err.txt
du -sk gpurun_out > gpurun_out/r2u_status.txt
