#!/bin/bash
# Round 2, call N: A/B of kernel variants on one box (scripts/build_variants.sh): today's kernel, without the fold code, with
# the scalar fprop epilogue, both, and the kernel files of the commit behind the 643 ms bench line (call I).
mkdir -p gpurun_out
timeout 300 scripts/micro/mma_rate 148 > gpurun_out/r2n_mma_rate_148.txt 2>&1
timeout 120 scripts/micro/mma_rate 1 > gpurun_out/r2n_mma_rate_1.txt 2>&1
V=h-denseunet_b200/variants
for rep in 1 2; do
for v in head callI nofold scalar nofold_scalar; do
  lib=$PWD/$V/libhdn_$v.so; [ $v = head ] && lib=$PWD/h-denseunet_b200/libhdn.so
  for c in 3dconv_up4 fianl_conv dense2_x2 dense2_x1 conv_up4; do
    for ps in fprop dgrad; do
      echo "== rep=$rep variant=$v $c $ps x3"; HDN_LIB=$lib timeout 180 python scripts/prof_conv.py $c $ps 5 2 2>&1 | tail -1
    done
  done
done
done > gpurun_out/r2n_variants.txt 2>&1
HDN_LIB=$PWD/$V/libhdn_callI.so timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2n_bench_callI.json 2> gpurun_out/r2n_bench_callI_err.txt
HDN_TC_X3FOLD=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2n_bench_head_fold1.json 2> gpurun_out/r2n_bench_head_fold1_err.txt
du -sk gpurun_out > gpurun_out/r2n_status.txt
