#!/bin/bash
# Round 2, call G: TMA mode of the fprop / dgrad kernel (unit tests of all producer forms, A/B timings, bench line);
# bisect of the in-process slice-reuse mismatch.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -8 > gpurun_out/r2g_pytest_tc.txt
for c in fianl_conv 3dconv_up4 conv_up4 dense2_x2 3ddense2_x2 dense2_x1 dense4_x1; do
  for w in fprop dgrad; do
    for lvl in 0 2; do echo "== $c $w x3 HDN_TC_TMA=$lvl"; HDN_TC_TMA=$lvl timeout 180 python scripts/prof_conv.py $c $w 5 2 2>&1 | tail -1; done
  done
done > gpurun_out/r2g_tma_times.txt 2>&1
for lvl in 1 2; do HDN_TC_TMA=$lvl timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2g_bench_tma$lvl.json 2> gpurun_out/r2g_bench_err$lvl.txt; done
for k in "sliding" "moving or sliding" "wce or sliding" "unet2d_training or sliding" "forward_full_shape and 224 or sliding"; do
  echo "== -k '$k'"; timeout 900 python -m pytest tests/test_gpu_parity2.py -q -s -k "$k" 2>&1 | grep -E "slice reuse|passed|failed" | tail -3
done > gpurun_out/r2g_reuse_bisect.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_post.py -q -x 2>&1 | tail -4 > gpurun_out/r2g_pytest_models_post.txt
echo done > gpurun_out/r2g_status.txt
