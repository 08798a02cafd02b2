#!/bin/bash
# Round 2, call E: elect_one_sync issue (no per-MMA serialisation loops): unit tests of every tcgen05 kernel, per-layer
# timings against round 1 / call A, a bench line; slice-reuse diagnostic; post-processing tests.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_post.py -q -x 2>&1 | tail -8 > gpurun_out/r2e_pytest_tc_post.txt
for c in fianl_conv 3dconv_up4 conv_up4 dense2_x2 dense2_x1 dense4_x1 3ddense2_x2; do
  for w in fprop dgrad; do
    for prec in 1 2; do echo "== $c $w precision=$prec"; timeout 180 python scripts/prof_conv.py $c $w 5 $prec 2>&1 | tail -1; done
  done
  echo "== $c wgrad precision=1"; timeout 180 python scripts/prof_conv.py $c wgrad 5 1 2>&1 | tail -1
done > gpurun_out/r2e_conv_times.txt 2>&1
timeout 300 python scripts/diag_reuse.py mixed > gpurun_out/r2e_diag_reuse.txt 2>&1
timeout 300 python scripts/diag_reuse.py fp32 >> gpurun_out/r2e_diag_reuse.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench_default.json 2> gpurun_out/r2e_bench_err.txt
timeout 900 python -m pytest tests/test_gpu_models.py -q -x 2>&1 | tail -5 > gpurun_out/r2e_pytest_models.txt
echo done > gpurun_out/r2e_status.txt
