#!/bin/bash
# Round 2, call I: per-CTA statistic sums (no per-tile global double atomics), dgrad epilogue L2 prefetch, classifier kernels,
# staging-pool fix (slice reuse), default TMA level 2: unit tests, layer timings, bench line, model + parity tests.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -6 > gpurun_out/r2i_pytest_tc.txt
for c in fianl_conv 3dconv_up4 conv_up4 dense2_x2 dense2_x1 dense4_x1; do
  for w in fprop dgrad; do echo "== $c $w x3"; timeout 180 python scripts/prof_conv.py $c $w 5 2 2>&1 | tail -1; done
done > gpurun_out/r2i_conv_times.txt 2>&1
for pf in 0 1; do for c in dense2_x1 dense4_x1 fianl_conv; do echo "== $c dgrad x3 HDN_TC_EPIPF=$pf"; HDN_TC_EPIPF=$pf timeout 180 python scripts/prof_conv.py $c dgrad 5 2 2>&1 | tail -1; done; done > gpurun_out/r2i_epipf_times.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_err.txt
timeout 1500 python -m pytest tests/test_gpu_models.py tests/test_gpu_post.py -q -x 2>&1 | tail -4 > gpurun_out/r2i_pytest_models_post.txt
timeout 2400 python -m pytest tests/test_gpu_parity2.py -q -s -k "not headline" 2>&1 | grep -v "^$" | grep -E "passed|failed|FAILED|XFAIL|XPASS|grad errors|2d mixed|slice reuse|forward|Error" | tail -30 > gpurun_out/r2i_parity2.txt
echo done > gpurun_out/r2i_status.txt
