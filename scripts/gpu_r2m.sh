#!/bin/bash
# Round 2, call M: quad fprop epilogue + reduction-based dgrad epilogue restored (tests, timings), folded bf16x3 (tests, A/B),
# headline bench with fold 0/1.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -8 > gpurun_out/r2m_pytest_tc.txt
for f in 0 1; do
  for c in 3dconv_up4 fianl_conv dense2_x2 dense2_x1 conv_up4 dense4_x1 3ddense2_x2 3ddense2_x1; do
    for ps in fprop dgrad; do
      echo "== x3fold=$f $c $ps x3"; HDN_TC_X3FOLD=$f timeout 180 python scripts/prof_conv.py $c $ps 5 2 2>&1 | tail -1
    done
  done
done > gpurun_out/r2m_fold_times.txt 2>&1
HDN_TC_X3FOLD=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2m_bench_fold0.json 2> gpurun_out/r2m_bench_fold0_err.txt
HDN_TC_X3FOLD=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2m_bench_fold1.json 2> gpurun_out/r2m_bench_fold1_err.txt
HDN_TC_X3FOLD=1 timeout 900 python -m pytest tests/test_gpu_models.py -q -x 2>&1 | tail -6 > gpurun_out/r2m_pytest_models_fold1.txt
du -sk gpurun_out > gpurun_out/r2m_status.txt
