#!/bin/bash
# Round 2, call F: quad dgrad epilogue + tc2 weight gradient for the 1x1x1 layers: unit tests, A/B timings, bench line;
# slice-reuse diagnostic with a dirty allocator / poisoned buffers; the round-2 parity file.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -8 > gpurun_out/r2f_pytest_tc.txt
for c in dense2_x1 dense4_x1 dense2_x2 conv_up4 fianl_conv; do
  for e4 in 0 1; do echo "== $c dgrad x3 HDN_TC_EPI4=$e4"; HDN_TC_EPI4=$e4 timeout 180 python scripts/prof_conv.py $c dgrad 5 2 2>&1 | tail -1; done
done > gpurun_out/r2f_epi4_times.txt 2>&1
for c in dense2_x1 dense4_x1; do
  for t2 in 0 1; do echo "== $c wgrad HDN_WGRAD_TC2=$t2"; HDN_WGRAD_TC2=$t2 timeout 180 python scripts/prof_conv.py $c wgrad 5 1 2>&1 | tail -1; done
done > gpurun_out/r2f_wgrad1x1_times.txt 2>&1
HDN_DIAG_DIRTY=1 timeout 300 python scripts/diag_reuse.py mixed > gpurun_out/r2f_diag_reuse_dirty.txt 2>&1
HDN_DIAG_DIRTY=1 HDN_POISON=1 timeout 300 python scripts/diag_reuse.py mixed > gpurun_out/r2f_diag_reuse_poison.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_err.txt
timeout 2400 python -m pytest tests/test_gpu_parity2.py -q -s -k "not headline" 2>&1 | grep -v "^$" | grep -E "passed|failed|FAILED|XFAIL|XPASS|grad errors|2d mixed|slice reuse|forward|K ours|Error" | tail -40 > gpurun_out/r2f_parity2.txt
echo done > gpurun_out/r2f_status.txt
