#!/bin/bash
# Round 2, call K: read-modify-write dgrad epilogue (tests + timings), then the evidence captures (scripts/gpu_evidence.sh).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_tc.py -q -x 2>&1 | tail -6 > gpurun_out/r2k_pytest_tc.txt
for c in dense2_x1 dense4_x1 dense2_x2 conv_up4 fianl_conv 3dconv_up4; do echo "== $c dgrad x3"; timeout 180 python scripts/prof_conv.py $c dgrad 5 2 2>&1 | tail -1; done > gpurun_out/r2k_dgrad_times.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2k_bench_default.json 2> gpurun_out/r2k_bench_err.txt
bash scripts/gpu_evidence.sh > gpurun_out/r2k_evidence.log 2>&1
echo done > gpurun_out/r2k_status.txt
