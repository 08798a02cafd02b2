#!/bin/bash
# tcgen05 kernel unit tests first (under a short timeout: a deadlocked kernel must not eat the budget)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q --timeout 120 --timeout-method thread 2>&1 | tail -60 | tee gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_models.py -q 2>&1 | tail -30 | tee gpurun_out/pytest_models.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --size 256 --cols 8 --batch 1 --steps 2 --warmup 3 --precision fp32 --no-cpu 2>&1 | tail -2 | tee gpurun_out/bench_small_fp32.json
timeout 600 python bench.py --size 256 --cols 8 --batch 1 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -2 | tee gpurun_out/bench_small_bf16.json
