#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q --timeout 60 --timeout-method thread > gpurun_out/pytest_tc.log 2>&1; tail -25 gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_models.py -q --timeout 300 --timeout-method thread > gpurun_out/pytest_models.log 2>&1; tail -8 gpurun_out/pytest_models.log
timeout 300 python bench.py --size 512 --cols 8 --batch 1 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_512x8_bf16.json
timeout 600 python bench.py --batch 1 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_512x48_b1.json
