#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -q --timeout 300 --timeout-method thread > gpurun_out/pytest_models.log 2>&1; tail -5 gpurun_out/pytest_models.log
timeout 600 python bench.py --batch 1 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_512x48_b1.json
