#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q --timeout 120 --timeout-method thread > gpurun_out/pytest_tc.log 2>&1; tail -5 gpurun_out/pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_models.py -q > gpurun_out/pytest_models.log 2>&1; tail -8 gpurun_out/pytest_models.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_256x8_bf16.csv \
   python bench.py --size 256 --cols 8 --batch 1 --steps 1 --warmup 1 --no-cpu --profile-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
timeout 600 python bench.py --size 512 --cols 8 --batch 1 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_512x8_bf16.json
