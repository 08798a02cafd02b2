#!/bin/bash
# first GPU pass: parity tests, smoke, a reduced-size bench line, ncu launch list (fp32 SIMT path)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --size 256 --cols 8 --batch 1 --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_small.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_small.csv \
   python bench.py --size 128 --cols 8 --batch 1 --steps 1 --warmup 1 --no-cpu --profile-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
