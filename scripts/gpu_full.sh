#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --batch 1 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_512x48_b1.json
timeout 900 python bench.py --batch 2 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_512x48_b2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 2600 --csv --log-file gpurun_out/launches_512x48_b1.csv \
   python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu --profile-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
