#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -q --timeout 300 --timeout-method thread > gpurun_out/pytest_models.log 2>&1; tail -12 gpurun_out/pytest_models.log | cut -c1-300
timeout 600 python bench.py --batch 1 --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_512x48_b1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_512x48_b1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d.get('cpu_baseline'))"
