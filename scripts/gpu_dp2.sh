#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --batch 1 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_2gpu.log 2>&1
echo "bench rc=$?"
grep -v "^W0\|^\[W\|^\*\*\*" gpurun_out/bench_2gpu.log | grep -B2 -A25 "Traceback" | head -60
grep '"metric"' gpurun_out/bench_2gpu.log | tee gpurun_out/bench_512x48_b1_2gpu.json | cut -c1-900
