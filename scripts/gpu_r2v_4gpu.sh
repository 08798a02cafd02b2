#!/bin/bash
# Round 2, 4-GPU call (gpurun --gpus 4): the headline step with the device-flag hand-over at world size 4.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_c4_4gpu.json 2> gpurun_out/r02_bench_c4_4gpu_err.txt
echo done > gpurun_out/r2v_status.txt
