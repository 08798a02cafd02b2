#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus 2): device-flag hand-over of the data-parallel exchange against the NCCL baseline,
# the headline bench on 2 GPUs, the z-sharded sliding-window inference (config 5) on 2 GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/dp_check.py > gpurun_out/r2h_dp_check_flags.txt 2>&1
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2h_bench_c4_2gpu.json 2> gpurun_out/r2h_bench_c4_2gpu_err.txt
timeout 1200 $TR bench.py --gpus 2 --config c5 --depth 128 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2h_bench_c5_2gpu_z128.json 2> gpurun_out/r2h_bench_c5_2gpu_err.txt
echo done > gpurun_out/r2h_status.txt
