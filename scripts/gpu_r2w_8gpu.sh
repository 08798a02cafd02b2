#!/bin/bash
# Round 2, 8-GPU call (gpurun --gpus 8): BASELINE config 5 (sliding-window inference of a 512-slice volume, z-sharded over 8 GPUs,
# with and without 2-D slice reuse) and the headline step at world size 8.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
timeout 600 $TR bench.py --gpus 8 --config c5 --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_bench_c5_8gpu.json 2> gpurun_out/r02_bench_c5_8gpu_err.txt

echo done > gpurun_out/r2w_status.txt
